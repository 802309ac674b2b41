// kernels.cuh -- the pipeline's kernels around core.h:
//   keys      (contig, pos) -> one linear sortable coordinate (+ input validation)
//   segment   chain-linkage boundary votes (max_cluster_bias_*) + ordered compaction of the
//             clusters that can reach min_support
//   cluster   one warp / one CTA / one CTA with global scratch per kept cluster -> core.h
//   order     candidates into the reference's emission order
//   genotype  windows binned on the linear coordinate, ONE streaming pass over the reads table,
//             cal_GL through the host-built libm table
#pragma once
#include "core.h"
#include "devprims.cuh"
#include "radix.cuh"
#include <type_traits>

namespace csv {

static constexpr int WARP_M = 128;    // largest cluster a warp-sized team handles (shared memory)
static constexpr int BLOCK_M = 2048;  // largest cluster a CTA-sized team handles in shared memory
static constexpr int CL_THREADS = 256;

// ------------------------------------------------------------------------------------------
// description of one SV type's sorted domain, passed by value to the kernels
// ------------------------------------------------------------------------------------------
struct TypeJob {
    int svtype;
    int64_t n_host;          // upper bound of the sorted-domain size
    const uint32_t* n_dev;   // actual size when it is only known on the device (small types)
    // INDEL: linear keys in sorted order
    const uint32_t* keys32;
    const uint64_t* keys64;
    IndelView iv;
    // DUP / INV / TRA: fully sorted, de-duplicated columns
    SortedView sv;
    ClusterParams cp;
    uint32_t kslot_base;     // first global kept-cluster slot of this type
    const uint32_t* kept_start;
    uint32_t* big_list;
    uint32_t* giant_list;
    uint32_t big_cap, giant_cap;
    char* giant_arena;
    int small_path;          // INS / DEL: clusters of <= 32 members go through the register kernel (k_cluster_small) first
    uint32_t* rest_list;     // ... which lists the others here (kept-cluster ordinals); null: the general kernel takes every cluster
    const uint32_t* n_rest;
    uint2* small_list;       // (ordinal, members) of the clusters of <= 32 members, written by k_select_heads in record mode
    const uint32_t* n_small;
};

__device__ __forceinline__ int64_t job_n(const TypeJob& J) { return J.n_dev ? (int64_t)*J.n_dev : J.n_host; }

// element i (> 0) belongs to the same chain cluster as element i-1
__device__ __forceinline__ bool job_linked(const TypeJob& J, int64_t i) {
    const int32_t bias = J.cp.bias;
    switch (J.svtype) {
        case CSV_DEL:
        case CSV_INS:  // resolveINDEL.py:61,271 -- contigs are padded by > bias in the linear key
            if (J.keys32) return (J.keys32[i] - J.keys32[i - 1]) <= (uint32_t)bias;
            return (J.keys64[i] - J.keys64[i - 1]) <= (uint64_t)bias;
        case CSV_DUP:  // resolveDUP.py:35
            return J.sv.chrom[i] == J.sv.chrom[i - 1] && !(J.sv.a[i] - J.sv.a[i - 1] > bias);
        case CSV_INV:  // resolveINV.py:56
            return J.sv.chrom[i] == J.sv.chrom[i - 1] && J.sv.c[i] == J.sv.c[i - 1] && !(J.sv.a[i] - J.sv.a[i - 1] > bias) &&
                   !(J.sv.b[i] - J.sv.b[i - 1] > bias);
        default:       // TRA resolveTRA.py:41,65
            return J.sv.chrom[i] == J.sv.chrom[i - 1] && J.sv.c[i] == J.sv.c[i - 1] && !(J.sv.a[i] - J.sv.a[i - 1] > bias);
    }
}

// predicate of the segment step: i starts a chain cluster with at least min_support members
struct HeadPred {
    TypeJob J;
    __device__ __forceinline__ bool operator()(int64_t i) const {
        if (i > 0 && job_linked(J, i)) return false;
        const int64_t n = job_n(J);
        const int need = J.cp.min_support;
        int cnt = 1;
        int64_t j = i + 1;
        while (cnt < need && j < n && job_linked(J, j)) { cnt++; j++; }
        return cnt >= need;
    }
};

// Segment step, specialised: the chain-linkage votes of a 2048-element tile are taken with
// coalesced loads and kept as a bit mask in shared memory (one __ballot_sync per 32 elements);
// "i starts a cluster of >= min_support members" is then a bit test: link[i] == 0 and
// link[i+1 .. i+min_support-1] all 1.  Ordered compaction as in k_select.
static constexpr int HEAD_MAX_NEED_WORDS = 64;  // supports min_support up to ~2000 via the mask; above -> generic path
// INS / DEL record mode (MR.rec != nullptr): the warps of the CTA then walk the kept clusters whose head lies in the
// tile and gather ONE 16 B record (+ column c of INS) per member through the input index, written at the member's
// sorted position -- the cluster kernels read a cluster's members as one contiguous range instead of gathering four
// or five columns per member.  Only members of kept clusters are touched (a fifth of the filter's survivors on 30x ONT).
struct MemberRec {
    IndelRec* rec; int32_t* recc;
    const int32_t *a, *b, *rid, *c;
    const uint32_t* sidx;
    // classification of the kept clusters by size while their records are gathered (null: not wanted)
    uint2* small_list; uint32_t* n_small;    // (kept ordinal, members) for <= 32 members
    uint32_t* rest_list; uint32_t* n_rest;   // kept ordinal for the others
};
__global__ void __launch_bounds__(SEL_THREADS) k_select_heads(TypeJob J, uint32_t* out, uint32_t out_cap, uint32_t* out_count,
                                                              TileSync ts, uint32_t* status_word, uint32_t overflow_bit, MemberRec MR) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    constexpr int WORDS = SEL_TILE / 32;
    __shared__ uint32_t s_link[WORDS + HEAD_MAX_NEED_WORDS + 2];
    __shared__ uint32_t s_warp[9];
    __shared__ uint32_t s_tile, s_excl;
    __shared__ uint32_t s_nheads;
    __shared__ uint16_t s_heads[SEL_TILE];
    const int64_t n = job_n(J);
    const uint32_t gen = ts_gen(ts);
    const int need = J.cp.min_support;
    const int halo_words = (need + 31) / 32 + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    while (true) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ts.ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        const int64_t base = (int64_t)tile * SEL_TILE;
        if (base >= n) break;
        // link bits for [base, base + SEL_TILE + 32*halo_words): votes are evaluated in batches
        // of 8 words per warp so that the key loads of a batch are all in flight together
        for (int w0 = warp; w0 < WORDS + halo_words; w0 += 8 * (SEL_THREADS / 32)) {
            bool l[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int w = w0 + u * (SEL_THREADS / 32);
                const int64_t i = base + (int64_t)w * 32 + lane;
                l[u] = (w < WORDS + halo_words && i > 0 && i < n) ? job_linked(J, i) : false;
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int w = w0 + u * (SEL_THREADS / 32);
                const uint32_t m = __ballot_sync(0xffffffffu, l[u]);
                if (lane == 0 && w < WORDS + halo_words) s_link[w] = m;
            }
        }
        __syncthreads();
        const int p0 = threadIdx.x * SEL_ITEMS;
        uint32_t flags = 0, cnt = 0;
#pragma unroll
        for (int j = 0; j < SEL_ITEMS; j++) {
            const int p = p0 + j;
            if (base + p >= n) continue;
            if (s_link[p >> 5] >> (p & 31) & 1u) continue;   // chained to its predecessor: not a head
            // bits p+1 .. p+need-1 must all be set
            int remaining = need - 1, q = p + 1;
            bool ok = true;
            while (remaining > 0) {
                const int wi = q >> 5, bo = q & 31;
                const int take = min(32 - bo, remaining);
                const uint32_t mask = (take == 32 ? 0xffffffffu : ((1u << take) - 1u)) << bo;
                if ((s_link[wi] & mask) != mask) { ok = false; break; }
                q += take; remaining -= take;
            }
            if (ok) { flags |= 1u << j; cnt++; }
        }
        uint32_t total;
        const uint32_t local = block_excl_scan_256(cnt, s_warp, &total);
        if (threadIdx.x < 32) {
            const uint32_t ex = lookback_exclusive_warp(ts.status, gen, (int)tile, total);
            if (threadIdx.x == 0) {
                s_excl = ex;
                if (base + SEL_TILE >= n) *out_count = ex + total;
            }
        }
        __syncthreads();
        uint32_t o = s_excl + local;
#pragma unroll
        for (int j = 0; j < SEL_ITEMS; j++) {
            if (flags >> j & 1u) {
                if (o < out_cap) out[o] = (uint32_t)(base + p0 + j);
                else atomicOr(status_word, overflow_bit);
                if (MR.rec) s_heads[local + (o - (s_excl + local))] = (uint16_t)(p0 + j);
                o++;
            }
        }
        if (MR.rec) {
            if (threadIdx.x == 0) s_nheads = total;
            __syncthreads();
            const uint32_t nh = s_nheads;
            uint32_t bad = 0;
            for (uint32_t h = warp; h < nh; h += SEL_THREADS / 32) {
                const int64_t s0 = base + s_heads[h];
                // members s0 .. s0+m-1: 32 links per step until the first missing one
                int64_t done = 1;
                while (true) {
                    const int64_t i = s0 + done + lane;
                    const bool brk = (i >= n) || !job_linked(J, i);
                    const uint32_t bm = __ballot_sync(0xffffffffu, brk);
                    const int take = bm ? __ffs(bm) - 1 : 32;
                    if (lane < take) {   // element i is a member
                        const uint32_t x = MR.sidx[i];
                        IndelRec r;
                        r.a = MR.a[x]; r.b = MR.b[x]; r.rid = MR.rid[x]; r.idx = x;
                        const int32_t c5 = MR.c ? MR.c[x] : 0;
                        if (r.rid < 0 || r.b < 0 || c5 < 0) bad |= ST_NEG_FIELD;
                        *reinterpret_cast<int4*>(&MR.rec[i]) = *reinterpret_cast<const int4*>(&r);
                        if (MR.recc) MR.recc[i] = c5;
                    }
                    if (bm) { done += take; break; }
                    done += 32;
                }
                if (MR.small_list && lane == 0) {   // `done` = members of the cluster; its kept ordinal = s_excl + h
                    if (done <= 32) MR.small_list[atomicAdd(MR.n_small, 1u)] = make_uint2(s_excl + h, (uint32_t)done);
                    else MR.rest_list[atomicAdd(MR.n_rest, 1u)] = s_excl + h;
                }
                if (lane == 0) {   // the head itself
                    const uint32_t x = MR.sidx[s0];
                    IndelRec r;
                    r.a = MR.a[x]; r.b = MR.b[x]; r.rid = MR.rid[x]; r.idx = x;
                    const int32_t c5 = MR.c ? MR.c[x] : 0;
                    if (r.rid < 0 || r.b < 0 || c5 < 0) bad |= ST_NEG_FIELD;
                    *reinterpret_cast<int4*>(&MR.rec[s0]) = *reinterpret_cast<const int4*>(&r);
                    if (MR.recc) MR.recc[s0] = c5;
                }
            }
            if (bad) atomicOr(status_word, bad);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// keys
// ------------------------------------------------------------------------------------------
struct ContigTab {
    const uint64_t* off;   // linear offset of every contig (padded by > max bias), n+1 entries
    const int64_t* len;
    int32_t n;
};

// Density pre-filter (INS/DEL).  A chain cluster with >= min_support members has all of them
// within R = (min_support-1)*bias of each other, so a signature whose +-R neighbourhood holds fewer
// than min_support signatures can never be part of a cluster the reference would keep
// (resolveINDEL.py:62: len(cluster) >= read_count) and dropping it cannot merge or complete any
// other cluster.  Neighbourhood counts come from a coarse bucket histogram (conservative: whole
// buckets).  On 30x ONT noise this removes ~90 % of the signatures BEFORE the sort.
static constexpr int BKT_SHIFT = 8;   // 256 bp buckets
static constexpr int BKT_PAD = 64;   // >= largest neighbourhood radius in buckets

// grouped uploads: chrom[i] = the contig k with off[k] <= i < off[k+1] (four consecutive rows per thread)
__global__ void __launch_bounds__(256) k_expand_contigs(const int64_t* __restrict__ off, int n_contigs, int64_t n, int32_t* __restrict__ chrom) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v * 4 < n; v += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = v * 4;
        int lo = 0, hi = n_contigs;   // last k with off[k] <= i0
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (__ldg(&off[mid]) <= i0) lo = mid; else hi = mid; }
        int k = lo;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t i = i0 + j;
            if (i >= n) break;
            while (__ldg(&off[k + 1]) <= i) k++;   // skips empty contigs; off[n_contigs] == n > i ends it
            chrom[i] = k;
        }
    }
}

// Four consecutive signatures per thread and iteration through 128-bit loads (64 B of loads in flight per
// thread: the kernel is a latency-bound stream); a scalar loop takes the tail / unaligned columns.
template <typename K, bool HIST>
__global__ void __launch_bounds__(256) k_indel_keys(const int32_t* __restrict__ chrom, const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                             const int32_t* __restrict__ rid, int64_t n, int is_ins, ContigTab ct, K* __restrict__ keys,
                             uint32_t* status, uint32_t* __restrict__ bkt) {
    auto one = [&](int32_t c, int32_t raw, int32_t bb, int32_t rr, uint32_t& bad) -> K {
        K key = 0;
        if (c < 0 || c >= ct.n) bad |= ST_BAD_CHROM;
        else {
            const int64_t pos = is_ins ? (raw >> 1) : raw;
            if (raw < 0 || pos > ct.len[c]) bad |= ST_BAD_POS;
            else key = (K)(ct.off[c] + (uint64_t)pos);
        }
        if (rr < 0 || bb < 0) bad |= ST_NEG_FIELD;
        return key;
    };
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((((uintptr_t)chrom) | ((uintptr_t)a) | ((uintptr_t)b) | ((uintptr_t)rid) | ((uintptr_t)keys)) & 15) == 0;
    const int64_t nv = aligned ? (n >> 2) : 0;
    uint32_t bad = 0;
    for (int64_t v = tid; v < nv; v += stride) {
        const int4 c4 = reinterpret_cast<const int4*>(chrom)[v], a4 = reinterpret_cast<const int4*>(a)[v];
        const int4 b4 = reinterpret_cast<const int4*>(b)[v], r4 = reinterpret_cast<const int4*>(rid)[v];
        const K k0 = one(c4.x, a4.x, b4.x, r4.x, bad), k1 = one(c4.y, a4.y, b4.y, r4.y, bad);
        const K k2 = one(c4.z, a4.z, b4.z, r4.z, bad), k3 = one(c4.w, a4.w, b4.w, r4.w, bad);
        if (sizeof(K) == 4) reinterpret_cast<uint4*>(keys)[v] = make_uint4((uint32_t)k0, (uint32_t)k1, (uint32_t)k2, (uint32_t)k3);
        else {
            reinterpret_cast<ulonglong2*>(keys)[2 * v] = make_ulonglong2((uint64_t)k0, (uint64_t)k1);
            reinterpret_cast<ulonglong2*>(keys)[2 * v + 1] = make_ulonglong2((uint64_t)k2, (uint64_t)k3);
        }
        if (HIST) {
            atomicAdd(&bkt[(uint32_t)(k0 >> BKT_SHIFT) + BKT_PAD], 1u); atomicAdd(&bkt[(uint32_t)(k1 >> BKT_SHIFT) + BKT_PAD], 1u);
            atomicAdd(&bkt[(uint32_t)(k2 >> BKT_SHIFT) + BKT_PAD], 1u); atomicAdd(&bkt[(uint32_t)(k3 >> BKT_SHIFT) + BKT_PAD], 1u);
        }
    }
    for (int64_t i = nv * 4 + tid; i < n; i += stride) {
        const K key = one(chrom[i], a[i], b[i], rid[i], bad);
        keys[i] = key;
        if (HIST) atomicAdd(&bkt[(uint32_t)(key >> BKT_SHIFT) + BKT_PAD], 1u);
    }
    if (bad) atomicOr(status, bad);
}

// pass 2 of the density filter: one bit per bucket = "the +-rb bucket neighbourhood holds >= need
// signatures" (a superset of the +-R window of every signature in the bucket).  The bit map is
// 1.5 MB for hg19 and stays cache resident for the per-signature test.
__global__ void __launch_bounds__(256) k_bucket_flags(const uint32_t* __restrict__ bkt, uint32_t n_buckets, int rb, uint32_t need,
                                                      uint32_t* __restrict__ flags, uint32_t* __restrict__ hist_to_clear, int hist_words) {
    if (blockIdx.x == 0)   // the radix digit histograms k_prefilter accumulates into
        for (int i = threadIdx.x; i < hist_words; i += 256) hist_to_clear[i] = 0u;
    // A CTA stages 4096 buckets (+ halo) in shared memory with coalesced loads; every thread then
    // slides the window sum along its 16 consecutive buckets; lane pairs assemble a 32-bit word.
    constexpr int PER = 16, TILE = 256 * PER;
    __shared__ uint32_t s_b[(TILE + 2 * BKT_PAD + 8) * 17 / 16 + 8];
    auto P = [](int i) { return i + (i >> 4); };  // +1 word per 16: threads stride 17 words -> no bank conflicts
    const uint32_t n_tiles = (n_buckets + TILE - 1) / TILE;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = (int64_t)tile * TILE;                 // first bucket of the tile
        // s_b[i] = bkt[t0 - rb + i + PAD], i in [0, TILE + 2*rb + 1)
        for (int i = threadIdx.x; i < TILE + 2 * rb + 1; i += 256) s_b[P(i)] = bkt[t0 - rb + i + BKT_PAD];
        __syncthreads();
        const int l0 = threadIdx.x * PER;                        // local index of the thread's first bucket
        uint32_t sum = 0, mask = 0;
        for (int k = 0; k <= 2 * rb; k++) sum += s_b[P(l0 + k)];
#pragma unroll
        for (int j = 0; j < PER; j++) {
            if (t0 + l0 + j < n_buckets && sum >= need) mask |= 1u << j;
            sum += s_b[P(l0 + j + 2 * rb + 1)] - s_b[P(l0 + j)];
        }
        const uint32_t other = __shfl_down_sync(0xffffffffu, mask, 1);
        const int64_t word = (t0 + l0) >> 5;
        if ((threadIdx.x & 1) == 0 && t0 + l0 < n_buckets) flags[word] = mask | (other << 16);
        __syncthreads();
    }
}

// survivors of the density filter, compacted (order irrelevant: every later tie-break uses the
// original input index).  2048 signatures per CTA iteration, one reservation atomic per iteration.
__global__ void __launch_bounds__(256) k_prefilter(const uint32_t* __restrict__ keys, int64_t n, const uint32_t* __restrict__ flags,
                                                   uint32_t* __restrict__ out_keys, uint32_t* __restrict__ out_idx,
                                                   uint32_t* out_count, uint32_t* __restrict__ bkt, int64_t n_bkt,
                                                   uint32_t* __restrict__ hist, int passes) {
    constexpr int ITEMS = 8;
    __shared__ uint32_t s_warp[10];
    __shared__ uint32_t s_hist[RS_MAX_PASSES * 256];
    // Two chores ride along: the bucket histogram is cleared for the next call (it was consumed by k_bucket_flags; the
    // buffer is all-zero between calls), and the survivors' digit histograms of every radix pass are accumulated here
    // instead of in a separate read of the compacted keys (hist was zeroed by k_bucket_flags).
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_bkt; i += (int64_t)gridDim.x * 256) bkt[i] = 0u;
    for (int i = threadIdx.x; i < passes * 256; i += 256) s_hist[i] = 0;
    __syncthreads();
    const int64_t n_tiles = (n + 256 * ITEMS - 1) / (256 * ITEMS);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * 256 * ITEMS;
        uint32_t key[ITEMS];
        uint32_t passm = 0, cnt = 0;
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const int64_t i = base + j * 256 + threadIdx.x;
            key[j] = i < n ? keys[i] : 0u;
        }
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            const int64_t i = base + j * 256 + threadIdx.x;
            const uint32_t b = key[j] >> BKT_SHIFT;
            const bool pass = i < n && ((__ldg(&flags[b >> 5]) >> (b & 31)) & 1u);
            if (pass) { passm |= 1u << j; cnt++; }
        }
        uint32_t o = block_reserve_256(cnt, out_count, s_warp);
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            if (passm >> j & 1u) {
                out_keys[o] = key[j];
                out_idx[o] = (uint32_t)(base + j * 256 + threadIdx.x);
                o++;
                for (int p = 0; p < passes; p++) atomicAdd(&s_hist[p * 256 + (int)((key[j] >> (8 * p)) & 0xff)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < passes * 256; i += 256) {
        const uint32_t v = s_hist[i];
        if (v) atomicAdd(&hist[i], v);
    }
}

// ------------------------------------------------------------------------------------------
// INS / DEL front end, filter-first: "density filter -> bucket counting sort -> in-bucket order -> member records"
// (replaces: keys pass + compaction + four stable radix passes over (key, index) + the member gathers of the
//  cluster kernels).  The bucket histogram the filter needs anyway IS a counting-sort histogram: a prefix sum over
//  the counts of the flagged buckets gives every flagged bucket its slot range, so ONE scatter pass puts every
//  survivor into bucket order and a tile-local pass orders the few signatures inside each 256-bp bucket.  The order
//  among equal keys is irrelevant: every later tie-break uses the input index.
//  What bounds these kernels on B200 is not DRAM bytes but the rate of UNCOALESCED 4-8 B accesses (one L1 wavefront
//  each, ~1.5e11/s measured): the design minimises those -- per signature one RED and one table look-up, per
//  survivor one returning atomic and one 8 B store, per member of a kept cluster one record gather.
//    k_indel_hist      8 B/sig read (chrom, a)                 + one RED per signature
//    k_bucket_prefix   4 B/bucket read, 4 B/bucket written     flag + slot offset inside the 4096-bucket tile
//    k_scan_small      tile totals -> tile bases               (a few thousand words, one CTA)
//    k_indel_scatter   8 B/sig read + one look-up              -> (key, index) 8 B per survivor, bucket order
//    k_bucket_fixup    8 B/survivor read + written             in-bucket order (+ clears the histogram)
//    k_select_heads    4 B/survivor read (chain votes)         -> 16-20 B record per member of a kept cluster
// ------------------------------------------------------------------------------------------
static constexpr int FIX_SMALL = 64;                    // buckets with more survivors than this are ordered by a CTA (k_bucket_fixup_big)
static constexpr int BP_PER = 16, BP_TILE = 256 * BP_PER;   // buckets per tile of k_bucket_prefix
static constexpr int BP_TILE_SHIFT = 12;
static_assert((1 << BP_TILE_SHIFT) == BP_TILE, "tile size");
static constexpr uint32_t BP_NONE = 0xffffffffu;        // bpre[] of a bucket that is not flagged

__device__ __forceinline__ uint32_t indel_key32(int32_t c, int32_t raw, int is_ins, const ContigTab& ct, uint32_t& bad) {
    if (c < 0 || c >= ct.n) { bad |= ST_BAD_CHROM; return 0u; }
    const int64_t pos = is_ins ? (raw >> 1) : raw;
    if (raw < 0 || pos > ct.len[c]) { bad |= ST_BAD_POS; return 0u; }   // len < 0: contig not owned by this shard
    return (uint32_t)(ct.off[c] + (uint64_t)pos);
}

__global__ void __launch_bounds__(256) k_indel_hist(const int32_t* __restrict__ chrom, const int32_t* __restrict__ a, int64_t n, int is_ins,
                                                    ContigTab ct, uint32_t* status, uint32_t* __restrict__ bkt) {
    pdl_launch_dependents();   // the next kernel of the chain may become resident now (it waits for this grid to finish)
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((((uintptr_t)chrom) | ((uintptr_t)a)) & 15) == 0;
    const int64_t nv = aligned ? (n >> 2) : 0;
    uint32_t bad = 0;
    for (int64_t v = tid; v < nv; v += stride) {
        const int4 c4 = __ldcs(reinterpret_cast<const int4*>(chrom) + v), a4 = __ldcs(reinterpret_cast<const int4*>(a) + v);
        const uint32_t k0 = indel_key32(c4.x, a4.x, is_ins, ct, bad), k1 = indel_key32(c4.y, a4.y, is_ins, ct, bad);
        const uint32_t k2 = indel_key32(c4.z, a4.z, is_ins, ct, bad), k3 = indel_key32(c4.w, a4.w, is_ins, ct, bad);
        atomicAdd(&bkt[(k0 >> BKT_SHIFT) + BKT_PAD], 1u); atomicAdd(&bkt[(k1 >> BKT_SHIFT) + BKT_PAD], 1u);
        atomicAdd(&bkt[(k2 >> BKT_SHIFT) + BKT_PAD], 1u); atomicAdd(&bkt[(k3 >> BKT_SHIFT) + BKT_PAD], 1u);
    }
    for (int64_t i = nv * 4 + tid; i < n; i += stride) atomicAdd(&bkt[(indel_key32(chrom[i], a[i], is_ins, ct, bad) >> BKT_SHIFT) + BKT_PAD], 1u);
    if (bad) atomicOr(status, bad);
}

// per bucket: bpre[b] = number of survivors in flagged buckets before b INSIDE its tile, or BP_NONE when the +-rb
// bucket neighbourhood holds fewer than `need` signatures; tile_tot[tile] = survivors of the tile.  Buckets with
// more than FIX_SMALL survivors are listed for the CTA-sized in-bucket pass.  Every tile is independent (no
// look-back chain): one streaming read of the histogram, one streaming write of bpre.
struct BigBuckets { uint4* list; uint32_t cap; uint32_t* count; };   // (tile, offset inside the tile, count, -)
// RB in 1..8: neighbourhood radius known at compile time, every thread keeps its 16 buckets + halo (32 counts, eight
// 128-bit loads) in registers; RB == 0: any radius <= BKT_PAD through a shared-memory tile.
template <int RB>
__global__ void __launch_bounds__(256) k_bucket_prefix(const uint32_t* __restrict__ bkt, uint32_t n_buckets, int rb, uint32_t need,
                                                       uint32_t* __restrict__ bpre, uint32_t* __restrict__ tile_tot, BigBuckets BB,
                                                       uint32_t* status, uint32_t* done_ctr, uint32_t* n_out) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    __shared__ uint32_t s_b[RB == 0 ? (BP_TILE + 2 * BKT_PAD + 8) * 17 / 16 + 8 : 1];
    __shared__ uint32_t s_warp[9];
    __shared__ uint32_t s_last;
    auto P = [](int i) { return i + (i >> 4); };   // +1 word per 16: threads stride 17 words -> no bank conflicts
    const uint32_t n_tiles = (n_buckets + BP_TILE - 1) / BP_TILE;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = (int64_t)tile * BP_TILE;
        const int l0 = threadIdx.x * BP_PER;
        uint32_t mask = 0, mine = 0;
        uint32_t own[BP_PER];
        if (RB > 0) {
            // w[k] = count of bucket t0 + l0 - 8 + k; (BKT_PAD - 8) words keep the 16 B alignment
            uint32_t w[32];
            const uint4* src = reinterpret_cast<const uint4*>(bkt + BKT_PAD + t0 + l0 - 8);
#pragma unroll
            for (int k = 0; k < 8; k++) { const uint4 x = __ldg(src + k); w[4 * k] = x.x; w[4 * k + 1] = x.y; w[4 * k + 2] = x.z; w[4 * k + 3] = x.w; }
            uint32_t sum = 0;
#pragma unroll
            for (int k = 8 - RB; k <= 8 + RB; k++) sum += w[k];
#pragma unroll
            for (int j = 0; j < BP_PER; j++) {
                own[j] = w[8 + j];
                if (t0 + l0 + j < n_buckets && sum >= need) { mask |= 1u << j; mine += own[j]; }
                if (j + 1 < BP_PER) sum += w[8 + j + RB + 1] - w[8 + j - RB];
            }
        } else {
            for (int i = threadIdx.x; i < BP_TILE + 2 * rb + 1; i += 256) s_b[P(i)] = bkt[t0 - rb + i + BKT_PAD];
            __syncthreads();
            uint32_t sum = 0;
            for (int k = 0; k <= 2 * rb; k++) sum += s_b[P(l0 + k)];
#pragma unroll
            for (int j = 0; j < BP_PER; j++) {
                own[j] = s_b[P(l0 + j + rb)];
                if (t0 + l0 + j < n_buckets && sum >= need) { mask |= 1u << j; mine += own[j]; }
                sum += s_b[P(l0 + j + 2 * rb + 1)] - s_b[P(l0 + j)];
            }
        }
        uint32_t total;
        uint32_t run = block_excl_scan_256(mine, s_warp, &total);
        if (threadIdx.x == 0) tile_tot[tile] = total;
        uint32_t o[BP_PER];
#pragma unroll
        for (int j = 0; j < BP_PER; j++) {
            o[j] = BP_NONE;
            if (mask >> j & 1u) {
                const uint32_t cnt = own[j];
                o[j] = run;
                if (cnt > FIX_SMALL) {
                    const uint32_t q = atomicAdd(BB.count, 1u);
                    if (q < BB.cap) BB.list[q] = make_uint4(tile, run, cnt, 0u); else atomicOr(status, ST_LIST_OVERFLOW);
                }
                run += cnt;
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(bpre + t0 + l0);   // bpre is sized to whole tiles
#pragma unroll
        for (int j = 0; j < BP_PER; j += 4) dst[j >> 2] = make_uint4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        if (RB == 0) __syncthreads();
    }
    // the CTA that finishes last turns the tile totals into tile bases (a few thousand words) and reports the number of
    // survivors: no separate scan launch
    __threadfence();
    if (threadIdx.x == 0) s_last = atomicAdd(done_ctr, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n_tiles; base += 256) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? __ldcg(&tile_tot[i]) : 0u;
        uint32_t total;
        const uint32_t ex = block_excl_scan_256(v, s_warp, &total);
        if (i < n_tiles) tile_tot[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *n_out = carry;
}

// exclusive scan of arr[0..n) by ONE CTA (n: a few thousand words), *total_out = sum
__global__ void __launch_bounds__(1024) k_scan_small(uint32_t* arr, int64_t n, uint32_t* total_out) {
    __shared__ uint32_t s_w[33];
    __shared__ uint32_t s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = i < n ? arr[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += y; }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_w[lane];
            uint32_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += y; }
            s_w[lane] = wi - w;
            if (lane == 31) s_w[32] = wi;
        }
        __syncthreads();
        if (i < n) arr[i] = s_run + s_w[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_run += s_w[32];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = s_run;
}

// survivors into bucket order: slot = tile base + offset of the bucket inside its tile + (count of the bucket,
// counted down by one returning atomic).  Afterwards every flagged bucket's count is zero again.
// (Tried: counting the slots up in bpre[] itself, so that the atomic lands on the sector the look-up has just brought
//  into L2 and the histogram is not touched again -- 48 MB less DRAM traffic per launch, but 3 % slower: the atomics
//  then share sectors with the 8.4 M look-ups still in flight.)
__global__ void __launch_bounds__(256) k_indel_scatter(const int32_t* __restrict__ chrom, const int32_t* __restrict__ a, int64_t n, int is_ins,
                                                       ContigTab ct, const uint32_t* __restrict__ bpre, const uint32_t* __restrict__ tile_base,
                                                       uint32_t* __restrict__ bkt, uint2* __restrict__ pairs_out) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((((uintptr_t)chrom) | ((uintptr_t)a)) & 15) == 0;
    const int64_t nv = aligned ? (n >> 2) : 0;
    uint32_t dummy = 0;   // validated by k_indel_hist
    auto place = [&](uint32_t key, uint32_t pre, int64_t i) {
        const uint32_t bu = key >> BKT_SHIFT;
        const uint32_t old = atomicSub(&bkt[bu + BKT_PAD], 1u);
        pairs_out[__ldg(&tile_base[bu >> BP_TILE_SHIFT]) + pre + old - 1u] = make_uint2(key, (uint32_t)i);
    };
    for (int64_t v = tid; v < nv; v += stride) {
        const int4 c4 = __ldcs(reinterpret_cast<const int4*>(chrom) + v), a4 = __ldcs(reinterpret_cast<const int4*>(a) + v);
        const uint32_t k0 = indel_key32(c4.x, a4.x, is_ins, ct, dummy), k1 = indel_key32(c4.y, a4.y, is_ins, ct, dummy);
        const uint32_t k2 = indel_key32(c4.z, a4.z, is_ins, ct, dummy), k3 = indel_key32(c4.w, a4.w, is_ins, ct, dummy);
        // the four look-ups are independent: all in flight before the first is used
        const uint32_t p0 = __ldg(&bpre[k0 >> BKT_SHIFT]), p1 = __ldg(&bpre[k1 >> BKT_SHIFT]);
        const uint32_t p2 = __ldg(&bpre[k2 >> BKT_SHIFT]), p3 = __ldg(&bpre[k3 >> BKT_SHIFT]);
        if (p0 != BP_NONE) place(k0, p0, 4 * v);
        if (p1 != BP_NONE) place(k1, p1, 4 * v + 1);
        if (p2 != BP_NONE) place(k2, p2, 4 * v + 2);
        if (p3 != BP_NONE) place(k3, p3, 4 * v + 3);
    }
    for (int64_t i = nv * 4 + tid; i < n; i += stride) {
        const uint32_t key = indel_key32(chrom[i], a[i], is_ins, ct, dummy);
        const uint32_t pre = __ldg(&bpre[key >> BKT_SHIFT]);
        if (pre != BP_NONE) place(key, pre, i);
    }
}

// order inside every bucket: destination = bucket start + number of bucket members with a smaller (key, slot).
// A CTA stages 2048 slots + a halo of FIX_SMALL on both sides in shared memory (coalesced), every thread ranks its
// 8 slots against their buckets there.  Buckets of more than FIX_SMALL survivors: k_bucket_fixup_big.  Rides along:
// the bucket histogram is cleared for the next call (flagged buckets were counted down to zero, the others not).
static constexpr int FX_TILE = 2048;
__global__ void __launch_bounds__(256) k_bucket_fixup(const uint2* __restrict__ pairs, const uint32_t* n_dev, uint32_t* __restrict__ keys_out,
                                                      uint32_t* __restrict__ idx_out, uint32_t* __restrict__ bkt, int64_t n_bkt, BigBuckets BB,
                                                      const uint32_t* __restrict__ tile_base) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    __shared__ uint32_t s_k[FX_TILE + 2 * FIX_SMALL];
    __shared__ uint32_t s_warp[9];
    {
        uint4* b4 = reinterpret_cast<uint4*>(bkt);   // n_bkt is a multiple of 4, cudaMalloc alignment
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < (n_bkt >> 2); i += (int64_t)gridDim.x * 256) b4[i] = z;
    }
    const int64_t n = *n_dev;
    const int64_t n_tiles = (n + FX_TILE - 1) / FX_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * FX_TILE;
        for (int q = threadIdx.x; q < FX_TILE + 2 * FIX_SMALL; q += 256) {
            const int64_t i = base - FIX_SMALL + q;
            s_k[q] = (i >= 0 && i < n) ? pairs[i].x : 0u;
        }
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < FX_TILE / 256; j++) {
            const int q = FIX_SMALL + j * 256 + threadIdx.x;      // position in s_k
            const int64_t i = base + j * 256 + threadIdx.x;
            if (i >= n) continue;
            const uint32_t key = s_k[q], bu = key >> BKT_SHIFT;
            int lo = q, hi = q + 1, seen = 1;
            uint32_t rank = 0;
            const int qmin = (int)(base - FIX_SMALL < 0 ? FIX_SMALL - base : 0);                       // first valid position
            const int64_t last = n - (base - FIX_SMALL);                                               // one past the last valid position
            const int qmax = (int)(last < FX_TILE + 2 * FIX_SMALL ? last : FX_TILE + 2 * FIX_SMALL);
            while (lo > qmin && seen <= FIX_SMALL) {
                const uint32_t k = s_k[lo - 1];
                if ((k >> BKT_SHIFT) != bu) break;
                rank += k <= key ? 1u : 0u;     // earlier slot: smaller (key, slot) iff key <= mine
                lo--; seen++;
            }
            while (hi < qmax && seen <= FIX_SMALL) {
                const uint32_t k = s_k[hi];
                if ((k >> BKT_SHIFT) != bu) break;
                rank += k < key ? 1u : 0u;
                hi++; seen++;
            }
            if (seen > FIX_SMALL) continue;     // a big bucket: ordered by k_bucket_fixup_big
            const int64_t dst = base - FIX_SMALL + lo + rank;
            keys_out[dst] = key;
            idx_out[dst] = pairs[i].y;
        }
        __syncthreads();
    }
    // big buckets (pile-ups, listed by k_bucket_prefix): one CTA per bucket, counting sort on the low BKT_SHIFT bits
    constexpr int NB = 1 << BKT_SHIFT;
    static_assert(NB == 256, "one histogram bin per thread");
    uint32_t* s_cnt = s_k;
    const uint32_t n_list = min(*BB.count, BB.cap);
    for (uint32_t q = blockIdx.x; q < n_list; q += gridDim.x) {
        const uint4 e = BB.list[q];
        const uint32_t first = tile_base[e.x] + e.y, cnt = e.z;
        s_cnt[threadIdx.x] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += 256) atomicAdd(&s_cnt[pairs[first + i].x & (NB - 1)], 1u);
        __syncthreads();
        uint32_t total;
        const uint32_t ex = block_excl_scan_256(s_cnt[threadIdx.x], s_warp, &total);
        s_cnt[threadIdx.x] = ex;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cnt; i += 256) {
            const uint2 pr = pairs[first + i];
            const uint32_t dst = first + atomicAdd(&s_cnt[pr.x & (NB - 1)], 1u);
            keys_out[dst] = pr.x;
            idx_out[dst] = pr.y;
        }
        __syncthreads();
    }
}
// small types: three sort keys per signature (name, second coordinate, primary)
__global__ void k_other_keys(const int32_t* __restrict__ chrom, const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                             const int32_t* __restrict__ rid, const int32_t* __restrict__ c, int64_t n, int svtype, ContigTab ct,
                             uint32_t* __restrict__ k_rid, uint32_t* __restrict__ k_b, uint64_t* __restrict__ k_prim,
                             uint32_t* status) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t bad = 0;
        const int32_t ch = chrom[i], ai = a[i], bi = b[i], ri = rid[i], ci = c ? c[i] : 0;
        if (ch < 0 || ch >= ct.n) bad |= ST_BAD_CHROM;
        else if (ct.len[ch] < 0) bad |= ST_BAD_POS;   // contig not owned by this shard
        if (ai < 0 || bi < 0 || ri < 0 || ci < 0) bad |= ST_NEG_FIELD;
        if (svtype == CSV_TRA && (ci >> 2) >= ct.n) bad |= ST_BAD_CHROM;
        if (svtype == CSV_INV && ci > 1) bad |= ST_NEG_FIELD;
        if (bad) atomicOr(status, bad);
        if (k_rid) { k_rid[i] = (uint32_t)ri; k_b[i] = (uint32_t)bi; }
        // (chr, a) for DUP cuteSV:783; (chr, strand, bp1) for INV cuteSV:792; (chr1, chr2, type, pos1) for TRA :801
        uint64_t hi = (uint64_t)(uint32_t)ch;
        if (svtype == CSV_INV) hi = hi * 2 + (uint32_t)ci;
        if (svtype == CSV_TRA) hi = hi * (uint64_t)(4 * ct.n) + (uint32_t)ci;   // chr2*4+type < 4*n_contigs
        k_prim[i] = (hi << 31) | (uint32_t)ai;                     // a < 2^31
    }
}

// Small types, second half of the tuple order: `perm` is sorted by the primary key only; inside every run of
// equal primary keys order by (b, name, input index) -- the rest of the reference's sort key (cuteSV:783,792,801) --
// by ranking each element against its run.  Runs are short (reads that report the same breakpoint); a run longer
// than RUN_MAX sets ST_BIG_RUN and the host reruns the type with the chained-sorts path.
static constexpr int RUN_MAX = 2048;
__global__ void __launch_bounds__(256) k_run_fixup(const uint64_t* __restrict__ kprim_sorted, const uint32_t* __restrict__ perm, int64_t n,
                                                   const int32_t* __restrict__ b, const int32_t* __restrict__ rid,
                                                   uint32_t* __restrict__ perm_out, uint32_t* status) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = kprim_sorted[i];
        const uint32_t x = perm[i];
        const bool alone = (i == 0 || kprim_sorted[i - 1] != k) && (i + 1 >= n || kprim_sorted[i + 1] != k);
        if (alone) { perm_out[i] = x; continue; }
        const int32_t bx = b[x], rx = rid[x];
        int64_t lo = i, hi = i + 1;
        uint32_t rank = 0;
        int seen = 1;
        auto less = [&](uint32_t y) { const int32_t by = b[y], ry = rid[y]; return by < bx || (by == bx && (ry < rx || (ry == rx && y < x))); };
        while (lo > 0 && kprim_sorted[lo - 1] == k && seen <= RUN_MAX) { rank += less(perm[lo - 1]) ? 1u : 0u; lo--; seen++; }
        while (hi < n && kprim_sorted[hi] == k && seen <= RUN_MAX) { rank += less(perm[hi]) ? 1u : 0u; hi++; seen++; }
        if (seen > RUN_MAX) { atomicOr(status, ST_BIG_RUN); perm_out[i] = x; continue; }
        perm_out[lo + rank] = x;
    }
}

// keys_out[i] = src[perm[i]] (perm == nullptr: identity)
template <typename K>
__global__ void k_gather_keys(const K* __restrict__ src, const uint32_t* __restrict__ perm, int64_t n, K* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = src[perm ? perm[i] : (uint32_t)i];
}

// small types: exact-duplicate removal (cuteSV:958-969) in the final order
struct DedupPred {
    const int32_t *chrom, *a, *b, *rid, *c;
    const uint32_t* perm;
    __device__ __forceinline__ bool operator()(int64_t i) const {
        if (i == 0) return true;
        const uint32_t x = perm[i], y = perm[i - 1];
        return !(chrom[x] == chrom[y] && a[x] == a[y] && b[x] == b[y] && rid[x] == rid[y] && (c ? c[x] == c[y] : true));
    }
};
// U columns = input columns gathered through perm[sel[k]]
__global__ void k_other_gather(const int32_t* __restrict__ chrom, const int32_t* __restrict__ a, const int32_t* __restrict__ b,
                               const int32_t* __restrict__ rid, const int32_t* __restrict__ c, const uint32_t* __restrict__ perm,
                               const uint32_t* __restrict__ sel, const uint32_t* n_sel, int32_t* u_chrom, int32_t* u_a, int32_t* u_b,
                               int32_t* u_rid, int32_t* u_c) {
    const int64_t n = *n_sel;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t x = perm[sel[k]];
        u_chrom[k] = chrom[x]; u_a[k] = a[x]; u_b[k] = b[x]; u_rid[k] = rid[x]; u_c[k] = c ? c[x] : 0;
    }
}

// ------------------------------------------------------------------------------------------
// cluster kernels
// ------------------------------------------------------------------------------------------
// KIND selects the one per-type routine a kernel instantiation contains (0 INS/DEL, 1 DUP, 2 INV, 3 TRA):
// one routine per kernel keeps the hot code inside the instruction cache.
__host__ __device__ constexpr int kind_of(int svtype) { return (svtype == CSV_DEL || svtype == CSV_INS) ? 0 : svtype == CSV_DUP ? 1 : svtype == CSV_INV ? 2 : 3; }
// KIND 4..7: INS / DEL with the type (and "every member is kept") fixed at compile time: DEL, INS, DEL keep-all, INS keep-all
template <int KIND, class Team>
__device__ __forceinline__ void run_cluster(Team tm, const TypeJob& J, int64_t s, int m, int M, char* arena, int64_t* red,
                                            uint32_t kslot, const Emit& E) {
    if (KIND == 4) indel_cluster<Team, CSV_DEL, 0>(tm, J.iv, s, m, M, arena, red, J.cp, J.svtype, kslot, E);
    else if (KIND == 5) indel_cluster<Team, CSV_INS, 0>(tm, J.iv, s, m, M, arena, red, J.cp, J.svtype, kslot, E);
    else if (KIND == 6) indel_cluster<Team, CSV_DEL, 1>(tm, J.iv, s, m, M, arena, red, J.cp, J.svtype, kslot, E);
    else if (KIND == 7) indel_cluster<Team, CSV_INS, 1>(tm, J.iv, s, m, M, arena, red, J.cp, J.svtype, kslot, E);
    else if (KIND == 0) indel_cluster(tm, J.iv, s, m, M, arena, red, J.cp, J.svtype, kslot, E);
    else if (KIND == 1) dup_cluster(tm, J.sv, s, m, M, arena, red, J.cp, kslot, E);
    else if (KIND == 2) inv_cluster(tm, J.sv, s, m, M, arena, red, J.cp, kslot, E);
    else tra_cluster(tm, J.sv, s, m, M, arena, red, J.cp, kslot, E);
}
__device__ __forceinline__ int arena_per(const TypeJob& J) {
    return (J.svtype == CSV_DEL || J.svtype == CSV_INS) ? INDEL_ARENA_PER : OTHER_ARENA_PER;
}
static constexpr int ARENA_PER_MAX = INDEL_ARENA_PER > OTHER_ARENA_PER ? INDEL_ARENA_PER : OTHER_ARENA_PER;

// size of the chain cluster starting at s, counting at most `limit`+1 members (warp-cooperative)
__device__ __forceinline__ int cluster_size_warp(const TypeJob& J, int64_t s, int64_t n, int limit) {
    const int lane = threadIdx.x & 31;
    int m = 1;
    while (m <= limit) {
        const int64_t i = s + m + lane;
        const bool brk = (i >= n) || !job_linked(J, i);
        const uint32_t mask = __ballot_sync(0xffffffffu, brk);
        if (mask) { m += __ffs(mask) - 1; return m; }
        m += 32;
    }
    return m;  // > limit
}

// ------------------------------------------------------------------------------------------
// INS / DEL clusters of at most 32 members, every member kept (remain_reads_ratio >= 1): the whole of
// generate_del_cluster / generate_ins_cluster (resolveINDEL.py:110-219, 319-432) in registers, one member per lane --
// no shared-memory arena, ~3x fewer instructions than the general routine (core.h indel_cluster) and, as a kernel of its
// own (k_cluster_small), a loop that stays inside the instruction cache (the general kernel spends most of its issue
// slots waiting for instructions: 47 KB of hot code, 24 warps per SM at different places in it).
//   exact duplicates / several signatures of one read: every lane looks at its peers (match.any on the read id); the
//   lowest lane of a read carries the read's entry = its longest signature (first of equal length in (pos, len, input)
//   order, resolveINDEL.py:125-131), ordered by the read's first occurrence (the dict order the stable sort by length
//   keeps): sort key (len_best, pos_first, len_first, name).
// 85 % of the kept clusters of 30x ONT have <= 32 members.
// ------------------------------------------------------------------------------------------
template <bool IS_INS>
__device__ __forceinline__ void indel_cluster_small(const IndelView& in, int64_t s, int m, const ClusterParams& P, uint32_t kslot, const Emit& E) {
    const int lane = (int)(threadIdx.x & 31);
    const bool act = lane < m;
    int32_t a = 0, len = 0, rid = -1 - lane, aux = 0;   // idle lanes: distinct negative ids
    uint32_t idx = 0;
    if (act) {
        if (in.rec) {
            const IndelRec r = in.rec[s + lane];
            a = r.a; len = r.b; rid = r.rid; idx = r.idx;
            if (IS_INS) aux = in.recc ? in.recc[s + lane] : 0;
        } else {
            idx = in.sidx[s + lane];
            a = in.a[idx]; len = in.b[idx]; rid = in.rid[idx];
            if (IS_INS) aux = in.c ? in.c[idx] : 0;
        }
    }
    const int32_t pos = IS_INS ? (a >> 1) : a;
    // the read's entry: best = its longest signature, first = its first occurrence in (pos, len, input index) order
    int32_t len_b = len, pos_b = pos, aux_b = aux, pos_f = pos, len_f = len;
    uint32_t idx_b = idx, idx_f = idx;
    bool rep = act;
    const uint32_t peers = __match_any_sync(0xffffffffu, rid);
    if (__any_sync(0xffffffffu, peers != (1u << lane))) {
        uint32_t rem = peers & ~(1u << lane);
        while (__any_sync(0xffffffffu, rem != 0u)) {
            const int j = rem ? __ffs(rem) - 1 : lane;
            rem &= rem - 1u;
            const int32_t pj = __shfl_sync(0xffffffffu, pos, j), lj = __shfl_sync(0xffffffffu, len, j), xj = __shfl_sync(0xffffffffu, aux, j);
            const uint32_t ij = __shfl_sync(0xffffffffu, idx, j);
            if (j != lane) {
                if (pj < pos_f || (pj == pos_f && (lj < len_f || (lj == len_f && ij < idx_f)))) { pos_f = pj; len_f = lj; idx_f = ij; }
                if (lj > len_b || (lj == len_b && (pj < pos_b || (pj == pos_b && ij < idx_b)))) { len_b = lj; pos_b = pj; aux_b = xj; idx_b = ij; }
            }
        }
        rep = act && (int)(__ffs(peers) - 1) == lane;   // one lane per read
    }
    const uint32_t reps = __ballot_sync(0xffffffffu, rep);
    const int u = __popc(reps);
    if (u < P.min_support) {   // len(read_tag) < read_count (:133); (the signature-count test :62 is implied)
        if (lane == 0) E.cnt[kslot] = 0;
        return;
    }
    // bitonic sort of the reads by (len_best, pos_first, len_first, name), carrying the lane that holds the entry
    uint64_t k0 = rep ? (((uint64_t)(uint32_t)len_b << 32) | (uint32_t)pos_f) : ~0ull;
    uint64_t k1 = rep ? (((uint64_t)(uint32_t)len_f << 32) | (uint32_t)rid) : ~0ull;
    int src = lane;
#pragma unroll 1
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
            const uint64_t o0 = __shfl_xor_sync(0xffffffffu, k0, j), o1 = __shfl_xor_sync(0xffffffffu, k1, j);
            const int os = __shfl_xor_sync(0xffffffffu, src, j);
            const bool lower = (lane & j) == 0, up = (lane & k) == 0;
            const bool gt = k0 > o0 || (k0 == o0 && k1 > o1);
            const bool lt = k0 < o0 || (k0 == o0 && k1 < o1);
            if ((lower == up) ? gt : lt) { k0 = o0; k1 = o1; src = os; }
        }
    }
    // lane p < u now holds the read of sorted position p
    const bool act_s = lane < u;
    const int32_t len_s = (int32_t)(k0 >> 32), rid_s = (int32_t)(uint32_t)k1;
    const int32_t pos_s = __shfl_sync(0xffffffffu, pos_b, src);
    const int32_t aux_s = IS_INS ? __shfl_sync(0xffffffffu, aux_b, src) : 0;
    const uint32_t idx_s = IS_INS ? __shfl_sync(0xffffffffu, idx_b, src) : 0u;
    // allele split on the length-sorted reads (:137-162)
    int64_t sum_len = act_s ? (int64_t)len_s : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum_len += __shfl_xor_sync(0xffffffffu, sum_len, o);
    const double thr = P.ratio * ((double)sum_len / (double)u);
    const int32_t len_prev = __shfl_up_sync(0xffffffffu, len_s, 1);
    const bool brk = act_s && lane > 0 && (double)(len_s - len_prev) > thr;
    const uint32_t B = __ballot_sync(0xffffffffu, brk) | 1u;     // bit p: an allele starts at sorted position p
    const int na = __popc(B);
    // alleles in (support, order) order = sorted(allele_collect, key=[support]) stable (:163): lane q < na owns allele q
    int a_st = 0, a_n = 0, a_rank = 0;
    if (lane < na) {
        a_st = (int)__fns(B, 0, lane + 1);
        const int a_en = lane + 1 < na ? (int)__fns(B, 0, lane + 2) : u;
        a_n = a_en - a_st;
    }
    for (int q = 0; q < na; q++) {
        const int nq = __shfl_sync(0xffffffffu, a_n, q);
        if (lane < na && (nq < a_n || (nq == a_n && q < lane))) a_rank++;
    }
    uint32_t n_emit = 0;
    const int32_t chrom = in.chrom[in.rec ? in.rec[s].idx : in.sidx[s]];
    for (int k = 0; k < na; k++) {
        const uint32_t who = __ballot_sync(0xffffffffu, lane < na && a_rank == k);
        const int owner = __ffs(who) - 1;
        const int st = __shfl_sync(0xffffffffu, a_st, owner), n = __shfl_sync(0xffffffffu, a_n, owner);
        if (n < P.min_support_allele) continue;
        const bool mem = lane >= st && lane < st + n;
        int64_t sp = mem ? (int64_t)pos_s : 0, sl = mem ? (int64_t)len_s : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { sp += __shfl_xor_sync(0xffffffffu, sp, o); sl += __shfl_xor_sync(0xffffffffu, sl, o); }
        int32_t pos_out, search, aux_out = 0;
        if (IS_INS) {
            // first member (allele order) whose sequence is long enough (:399-405); signalLen = sl / n (every member kept)
            const int32_t need = (int32_t)((double)sl / (double)n);
            const uint32_t okm = __ballot_sync(0xffffffffu, mem && aux_s >= need);
            if (!okm) continue;   // ideal_ins_seq == '<INS>' -> dropped
            const int pick = __ffs(okm) - 1;
            pos_out = __shfl_sync(0xffffffffu, pos_s, pick);
            aux_out = (int32_t)__shfl_sync(0xffffffffu, idx_s, pick);
            search = pos_out;
        } else {
            // member closest to the mean position, index tie-break: |x - mean| ordered like |n*x - sum| (exact)
            int64_t d = (int64_t)n * pos_s - sp;
            if (d < 0) d = -d;
            uint64_t key = mem ? (((uint64_t)d << 5) | (uint32_t)(lane - st)) : ~0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const uint64_t y = __shfl_xor_sync(0xffffffffu, key, o); if (y < key) key = y; }
            search = __shfl_sync(0xffffffffu, pos_s, st + (int)(key & 31u));   // search_threshold (:177)
            pos_out = (int32_t)((double)sp / (double)n);
        }
        uint32_t slot = 0, noff = 0;
        if (lane == 0) emit_reserve_issue(E, (uint32_t)n, &slot, &noff);   // looked at after the std work
        // CIPOS / CILEN: np.std over the allele (:191-194), lanes 0-7 pos / 8-15 len = numpy's eight strided accumulators
        int32_t cipos, cilen;
        {
            const int which = (lane >> 3) & 1, j8 = lane & 7;
            const double mean = (double)(which ? sl : sp) / (double)n;
            auto sq = [&](int i) {
                const int32_t v0 = __shfl_sync(0xffffffffu, pos_s, (st + i) & 31), v1 = __shfl_sync(0xffffffffu, len_s, (st + i) & 31);
                const double x = (double)(which ? v1 : v0) - mean;
                return x * x;
            };
            double res;
            if (n < 8) {
                res = 0.;
                for (int i = 0; i < n; i++) res += sq(i);
            } else {
                const int n8 = n - (n % 8);
                double r = sq(j8);
                for (int i = 8 + j8; i < n8; i += 8) r += sq(i);
                r = r + __shfl_xor_sync(0xffffffffu, r, 1);
                r = r + __shfl_xor_sync(0xffffffffu, r, 2);
                r = r + __shfl_xor_sync(0xffffffffu, r, 4);
                res = r;
                for (int i = n8; i < n; i++) res += sq(i);
            }
            res = res / (double)n;
            const int32_t ci = cal_cipos(sqrt(res), n, E.pow_half);
            cipos = __shfl_sync(0xffffffffu, ci, 0);
            cilen = __shfl_sync(0xffffffffu, ci, 8);
        }
        int ok = 0;
        if (lane == 0) {
            note_support(E, (uint32_t)n);
            ok = emit_reserve_check(E, (uint32_t)n, slot, noff) ? 1 : 0;
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        noff = __shfl_sync(0xffffffffu, noff, 0);
        if (ok) {
            if (mem) E.names[noff + (uint32_t)(lane - st)] = rid_s;
            if (lane == 0) {
                const double signalLen = (double)sl / (double)n;
                csv_cand c;
                c.svtype = IS_INS ? CSV_INS : CSV_DEL; c.chrom = chrom; c.pos = pos_out;
                c.len = IS_INS ? (int32_t)signalLen : (int32_t)(-signalLen);
                c.support = n; c.cipos = cipos; c.cilen = cilen; c.search_pos = search; c.pos2 = 0; c.aux = aux_out;
                c.names_off = (int32_t)noff; c.names_cnt = n; c.cluster = (int32_t)kslot; c.flags = 0;
                c.reserved[0] = (int32_t)n_emit; c.reserved[1] = 0;
                E.cand[slot] = c;
            }
        }
        n_emit++;
    }
    if (lane == 0) E.cnt[kslot] = n_emit;
}

// one warp per kept cluster of <= 32 members.  Two ways in: (a) k_select_heads already sorted the kept clusters into
// `small_list` (ordinal, members) and `rest_list` while it gathered their records, so this kernel and the general kernel run
// side by side on two streams; (b) no lists (gather mode): every kept cluster is sized here and the larger ones are listed
// for the general kernel, which then runs after this one.
template <bool IS_INS>
__global__ void __launch_bounds__(256) k_cluster_small(TypeJob J, Emit E, Counters* ctr, uint32_t* work, uint32_t* n_rest) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    const int lane = threadIdx.x & 31;
    const int64_t n = job_n(J);
    const uint32_t n_todo = J.small_list ? *J.n_small : ctr->n_kept[J.svtype];
    uint32_t k_next = 0, n_done = 0, n_mem = 0;
    if (lane == 0) k_next = atomicAdd(work, 1u);
    while (true) {
        const uint32_t q = __shfl_sync(0xffffffffu, k_next, 0);
        if (q >= n_todo) break;
        if (lane == 0) k_next = atomicAdd(work, 1u);
        uint32_t k;
        int m;
        if (J.small_list) { const uint2 e = J.small_list[q]; k = e.x; m = (int)e.y; }
        else {
            k = q;
            m = cluster_size_warp(J, J.kept_start[k], n, 32);
            if (m > 32) {
                if (lane == 0) J.rest_list[atomicAdd(n_rest, 1u)] = k;   // (capacity = kept capacity of the type)
                continue;
            }
        }
        indel_cluster_small<IS_INS>(J.iv, J.kept_start[k], m, J.cp, J.kslot_base + k, E);
        n_done++; n_mem += (uint32_t)m;
    }
    if (lane == 0) {
        if (n_done) atomicAdd(&ctr->pad[0], n_done);
        if (n_mem) atomicAdd(&ctr->n_members[J.svtype], n_mem);
    }
}

// one warp per kept cluster; clusters larger than WARP_M are deferred to the CTA kernel
template <int KIND>
__global__ void __launch_bounds__(CL_THREADS) k_cluster_warp(TypeJob J, Emit E, Counters* ctr, uint32_t* work) {
    pdl_launch_dependents();   // the next kernel of the chain may become resident now (it waits for this grid to finish)
    extern __shared__ __align__(16) char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int WARPS = CL_THREADS / 32;
    char* arena = smem + (size_t)warp * (WARP_M * ARENA_PER_MAX);
    int64_t* red = (int64_t*)(smem + (size_t)WARPS * (WARP_M * ARENA_PER_MAX)) + warp * 40;
    const int64_t n = job_n(J);
    const uint32_t n_kept = J.rest_list ? *J.n_rest : ctr->n_kept[J.svtype];
    CudaTeam<32> tm;
    // dynamic hand-out (one atomic per cluster): cluster costs vary, a static stride leaves a long tail
    // (the ticket of the NEXT cluster is requested before the current one is processed, so the atomic's round
    //  trip overlaps the work)
    uint32_t k_next = 0, n_mem = 0;
    if (lane == 0) k_next = atomicAdd(work, 1u);
    while (true) {
        const uint32_t q = __shfl_sync(0xffffffffu, k_next, 0);
        if (q >= n_kept) break;
        if (lane == 0) k_next = atomicAdd(work, 1u);
        const uint32_t k = J.rest_list ? J.rest_list[q] : q;   // after k_cluster_small: only the clusters it left
        const int64_t s = J.kept_start[k];
        const int m = cluster_size_warp(J, s, n, WARP_M);
        if (m > WARP_M) {
            if (lane == 0) {
                const uint32_t o = atomicAdd(&ctr->n_big[J.svtype], 1u);
                if (o < J.big_cap) J.big_list[o] = k; else atomicOr(&ctr->status, ST_LIST_OVERFLOW);
            }
            continue;
        }
        n_mem += (uint32_t)m;
        run_cluster<KIND>(tm, J, s, m, pow2ceil(m), arena, red, J.kslot_base + k, E);
        __syncwarp();
    }
    if (lane == 0 && n_mem) atomicAdd(&ctr->n_members[J.svtype], n_mem);
}

// size of the chain cluster starting at s (CTA-cooperative, exact)
__device__ __forceinline__ int64_t cluster_size_block(const TypeJob& J, int64_t s, int64_t n, int64_t* red) {
    CudaTeam<CL_THREADS> tm;
    int64_t done = 1;
    while (true) {
        const int64_t i = s + done + threadIdx.x;
        const int64_t cand = ((i >= n) || !job_linked(J, i)) ? (done + threadIdx.x) : INT64_MAX;
        const int64_t first = team_min(tm, cand, red);
        if (first != INT64_MAX) return first;
        done += CL_THREADS;
    }
}

// one CTA per deferred cluster; clusters beyond the shared-memory arena (> BLOCK_M members) use global scratch
template <int KIND>
__global__ void __launch_bounds__(CL_THREADS) k_cluster_block(TypeJob J, Emit E, Counters* ctr) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    extern __shared__ __align__(16) char smem[];
    __shared__ int64_t red[CL_THREADS + 8];
    const int64_t n = job_n(J);
    const uint32_t n_list = min(ctr->n_big[J.svtype], J.big_cap);
    CudaTeam<CL_THREADS> tm;
    for (uint32_t q = blockIdx.x; q < n_list; q += gridDim.x) {
        const uint32_t k = J.big_list[q];
        const int64_t s = J.kept_start[k];
        const int64_t m = cluster_size_block(J, s, n, red);
        const bool giant = m > BLOCK_M;
        const int M = pow2ceil((int)m);
        if (threadIdx.x == 0) {
            atomicAdd(&ctr->n_members[J.svtype], (uint32_t)m);
            if (giant) atomicAdd(&ctr->n_giant[J.svtype], 1u);
        }
        // global scratch: clusters are disjoint ranges of the sorted order and M <= 2m
        char* arena = giant ? (J.giant_arena + (size_t)(2 * s) * ARENA_PER_MAX) : smem;
        run_cluster<KIND>(tm, J, s, (int)m, M, arena, red, J.kslot_base + k, E);
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// order: final position = (exclusive scan of per-cluster counts)[kslot] + emission rank
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// genotype
// ------------------------------------------------------------------------------------------
// One genotype window, ready to be tested against a read: linear bounds of this window and (second window of DUP / INV
// only) of the candidate's first window, the slice of supporting read ids, the candidate index.  32 B = one sector.
struct alignas(16) WinRec { uint32_t S, E, S0, E0; int32_t names_off, names_cnt; uint32_t cand, second; };
struct GenoJob {
    csv_cand* cand;           // final order
    csv_geno* geno;
    const int32_t* names;
    const Counters* ctr;
    uint32_t cap_cand;
    ContigTab ct;
    GtParams gp;
    int shift;
    uint32_t n_bins;
    uint32_t* bin_start;      // n_bins + 1 (counts, then exclusive offsets)
    uint32_t* bin_fill;       // n_bins
    uint32_t* bin_bits;       // n_bins/32 + 1: bin holds at least one window (small, stays in L1)
    uint32_t* win_list;       // cand*2 + which, grouped by bin
    uint32_t win_cap;
    struct WinRec* win_rec;   // lin32 only: the same slots as 32 B records (window bounds in linear coordinates, names slice, cand)
    int lin32;                // the linear coordinate fits 32 bits: (read, window) pairs carry the read's coordinates
    uint32_t* dr;             // per candidate
    uint8_t* has_rows;        // per contig: reads table has rows (call_gt's `chr not in sigs_index["reads"]`)
    const csv_geno* gl_table;
    int genotype;
};

__device__ __forceinline__ uint32_t window_bin(const GenoJob& G, const csv_cand& c, int which) {
    int64_t s, e;
    window_of(c, which, G.gp, &s, &e);
    return (uint32_t)((G.ct.off[c.chrom] + (uint64_t)s) >> G.shift);
}

// candidates into the reference's emission order; when genotyping, the same pass counts the genotype windows per bin
__global__ void k_permute(const csv_cand* __restrict__ tmp, const uint32_t* __restrict__ base, Counters* ctr, uint32_t cap,
                          csv_cand* __restrict__ out, GenoJob G, const unsigned long long* cursor) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    const unsigned long long cur = *cursor;
    if (blockIdx.x == 0 && threadIdx.x == 0) { ctr->n_cand = (uint32_t)(cur >> 32); ctr->n_names = (uint32_t)cur; }   // for the kernels after this one
    const uint32_t n = min((uint32_t)(cur >> 32), cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        csv_cand c = tmp[i];
        const uint32_t dst = base[c.cluster] + (uint32_t)c.reserved[0];
        c.reserved[0] = 0;
        if (dst >= cap) continue;
        out[dst] = c;
        if (G.genotype) {
            const int nw = n_windows_of(c);
            for (int w = 0; w < nw; w++) {
                uint32_t b = window_bin(G, c, w);
                if (b >= G.n_bins) b = G.n_bins - 1;
                atomicAdd(&G.bin_start[b], 1u);
                atomicOr(&G.bin_bits[b >> 5], 1u << (b & 31));
            }
            G.dr[dst] = 0;
        }
    }
}

// pass 1: scatter the windows into their bins (bin_start already scanned; pass 0 = the counting is part of k_permute)
template <int PASS>
__global__ void k_windows(GenoJob G) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    const uint32_t n = min(G.ctr->n_cand, G.cap_cand);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const csv_cand c = G.cand[i];
        const int nw = n_windows_of(c);
        for (int w = 0; w < nw; w++) {
            uint32_t b = window_bin(G, c, w);
            if (b >= G.n_bins) b = G.n_bins - 1;
            if (PASS == 0) { atomicAdd(&G.bin_start[b], 1u); atomicOr(&G.bin_bits[b >> 5], 1u << (b & 31)); }
            else {
                const uint32_t o = G.bin_start[b] + atomicAdd(&G.bin_fill[b], 1u);
                if (o < G.win_cap) {
                    G.win_list[o] = i * 2u + (uint32_t)w;
                    if (G.lin32) {
                        const uint64_t coff = G.ct.off[c.chrom];
                        int64_t ws, we, s0 = 0, e0 = 0;
                        window_of(c, w, G.gp, &ws, &we);
                        if (w) window_of(c, 0, G.gp, &s0, &e0);
                        WinRec r;
                        r.S = (uint32_t)(coff + (uint64_t)ws); r.E = (uint32_t)(coff + (uint64_t)we);
                        r.S0 = (uint32_t)(coff + (uint64_t)s0); r.E0 = (uint32_t)(coff + (uint64_t)e0);
                        r.names_off = c.names_off; r.names_cnt = c.names_cnt; r.cand = i; r.second = (uint32_t)w;
                        *reinterpret_cast<uint4*>(&G.win_rec[o]) = *reinterpret_cast<const uint4*>(&r);
                        *(reinterpret_cast<uint4*>(&G.win_rec[o]) + 1) = *(reinterpret_cast<const uint4*>(&r) + 1);
                    }
                }
            }
        }
        if (PASS == 0) G.dr[i] = 0;
    }
}

// One (read, window) test of assign_gt / overlap_cover: a primary read covers window [s,e] iff
// start <= s and end >= e (cuteSV_genotype.py:100-138); DR counts covering reads that are not
// supporting reads (:161-173).  RS/RE are linear coordinates.
__device__ __forceinline__ void test_pair(const GenoJob& G, uint64_t RS, uint64_t RE, int32_t rid, uint32_t w) {
    const uint32_t ent = G.win_list[w];
    const csv_cand c = G.cand[ent >> 1];
    const uint64_t coff = G.ct.off[c.chrom];
    int64_t s, e;
    window_of(c, (int)(ent & 1u), G.gp, &s, &e);
    if (!(RS <= coff + (uint64_t)s && RE >= coff + (uint64_t)e)) return;
    if (ent & 1u) {  // union of the two breakpoint covers (resolveDUP.py:155-157): count once
        int64_t s0, e0;
        window_of(c, 0, G.gp, &s0, &e0);
        if (RS <= coff + (uint64_t)s0 && RE >= coff + (uint64_t)e0) return;
    }
    // supporting reads are not counted; four independent loads per step instead of a load-compare-branch chain
    const int32_t* nm = G.names + c.names_off;
    const int n = c.names_cnt;
    bool found = false;
    int k = 0;
    for (; k + 4 <= n && !found; k += 4)
        found = (__ldg(nm + k) == rid) | (__ldg(nm + k + 1) == rid) | (__ldg(nm + k + 2) == rid) | (__ldg(nm + k + 3) == rid);
    for (; k < n && !found; k++) found = __ldg(nm + k) == rid;
    if (!found) atomicAdd(&G.dr[ent >> 1], 1u);
}

// the same test on a WinRec (32-bit linear coordinates): one 32 B gather instead of the candidate record, its contig
// offset and the window arithmetic
__device__ __forceinline__ void test_pair32(const GenoJob& G, uint32_t RS, uint32_t RE, int32_t rid, uint32_t w) {
    const uint4 lo = __ldg(reinterpret_cast<const uint4*>(&G.win_rec[w])), hi = __ldg(reinterpret_cast<const uint4*>(&G.win_rec[w]) + 1);
    if (!(RS <= lo.x && RE >= lo.y)) return;
    if (hi.w && RS <= lo.z && RE >= lo.w) return;   // union of the two breakpoint covers (resolveDUP.py:155-157): count once
    const int32_t* nm = G.names + (int32_t)hi.x;
    const int n = (int32_t)hi.y;
    bool found = false;
    int k = 0;
    for (; k + 4 <= n && !found; k += 4)
        found = (__ldg(nm + k) == rid) | (__ldg(nm + k + 1) == rid) | (__ldg(nm + k + 2) == rid) | (__ldg(nm + k + 3) == rid);
    for (; k < n && !found; k++) found = __ldg(nm + k) == rid;
    if (!found) atomicAdd(&G.dr[hi.z], 1u);
}

// ONE streaming pass over the reads table (replaces overlap_cover's event sort + sweep,
// cuteSV_genotype.py:95-159).  A read can only cover windows whose start lies in a bin it
// overlaps; most reads overlap no occupied bin (bit test on an L1-resident map).  The few
// (read, window) pairs that remain are compacted (warp-aggregated append) and tested by a second,
// dense kernel so that the streaming pass keeps all 32 lanes busy.
// LIN32 (linear coordinate < 2^32): a pair is 16 B (read start, read end, read id, window slot), so the dense kernel
// needs no gather into the reads table; otherwise 8 B (read row, window slot).
struct PairBuf { uint2* pairs; uint4* pairs4; uint32_t cap; uint32_t* count; };

template <bool LIN32>
__global__ void __launch_bounds__(256) k_reads_pass(GenoJob G, PairBuf PB, const int32_t* __restrict__ r_chrom,
                                                    const int32_t* __restrict__ r_start, const int32_t* __restrict__ r_end,
                                                    const int32_t* __restrict__ r_id, const uint8_t* __restrict__ r_prim,
                                                    int64_t n_reads, uint32_t* status) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    constexpr int ITEMS = 4;
    constexpr int TAB = 1024;            // contigs whose (offset, validity) live in shared memory
    __shared__ uint32_t s_warp[10];
    __shared__ uint64_t s_off[TAB];      // linear offset, ~0 for a contig outside the shard
    __shared__ uint8_t s_seen[TAB];
    const int n_tab = G.ct.n < TAB ? G.ct.n : TAB;
    for (int i = threadIdx.x; i < n_tab; i += 256) { s_off[i] = G.ct.len[i] < 0 ? ~0ull : G.ct.off[i]; s_seen[i] = 0; }
    __syncthreads();
    const int64_t n_tiles = (n_reads + 256 * ITEMS - 1) / (256 * ITEMS);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * 256 * ITEMS;
        int32_t ch[ITEMS], st[ITEMS], en[ITEMS];
        uint8_t pr[ITEMS];
        // row of item j: four consecutive rows per thread (one 128-bit load per column) when the columns allow it
        const int64_t r0 = base + (int64_t)threadIdx.x * ITEMS;
        const bool vec = ITEMS == 4 && r0 + ITEMS <= n_reads &&
                         ((((uintptr_t)r_chrom) | ((uintptr_t)r_start) | ((uintptr_t)r_end)) & 15) == 0 && (((uintptr_t)r_prim) & 3) == 0;
        const bool blocked = __syncthreads_and(vec || r0 >= n_reads) != 0;   // one mapping per tile (uniform: the pair slots are reserved per CTA)
#define RP_ROW(j) (blocked ? r0 + (j) : base + (int64_t)(j) * 256 + threadIdx.x)
        if (blocked && vec) {
            const int4 c4 = __ldcs(reinterpret_cast<const int4*>(r_chrom + r0)), s4 = __ldcs(reinterpret_cast<const int4*>(r_start + r0));
            const int4 e4 = __ldcs(reinterpret_cast<const int4*>(r_end + r0));
            const uint32_t p4 = __ldcs(reinterpret_cast<const uint32_t*>(r_prim + r0));
            ch[0] = c4.x; ch[1] = c4.y; ch[2] = c4.z; ch[3] = c4.w;
            st[0] = s4.x; st[1] = s4.y; st[2] = s4.z; st[3] = s4.w;
            en[0] = e4.x; en[1] = e4.y; en[2] = e4.z; en[3] = e4.w;
            pr[0] = (uint8_t)p4; pr[1] = (uint8_t)(p4 >> 8); pr[2] = (uint8_t)(p4 >> 16); pr[3] = (uint8_t)(p4 >> 24);
        } else {
#pragma unroll
            for (int j = 0; j < ITEMS; j++) {
                const int64_t r = RP_ROW(j);
                const bool in = r < n_reads;
                ch[j] = in ? __ldcs(r_chrom + r) : -1; st[j] = in ? __ldcs(r_start + r) : 0; en[j] = in ? __ldcs(r_end + r) : 0;
                pr[j] = in ? __ldcs(r_prim + r) : 0;
            }
        }
        uint32_t w[ITEMS], cnt[ITEMS], total = 0;
        uint64_t lin[ITEMS];   // RS of the read
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            w[j] = 0; cnt[j] = 0; lin[j] = 0;
            const int64_t r = RP_ROW(j);
            if (r >= n_reads) continue;
            if (ch[j] < 0 || ch[j] >= G.ct.n) { atomicOr(status, ST_BAD_CHROM); continue; }
            uint64_t off;
            if (ch[j] < TAB) {
                off = s_off[ch[j]];
                // deliberate benign race (racecheck reports it, profiles/r02_sanitizer.txt): a byte that only goes 0 -> 1, every writer stores
                // the same value, read after the final __syncthreads().  The race-free variants (shared atomics) cost 10 registers = one CTA
                // per SM = 16-22 us in this latency-bound pass.
                if (!s_seen[ch[j]]) s_seen[ch[j]] = 1;
            } else {
                off = G.ct.len[ch[j]] < 0 ? ~0ull : G.ct.off[ch[j]];
                if (!G.has_rows[ch[j]]) G.has_rows[ch[j]] = 1;
            }
            if (off == ~0ull) { atomicOr(status, ST_BAD_CHROM); continue; }   // outside the shard
            if (!pr[j]) continue;
            const uint64_t RS = off + (uint64_t)(uint32_t)st[j], RE = off + (uint64_t)(uint32_t)en[j];
            lin[j] = RS;
            const uint32_t b0 = (uint32_t)(RS >> G.shift);
            uint32_t b1 = (uint32_t)(RE >> G.shift);
            if (b1 >= G.n_bins) b1 = G.n_bins - 1;
            if (b0 > b1) continue;
            // any occupied bin in [b0, b1]?  One or two words of the bit map for a typical read.
            const uint32_t w0 = b0 >> 5, w1 = b1 >> 5;
            uint32_t m = __ldg(&G.bin_bits[w0]) & (0xffffffffu << (b0 & 31));
            if (w1 == w0) m &= 0xffffffffu >> (31 - (b1 & 31));
            else {
                for (uint32_t wi = w0 + 1; wi < w1 && !m; wi++) m = __ldg(&G.bin_bits[wi]);
                if (!m) m = __ldg(&G.bin_bits[w1]) & (0xffffffffu >> (31 - (b1 & 31)));
            }
            if (m) { w[j] = G.bin_start[b0]; cnt[j] = G.bin_start[b1 + 1] - w[j]; total += cnt[j]; }
        }
        uint32_t o = block_reserve_256(total, PB.count, s_warp);
#pragma unroll
        for (int j = 0; j < ITEMS; j++) {
            if (!cnt[j]) continue;
            const int64_t r = RP_ROW(j);
            const int32_t rid = r_id[r];
            const uint64_t RS = lin[j], RE = lin[j] + (uint64_t)(uint32_t)(en[j] - st[j]);
            // slots below the capacity go to the pair buffer (every reserved slot < cap MUST be written:
            // k_pairs_test consumes [0, min(count, cap))); the rest is tested inline (correct, just slower)
            uint32_t k = 0;
            for (; k < cnt[j] && (uint64_t)o + k < PB.cap; k++) {
                if (LIN32) PB.pairs4[o + k] = make_uint4((uint32_t)RS, (uint32_t)RE, (uint32_t)rid, w[j] + k);
                else PB.pairs[o + k] = make_uint2((uint32_t)r, w[j] + k);
            }
            for (; k < cnt[j]; k++) {
                if (LIN32) test_pair32(G, (uint32_t)RS, (uint32_t)RE, rid, w[j] + k);
                else test_pair(G, RS, RE, rid, w[j] + k);
            }
            o += cnt[j];
        }
#undef RP_ROW
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_tab; i += 256)
        if (s_seen[i] && !G.has_rows[i]) G.has_rows[i] = 1;
}

template <bool LIN32>
__global__ void __launch_bounds__(256) k_pairs_test(GenoJob G, PairBuf PB, const int32_t* __restrict__ r_chrom,
                                                    const int32_t* __restrict__ r_start, const int32_t* __restrict__ r_end,
                                                    const int32_t* __restrict__ r_id) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    const uint32_t n = min(*PB.count, PB.cap);
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        if (LIN32) {
            const uint4 pr = PB.pairs4[p];
            test_pair32(G, pr.x, pr.y, (int32_t)pr.z, pr.w);
        } else {
            const uint2 pr = PB.pairs[p];
            const uint64_t off = G.ct.off[r_chrom[pr.x]];
            test_pair(G, off + (uint64_t)(uint32_t)r_start[pr.x], off + (uint64_t)(uint32_t)r_end[pr.x], r_id[pr.x], pr.y);
        }
    }
}

__global__ void k_finalize(GenoJob G) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    const uint32_t n = min(G.ctr->n_cand, G.cap_cand);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        csv_cand c = G.cand[i];
        csv_geno g;
        g.dr = -1; g.dv = c.names_cnt; g.gt = -1; g.pl[0] = g.pl[1] = g.pl[2] = 0; g.gq = 0; g.status = 1; g.qual = 0.0;
        if (G.genotype && n_windows_of(c) > 0) {
            if (!G.has_rows[c.chrom]) {
                c.flags |= CSV_F_NO_READS;
                G.cand[i].flags = c.flags;
            } else {
                const int32_t dr = (int32_t)G.dr[i];
                g = G.gl_table[gl_index(dr, c.names_cnt)];  // cal_GL(DR, DV) (cuteSV_genotype.py:171)
                g.dr = dr; g.dv = c.names_cnt;
            }
        }
        G.geno[i] = g;
    }
}

// ---- TRA genotyping from the packed all-alignments table ----
__global__ void k_aln_index(const int32_t* __restrict__ chrom, const int32_t* __restrict__ start, const int32_t* __restrict__ end, int64_t n,
                            int32_t n_contigs, uint32_t* off, int32_t* max_span, uint32_t* status) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t c = chrom[i];
        if (c < 0 || c >= n_contigs) { atomicOr(status, ST_BAD_CHROM); continue; }
        if (i == 0 || chrom[i - 1] != c) off[c] = (uint32_t)i;
        if (i > 0 && (chrom[i - 1] > c || (chrom[i - 1] == c && start[i - 1] > start[i]))) atomicOr(status, ST_UNSORTED);
        atomicMax(&max_span[c], end[i] - start[i]);
    }
}
__global__ void k_aln_fill(uint32_t* off, int32_t n_contigs, uint32_t n) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        off[n_contigs] = n;
        for (int c = n_contigs - 1; c >= 0; c--)
            if (off[c] == 0xffffffffu) off[c] = off[c + 1];
    }
}
// Warp-cooperative count_coverage (core.h tra_count_coverage is the scalar statement the emulator runs):
// 32 consecutive BAM-order records per step; the running counters become ballot prefix counts and the
// record that triggers an early return is the lowest lane whose inclusive counts satisfy a return test.
__device__ __forceinline__ int tra_count_coverage_warp(const AlnView& A, int32_t chr, int64_t s, int64_t e, const int32_t* sup, int n_sup,
                                                       int32_t up_bound, int32_t itround, int32_t* nset, int32_t* dr, int64_t xs, int64_t xe) {
    const int lane = threadIdx.x & 31;
    const uint32_t below = (1u << lane) - 1u, upto = below | (1u << lane);
    int64_t iteration = 0, primary = 0;
    const uint32_t lo0 = A.off[chr], hi0 = A.off[chr + 1];
    uint32_t lo = lo0, hi = hi0;
    const int64_t min_start = s - (int64_t)A.max_span[chr];
    while (lo < hi) { uint32_t mid = lo + (hi - lo) / 2; if ((int64_t)A.start[mid] < min_start) lo = mid + 1; else hi = mid; }
    for (uint32_t base = lo; base < hi0; base += 32) {
        const uint32_t i = base + lane;
        const bool valid = i < hi0;
        const int64_t st = valid ? (int64_t)A.start[i] : 0, en = valid ? (int64_t)A.end[i] : 0;
        const bool before_end = valid && st < e;                  // loop condition (starts ascend: a prefix of the lanes)
        const bool fetched = before_end && en > s;
        const bool prim = fetched && A.prim[i] != 0;
        const bool spanning = prim && st < s && en > e;
        const bool fresh = spanning && !(xs <= xe && st < xs && en > xe);
        const bool is_ref = fresh && !sorted_contains(sup, n_sup, A.rid[i]);
        const uint32_t m_f = __ballot_sync(0xffffffffu, fetched), m_p = __ballot_sync(0xffffffffu, prim);
        const uint32_t m_n = __ballot_sync(0xffffffffu, fresh), m_r = __ballot_sync(0xffffffffu, is_ref);
        const int64_t it_incl = iteration + __popc(m_f & upto), pr_incl = primary + __popc(m_p & upto);
        const int32_t ns_incl = *nset + __popc(m_n & upto), dr_incl = *dr + __popc(m_r & upto);
        const bool ret_up = spanning && ns_incl >= up_bound;       // return 1 inside the spanning block
        const bool ret_round = prim && it_incl >= itround;         // the itround test is only reached by primary records
        const uint32_t m_ret = __ballot_sync(0xffffffffu, ret_up || ret_round);
        if (m_ret) {
            const int src = __ffs(m_ret) - 1;
            int code = ret_up ? 1 : (((double)pr_incl / (double)it_incl) <= 0.2 ? 1 : -1);
            code = __shfl_sync(0xffffffffu, code, src);
            *nset = __shfl_sync(0xffffffffu, ns_incl, src);
            *dr = __shfl_sync(0xffffffffu, dr_incl, src);
            return code;
        }
        iteration += __popc(m_f); primary += __popc(m_p);
        *nset += __popc(m_n); *dr += __popc(m_r);
        if (__ballot_sync(0xffffffffu, before_end) != 0xffffffffu) break;   // a record with start >= e (or the contig end) was seen
    }
    return 0;
}
// one warp per TRA candidate; the candidates are ordered by SV type, TRA last, so warp w takes candidate n-1-w
__global__ void __launch_bounds__(128) k_tra_genotype(GenoJob G, AlnView A, int32_t bias, int32_t gt_round) {
    const uint32_t n = min(G.ctr->n_cand, G.cap_cand);
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint32_t w = warp; w < n; w += n_warps) {
        const uint32_t i = n - 1 - w;
        const csv_cand c = G.cand[i];
        if (c.svtype != CSV_TRA) break;   // uniform across the warp
        const int32_t* sup = G.names + c.names_off;
        const int32_t chr1 = c.chrom, chr2 = c.aux >> 2, n_sup = c.names_cnt;
        const int32_t up = threshold_ref_count(n_sup);
        int32_t nset = 0, dr = 0;
        int64_t s = (int64_t)c.pos - bias; if (s < 0) s = 0;
        int64_t e = (int64_t)c.pos + bias; if (e > A.contig_len[chr1]) e = A.contig_len[chr1];
        const int st = tra_count_coverage_warp(A, chr1, s, e, sup, n_sup, up, gt_round, &nset, &dr, 1, 0);
        const int64_t s1 = s, e1 = e;
        csv_geno g;
        if (st == -1) {  // DR '.', GT './.' (resolveTRA.py:277-282)
            g.dr = -1; g.dv = n_sup; g.gt = -1; g.pl[0] = g.pl[1] = g.pl[2] = 0; g.gq = 0; g.status = 2; g.qual = 0.0;
        } else {
            if (st == 0) {
                s = (int64_t)c.pos2 - bias; if (s < 0) s = 0;
                e = (int64_t)c.pos2 + bias; if (e > A.contig_len[chr2]) e = A.contig_len[chr2];
                if (chr2 == chr1) tra_count_coverage_warp(A, chr2, s, e, sup, n_sup, up, gt_round, &nset, &dr, s1, e1);
                else tra_count_coverage_warp(A, chr2, s, e, sup, n_sup, up, gt_round, &nset, &dr, 1, 0);
            }
            g = G.gl_table[gl_index(dr, n_sup)];
            g.dr = dr; g.dv = n_sup;
        }
        if (lane == 0) {
            G.geno[i] = g;
            G.cand[i].flags = c.flags & ~CSV_F_GT_HOST;
        }
    }
}

// cal_GL for arbitrary (c0, c1) pairs: special cases + rescale on the device, libm part from the table
__global__ void k_cal_gl(const int32_t* c0, const int32_t* c1, int64_t n, const csv_geno* gl_table, csv_geno* out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        csv_geno g = gl_table[gl_index(c0[i], c1[i])];
        g.dr = c0[i]; g.dv = c1[i];
        out[i] = g;
    }
}

}  // namespace csv
