// gather_api.inl -- multi-GPU: one csv_ctx per GPU (one process per GPU), contigs sharded over the ranks
// (csv_set_shard), no data-path collective, ONE NCCL all-gather of the final records (the unit of
// independence is (svtype, contig): cuteSV:1116-1189).  Included by cutesv_b200.cu.
//
// NCCL is resolved with dlopen at first use, so the library still loads on a box without NCCL (the ABI test) and a
// process that already carries a libnccl (torch) shares that copy.

struct NcclApi {
    void* h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    bool tried = false;
};
static NcclApi g_nccl;
static int nccl_load() {
    NcclApi& a = g_nccl;
    if (a.h) return CSV_OK;
    if (a.tried) return set_err(CSV_E_STATE, "NCCL is not available (libnccl.so.2 could not be loaded)");
    a.tried = true;
    const char* names[] = {getenv("CUTESV_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        if (!nm || !*nm) continue;
        a.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (a.h) break;
    }
    if (!a.h) return set_err(CSV_E_STATE, "NCCL is not available: %s", dlerror());
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(a.h, "ncclAllGather");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.h, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.AllReduce || !a.GetErrorString) {
        a.h = nullptr;
        return set_err(CSV_E_STATE, "NCCL is not available: symbols missing in libnccl");
    }
    return CSV_OK;
}
#define NC(call)                                                                                                   \
    do {                                                                                                           \
        ncclResult_t r__ = (call);                                                                                 \
        if (r__ != ncclSuccess) return set_err(CSV_E_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r__));    \
    } while (0)

static void p2p_release(csv_ctx* c);
static void comm_destroy(csv_ctx* c) {
    p2p_release(c);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    c->comm = nullptr;
}

extern "C" int csv_comm_unique_id(void* id, size_t bytes) {
    if (!id || bytes < sizeof(ncclUniqueId)) return set_err(CSV_E_INVALID, "unique id buffer must hold %zu bytes", sizeof(ncclUniqueId));
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId u;
    NC(g_nccl.GetUniqueId(&u));
    memset(id, 0, bytes);
    memcpy(id, &u, sizeof(u));
    return CSV_OK;
}

extern "C" int csv_comm_init(csv_ctx* c, const void* id, int rank, int world) {
    if (!c || !id || world < 1 || rank < 0 || rank >= world) return set_err(CSV_E_INVALID, "bad argument");
    int rc = nccl_load();
    if (rc) return rc;
    CU(cudaSetDevice(c->device));
    comm_destroy(c);
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    NC(g_nccl.CommInitRank(&c->comm, world, u, rank));
    c->rank = rank; c->world = world;
    c->pad_cand = c->pad_names = 0;
    c->p2p.failed = false; c->p2p.epoch = 0;
    if (!c->h_gather) CU(cudaMallocHost((void**)&c->h_gather, 4 * sizeof(int64_t) * 1024));
    return CSV_OK;
}

extern "C" int csv_set_gather(csv_ctx* c, int peer_to_peer) {
    if (!c) return set_err(CSV_E_INVALID, "null ctx");
    c->p2p_enabled = peer_to_peer != 0;
    return CSV_OK;
}
extern "C" int csv_gather_mode(csv_ctx* c) { return c && c->p2p_enabled && c->p2p.ready && !c->p2p.failed ? 1 : 0; }

extern "C" int csv_comm_destroy(csv_ctx* c) {
    if (!c) return set_err(CSV_E_INVALID, "null ctx");
    comm_destroy(c);
    c->world = 1; c->rank = 0;
    return CSV_OK;
}

// ---- message of one rank: header | pad_cand x csv_cand | pad_cand x csv_geno | pad_names x int32 ----
static constexpr int64_t GH_WORDS = 4;   // header: n_cand, n_names, overflow, rank (int64 each)
static constexpr int GM_MAX_KEYS = 2048; // (svtype, contig) groups the merge handles through a shared-memory table
struct GatherLayout {
    int64_t pad_cand, pad_names, off_geno, off_names, off_tab, msg_bytes;
    int32_t n_keys;   // 5 * n_contigs when <= GM_MAX_KEYS, else 0 (binary-search merge)
};
// message of one rank: header | pad_cand x csv_cand | pad_cand x csv_geno | pad_names x int32 | n_keys x (first, count)
static GatherLayout gather_layout(int64_t pad_cand, int64_t pad_names, int32_t n_contigs, int world) {
    GatherLayout L;
    L.pad_cand = pad_cand; L.pad_names = pad_names;
    L.n_keys = CSV_NTYPES * n_contigs <= GM_MAX_KEYS ? CSV_NTYPES * n_contigs : 0;
    if ((size_t)L.n_keys * (world + 1) * 4 > 160 * 1024) L.n_keys = 0;   // the merge keeps (world + 1) words per key in shared memory
    L.off_geno = GH_WORDS * 8 + pad_cand * (int64_t)sizeof(csv_cand);
    L.off_names = L.off_geno + pad_cand * (int64_t)sizeof(csv_geno);
    L.off_tab = (L.off_names + pad_names * 4 + 15) / 16 * 16;
    L.msg_bytes = (L.off_tab + (int64_t)L.n_keys * 8 + 15) / 16 * 16;
    return L;
}

// 16 B words: csv_cand = 4, csv_geno = 2.5 (copied as 8 B words: 5)
__global__ void __launch_bounds__(256) k_gather_pack(const csv_cand* __restrict__ cand, const csv_geno* __restrict__ geno,
                                                     const int32_t* __restrict__ names, const Counters* ctr, uint32_t cap_cand,
                                                     uint32_t cap_names, GatherLayout L, int rank, char* __restrict__ msg) {
    const int64_t nc = min(ctr->n_cand, cap_cand), nn = min(ctr->n_names, cap_names);
    const int64_t cc = nc < L.pad_cand ? nc : L.pad_cand, cn = nn < L.pad_names ? nn : L.pad_names;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (tid == 0) {
        int64_t* h = (int64_t*)msg;
        h[0] = nc; h[1] = nn; h[2] = (nc > L.pad_cand || nn > L.pad_names) ? 1 : 0; h[3] = rank;
    }
    const uint4* src_c = (const uint4*)cand;
    uint4* dst_c = (uint4*)(msg + GH_WORDS * 8);
    for (int64_t i = tid; i < cc * 4; i += stride) dst_c[i] = src_c[i];
    if (L.n_keys) {   // (first record, count) of every (svtype, contig) group of this rank: the records are sorted by that key
        uint2* tab = (uint2*)(msg + L.off_tab);   // zeroed before this kernel
        const int32_t nct = L.n_keys / CSV_NTYPES;
        for (int64_t i = tid; i < cc; i += stride) {
            const int32_t sv = cand[i].svtype, ch = cand[i].chrom;
            if (sv < 0 || sv >= CSV_NTYPES || ch < 0 || ch >= nct) continue;
            const int32_t k = sv * nct + ch;
            if (i == 0 || cand[i - 1].svtype != sv || cand[i - 1].chrom != ch) tab[k].x = (uint32_t)i;
            if (i + 1 == cc || cand[i + 1].svtype != sv || cand[i + 1].chrom != ch) tab[k].y = (uint32_t)(i + 1);   // end; count = end - first
        }
    }
    const uint2* src_g = (const uint2*)geno;
    uint2* dst_g = (uint2*)(msg + L.off_geno);
    for (int64_t i = tid; i < cc * 5; i += stride) dst_g[i] = src_g[i];
    int32_t* dst_n = (int32_t*)(msg + L.off_names);
    for (int64_t i = tid; i < cn; i += stride) dst_n[i] = names[i];
}

// Merged order = the single-GPU order: svtype, contig id, then (rank, emission order).  Every rank's records are already
// in that order, so a record's merged position is a sum of binary searches: (records of every rank with a smaller key)
// + (records of lower ranks with the same key) + its offset inside its own key group.  ONE kernel, no scratch: with
// contig shards every (svtype, contig) group comes from one rank, but the rule is well defined for any inputs.
struct MergeJob {
    const char* recv; GatherLayout L; int world; int32_t n_contigs;
    csv_cand* out_c; csv_geno* out_g; int32_t* out_n; int64_t cap_c, cap_n;
    int64_t* hdr;   // [world * GH_WORDS + 1]: compact copy of the headers + a status word (one D2H copy)
    const unsigned long long* flags;   // peer-to-peer path: flags[r] == epoch once rank r's message has landed (null: NCCL path)
    unsigned long long epoch;
};
__device__ __forceinline__ const int64_t* gm_header(const MergeJob& M, int r) { return (const int64_t*)(M.recv + (int64_t)r * M.L.msg_bytes); }
__device__ __forceinline__ int64_t gm_valid(const MergeJob& M, int r) {
    const int64_t n = gm_header(M, r)[0];
    return n < M.L.pad_cand ? n : M.L.pad_cand;
}
__device__ __forceinline__ int64_t gm_key(const csv_cand& c) { return ((int64_t)c.svtype << 32) | (uint32_t)c.chrom; }
// number of records of rank r with key < k (UPPER: <= k)
template <bool UPPER>
__device__ __forceinline__ int64_t gm_bound(const MergeJob& M, int r, int64_t k) {
    const csv_cand* cs = (const csv_cand*)(M.recv + (int64_t)r * M.L.msg_bytes + GH_WORDS * 8);
    int64_t lo = 0, hi = gm_valid(M, r);
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int64_t km = ((int64_t)cs[mid].svtype << 32) | (uint32_t)cs[mid].chrom;
        if (UPPER ? km <= k : km < k) lo = mid + 1; else hi = mid;
    }
    return lo;
}
__global__ void __launch_bounds__(256) k_gather_merge(MergeJob M) {
    pdl_launch_dependents(); pdl_wait();   // programmatic dependent launch: resident early, starts when the previous kernel has finished
    extern __shared__ uint32_t s_gm[];   // table path: [n_keys] exclusive offset of every key group, then [world * n_keys] of every (rank, key)
    __shared__ uint32_t s_warp[9];
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (M.flags) {
        // peer-to-peer gather: every rank stored its message into this rank's mail box over NVLink and then released its flag;
        // wait (system-scope acquire) until all of them have landed.  A peer that never arrives ends the wait after ~1 s
        // with an error status instead of hanging the GPU.
        if ((int)threadIdx.x < M.world) {
            const long long t0 = clock64();
            unsigned long long v;
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(M.flags + threadIdx.x) : "memory");
                if (v != M.epoch && clock64() - t0 > 2000000000ll) { M.hdr[M.world * GH_WORDS] = 3; break; }
            } while (v != M.epoch);
        }
        __syncthreads();
    }
    if (tid < (int64_t)M.world * GH_WORDS) M.hdr[tid] = gm_header(M, (int)(tid / GH_WORDS))[tid % GH_WORDS];
    const int64_t per = M.L.pad_cand, total = per * M.world;
    const int T = M.L.n_keys;
    if (T) {
        // merged position of a record = (records of all ranks with a smaller key) + (records of lower ranks with the same key)
        // + its offset inside its own group: O(1) per record from the (first, end) tables every message carries
        uint32_t* s_key = s_gm;            // [T]
        uint32_t* s_rk = s_gm + T;         // [world * T]
        constexpr int PER = GM_MAX_KEYS / 256;
        uint32_t loc[PER], sum = 0;
#pragma unroll
        for (int j = 0; j < PER; j++) {
            const int k = threadIdx.x * PER + j;
            uint32_t t = 0;
            if (k < T)
                for (int r = 0; r < M.world; r++) {
                    const uint2 e = ((const uint2*)(M.recv + (int64_t)r * M.L.msg_bytes + M.L.off_tab))[k];
                    s_rk[r * T + k] = t;   // records of lower ranks with this key
                    t += e.y - e.x;
                }
            loc[j] = t; sum += t;
        }
        uint32_t tot;
        uint32_t run = block_excl_scan_256(sum, s_warp, &tot);
#pragma unroll
        for (int j = 0; j < PER; j++) {
            const int k = threadIdx.x * PER + j;
            if (k < T) s_key[k] = run;
            run += loc[j];
        }
        __syncthreads();
        const int32_t nct = T / CSV_NTYPES;
        for (int64_t g = tid; g < total; g += stride) {
            const int r = (int)(g / per);
            const int64_t i = g - (int64_t)r * per;
            if (i >= gm_valid(M, r)) continue;
            const char* msg = M.recv + (int64_t)r * M.L.msg_bytes;
            csv_cand c = ((const csv_cand*)(msg + GH_WORDS * 8))[i];
            if (c.svtype < 0 || c.svtype >= CSV_NTYPES || c.chrom < 0 || c.chrom >= nct) { M.hdr[M.world * GH_WORDS] = 1; continue; }
            const int k = c.svtype * nct + c.chrom;
            const uint32_t first = ((const uint2*)(msg + M.L.off_tab))[k].x;
            const int64_t dst = (int64_t)s_key[k] + s_rk[r * T + k] + (i - (int64_t)first);
            int64_t nbase = 0;
            for (int q = 0; q < r; q++) nbase += gm_header(M, q)[1];
            c.names_off += (int32_t)nbase;
            c.reserved[1] = r;
            if (dst < M.cap_c) { M.out_c[dst] = c; M.out_g[dst] = ((const csv_geno*)(msg + M.L.off_geno))[i]; }
            else M.hdr[M.world * GH_WORDS] = 2;
        }
    } else
    for (int64_t g = tid; g < total; g += stride) {   // many contigs: binary searches over the other ranks' records
        const int r = (int)(g / per);
        const int64_t i = g - (int64_t)r * per;
        if (i >= gm_valid(M, r)) continue;
        const char* msg = M.recv + (int64_t)r * M.L.msg_bytes;
        csv_cand c = ((const csv_cand*)(msg + GH_WORDS * 8))[i];
        if (c.svtype < 0 || c.svtype >= CSV_NTYPES || c.chrom < 0 || c.chrom >= M.n_contigs) { M.hdr[M.world * GH_WORDS] = 1; continue; }
        const int64_t k = gm_key(c);
        int64_t dst = i, nbase = 0;   // i = (records of rank r with a smaller key) + offset inside the own group
        for (int q = 0; q < M.world; q++) {
            if (q == r) continue;
            dst += q < r ? gm_bound<true>(M, q, k) : gm_bound<false>(M, q, k);
        }
        for (int q = 0; q < r; q++) nbase += gm_header(M, q)[1];
        c.names_off += (int32_t)nbase;
        c.reserved[1] = r;   // source rank: csv_cand.aux of an INS row indexes THAT rank's INS signature array
        if (dst < M.cap_c) { M.out_c[dst] = c; M.out_g[dst] = ((const csv_geno*)(msg + M.L.off_geno))[i]; }
        else M.hdr[M.world * GH_WORDS] = 2;
    }
    int64_t nbase = 0;
    for (int r = 0; r < M.world; r++) {
        const int64_t nn = gm_header(M, r)[1];
        const int64_t cn = nn < M.L.pad_names ? nn : M.L.pad_names;
        const int32_t* src = (const int32_t*)(M.recv + (int64_t)r * M.L.msg_bytes + M.L.off_names);
        for (int64_t i = tid; i < cn; i += stride)
            if (nbase + i < M.cap_n) M.out_n[nbase + i] = src[i];
        nbase += nn;
    }
}

// ---- peer-to-peer all-gather: no NCCL call in the step -------------------------------------------------------------------
// Every rank owns a mail box (2 buffers x world slots + arrival flags) that all peers map through CUDA IPC.  A step's gather is
// one kernel that packs the message straight into slot [rank] of every peer's box over NVLink / NVSwitch and, when the last
// CTA has finished (system-scope fence), releases flag [rank] = epoch in every box; the merge kernel of each rank waits for
// the world's flags.  Two buffers alternate by step: a rank can only be one step ahead of a peer (its merge of step k
// waited for that peer's flag of step k), so a slot is never overwritten while its owner still reads it.
// pack + push in ONE kernel: every 16 B word of the message is read once from the result arrays and stored straight into
// slot [rank] of every box (own box included); no staging copy of the message, no separate clear of the key table.
// The (first, end) key table is the one part that needs two phases (boundaries are found by the threads that copy the
// records): it is built in a rank-local table, and the CTA that finishes last copies it into the boxes, clears the local
// table for the next step (zero when allocated) and releases the flags.
struct PackPushJob {
    const csv_cand* cand; const csv_geno* geno; const int32_t* names; const Counters* ctr;
    uint32_t cap_cand, cap_names;
    GatherLayout L; int rank, world;
    uint2* tab_local;                    // [L.n_keys], all zero between steps
    char* const* slots; unsigned long long* const* flags; unsigned long long epoch; uint32_t* done;
};
__global__ void __launch_bounds__(256) k_gather_pack_push(PackPushJob J) {
    pdl_launch_dependents();   // the next kernel of the chain may become resident now (it waits for this grid to finish)
    __shared__ uint32_t s_last;
    const GatherLayout& L = J.L;
    const int W = J.world;
    const int64_t nc = min(J.ctr->n_cand, J.cap_cand), nn = min(J.ctr->n_names, J.cap_names);
    const int64_t cc = nc < L.pad_cand ? nc : L.pad_cand, cn = nn < L.pad_names ? nn : L.pad_names;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (tid < W) {
        int64_t* h = (int64_t*)J.slots[tid];
        h[0] = nc; h[1] = nn; h[2] = (nc > L.pad_cand || nn > L.pad_names) ? 1 : 0; h[3] = J.rank;
    }
    const uint4* src_c = (const uint4*)J.cand;
    for (int64_t i = tid; i < cc * 4; i += stride) {
        const uint4 v = src_c[i];
        for (int p = 0; p < W; p++) ((uint4*)(J.slots[p] + GH_WORDS * 8))[i] = v;
    }
    if (L.n_keys) {
        const int32_t nct = L.n_keys / CSV_NTYPES;
        for (int64_t i = tid; i < cc; i += stride) {
            const int32_t sv = J.cand[i].svtype, ch = J.cand[i].chrom;
            if (sv < 0 || sv >= CSV_NTYPES || ch < 0 || ch >= nct) continue;
            const int32_t k = sv * nct + ch;
            if (i == 0 || J.cand[i - 1].svtype != sv || J.cand[i - 1].chrom != ch) J.tab_local[k].x = (uint32_t)i;
            if (i + 1 == cc || J.cand[i + 1].svtype != sv || J.cand[i + 1].chrom != ch) J.tab_local[k].y = (uint32_t)(i + 1);
        }
    }
    const uint2* src_g = (const uint2*)J.geno;
    for (int64_t i = tid; i < cc * 5; i += stride) {
        const uint2 v = src_g[i];
        for (int p = 0; p < W; p++) ((uint2*)(J.slots[p] + L.off_geno))[i] = v;
    }
    for (int64_t i = tid; i < cn; i += stride) {
        const int32_t v = J.names[i];
        for (int p = 0; p < W; p++) ((int32_t*)(J.slots[p] + L.off_names))[i] = v;
    }
    __threadfence_system();
    if (threadIdx.x == 0) s_last = atomicAdd(J.done, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int k = threadIdx.x; k < L.n_keys; k += blockDim.x) {
        const uint2 e = __ldcg(&J.tab_local[k]);
        for (int p = 0; p < W; p++) ((uint2*)(J.slots[p] + L.off_tab))[k] = e;
        J.tab_local[k] = make_uint2(0u, 0u);
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < W) {
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(J.flags[threadIdx.x]), "l"(J.epoch) : "memory");
    }
    if (threadIdx.x == 0) *J.done = 0u;   // ready for the next step (stream order)
}

static void p2p_release(csv_ctx* c) {
    csv_ctx::P2PState& X = c->p2p;
    for (size_t r = 0; r < X.peer.size(); r++)
        if (X.peer[r] && (int)r != c->rank) cudaIpcCloseMemHandle(X.peer[r]);
    X.peer.clear();
    if (X.box) cudaFree(X.box);
    X.box = nullptr;
    X.d_tab.release();
    X.ready = false;
}

// (re)build the mail boxes for messages of msg_bytes: collective (handles travel through one ncclAllGather)
static int p2p_setup(csv_ctx* c, int64_t msg_bytes) {
    csv_ctx::P2PState& X = c->p2p;
    const int W = c->world;
    CU(cudaStreamSynchronize(c->stream));
    p2p_release(c);
    X.slot_bytes = msg_bytes;
    X.flags_off = (size_t)2 * W * msg_bytes;
    const size_t total = X.flags_off + (size_t)2 * W * 8 + 256;
    cudaError_t e = cudaMalloc(&X.box, total);
    if (e != cudaSuccess) return set_err(CSV_E_CUDA, "cudaMalloc(mail box): %s", cudaGetErrorString(e));
    CU(cudaMemset(X.box, 0, total));
    CU(cudaDeviceSynchronize());
    cudaIpcMemHandle_t mine;
    e = cudaIpcGetMemHandle(&mine, X.box);
    // a rank that cannot export still takes part in the exchange (all-zero handle): every rank then falls back to NCCL together
    int ok = e == cudaSuccess ? 1 : 0;
    if (!ok) { cudaGetLastError(); memset(&mine, 0, sizeof(mine)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(c->g_scratch.ensure((size_t)(W + 1) * 80 + (size_t)(W * GH_WORDS + 8) * 8, true));
    char* d_send = c->g_scratch.as<char>() + (size_t)(W * GH_WORDS + 8) * 8;
    char* d_recv = d_send + 80;
    char hbuf[80];
    memset(hbuf, 0, sizeof(hbuf));
    memcpy(hbuf, &mine, 64);
    hbuf[64] = (char)ok;
    CU(cudaMemcpyAsync(d_send, hbuf, 80, cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllGather(d_send, d_recv, 80, ncclUint8, c->comm, c->stream));
    std::vector<char> all((size_t)W * 80);
    CU(cudaMemcpyAsync(all.data(), d_recv, (size_t)W * 80, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int r = 0; r < W; r++) ok &= all[(size_t)r * 80 + 64] ? 1 : 0;
    X.peer.assign(W, nullptr);
    if (ok) {
        for (int r = 0; r < W && ok; r++) {
            if (r == c->rank) { X.peer[r] = X.box; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, all.data() + (size_t)r * 80, 64);
            e = cudaIpcOpenMemHandle(&X.peer[r], h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { cudaGetLastError(); X.peer[r] = nullptr; ok = 0; }
        }
    }
    // every rank must take the same path: agree on "all ranks mapped all boxes"
    int64_t flag = ok;
    CU(cudaMemcpyAsync(d_send, &flag, 8, cudaMemcpyHostToDevice, c->stream));
    NC(g_nccl.AllReduce(d_send, d_send, 1, ncclInt64, ncclMin, c->comm, c->stream));
    CU(cudaMemcpyAsync(&flag, d_send, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (!flag) { p2p_release(c); X.failed = true; return CSV_OK; }   // NCCL path from now on
    // device tables: [buf][peer] slot pointer of this rank, then [buf][peer] flag pointer
    std::vector<void*> tab((size_t)4 * W);
    for (int b = 0; b < 2; b++)
        for (int r = 0; r < W; r++) {
            tab[(size_t)b * W + r] = (char*)X.peer[r] + ((size_t)b * W + c->rank) * msg_bytes;
            tab[(size_t)(2 + b) * W + r] = (char*)X.peer[r] + X.flags_off + ((size_t)b * W + c->rank) * 8;
        }
    CU(X.d_tab.ensure(tab.size() * sizeof(void*) + 64, true));
    CU(cudaMemcpy(X.d_tab.p, tab.data(), tab.size() * sizeof(void*), cudaMemcpyHostToDevice));
    CU(cudaMemset((char*)X.d_tab.p + tab.size() * sizeof(void*), 0, 64));   // the push kernel's done counter
    X.ready = true;
    return CSV_OK;
}

static int gather_enqueue(csv_ctx* c) {
    const GatherLayout L = gather_layout(c->pad_cand, c->pad_names, c->n_contigs, c->world);
    const int W = c->world;
    CU(c->g_scratch.ensure((size_t)(W * GH_WORDS + 8) * 8, true));   // compact headers + status word (zero when allocated)
    CU(c->g_cand.ensure((size_t)L.pad_cand * W * sizeof(csv_cand) + 64));
    CU(c->g_geno.ensure((size_t)L.pad_cand * W * sizeof(csv_geno) + 64));
    CU(c->g_names.ensure((size_t)L.pad_names * W * 4 + 64));
    csv_ctx::P2PState& X = c->p2p;
    const bool want_p2p = c->p2p_enabled && !X.failed && W > 1;
    if (want_p2p && (!X.ready || X.slot_bytes != L.msg_bytes)) {   // first gather / the padding changed: collective re-setup
        int prc = p2p_setup(c, L.msg_bytes);
        if (prc) return prc;
    }
    MergeJob M;
    if (want_p2p && X.ready) {
        // (pack + push into every peer's mail box) -> merge (which waits for the world's flags): two launches and one
        // small D2H copy per step, no NCCL call
        X.epoch++;
        const int b = (int)(X.epoch & 1ull);
        CU(c->g_tab.ensure((size_t)std::max<int64_t>(L.n_keys, 1) * 8, true));   // zero when allocated, re-zeroed by the kernel
        PackPushJob J;
        J.cand = c->cand.as<csv_cand>(); J.geno = c->geno.as<csv_geno>(); J.names = c->names.as<int32_t>();
        J.ctr = c->counters.as<Counters>(); J.cap_cand = c->cap_cand; J.cap_names = c->cap_names;
        J.L = L; J.rank = c->rank; J.world = W;
        J.tab_local = c->g_tab.as<uint2>();
        J.slots = (char* const*)((void**)X.d_tab.p + (size_t)b * W);
        J.flags = (unsigned long long* const*)((void**)X.d_tab.p + (size_t)(2 + b) * W);
        J.epoch = X.epoch;
        J.done = (uint32_t*)((void**)X.d_tab.p + (size_t)4 * W);
        LAUNCH(c, k_gather_pack_push, std::min(c->n_sm * 2, (int)std::max<int64_t>(L.msg_bytes / 16 / 256, 1)), 256, 0, J);
        M.recv = (const char*)X.box + (size_t)b * W * L.msg_bytes;
        M.flags = (const unsigned long long*)((const char*)X.box + X.flags_off) + (size_t)b * W;
        M.epoch = X.epoch;
    } else {
        // clear the key table -> pack -> ONE ncclAllGather -> merge
        CU(c->g_send.ensure((size_t)L.msg_bytes));
        CU(c->g_recv.ensure((size_t)L.msg_bytes * W));
        if (L.n_keys) CU(cudaMemsetAsync(c->g_send.as<char>() + L.off_tab, 0, (size_t)L.n_keys * 8, c->stream));
        LAUNCH(c, k_gather_pack, c->n_sm * 2, 256, 0, c->cand.as<csv_cand>(), c->geno.as<csv_geno>(), c->names.as<int32_t>(),
               c->counters.as<Counters>(), c->cap_cand, c->cap_names, L, c->rank, c->g_send.as<char>());
        NC(g_nccl.AllGather(c->g_send.p, c->g_recv.p, (size_t)L.msg_bytes, ncclUint8, c->comm, c->stream));
        M.recv = c->g_recv.as<char>();
        M.flags = nullptr; M.epoch = 0;
    }
    M.L = L; M.world = W; M.n_contigs = c->n_contigs;
    M.out_c = c->g_cand.as<csv_cand>(); M.out_g = c->g_geno.as<csv_geno>(); M.out_n = c->g_names.as<int32_t>();
    M.cap_c = L.pad_cand * W; M.cap_n = L.pad_names * W;
    M.hdr = c->g_scratch.as<int64_t>();
    const size_t gm_smem = (size_t)L.n_keys * (W + 1) * 4;
    if (gm_smem > 48 * 1024) CU(cudaFuncSetAttribute(k_gather_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gm_smem));
    c->pdl_now = c->pdl_enabled;   // one stream, the merge directly behind the push
    if (M.flags) LAUNCH_PDL(c, k_gather_merge, grid_for(c, std::max<int64_t>(L.pad_cand * W, L.pad_names), 256, 2), 256, gm_smem, M);   // behind k_gather_pack_push
    else LAUNCH(c, k_gather_merge, grid_for(c, std::max<int64_t>(L.pad_cand * W, L.pad_names), 256, 2), 256, gm_smem, M);
    CU(cudaMemcpyAsync(c->h_gather, M.hdr, (size_t)(W * GH_WORDS + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    c->gathered = true;
    return CSV_OK;
}

extern "C" int csv_allgather(csv_ctx* c) {
    if (!c) return set_err(CSV_E_INVALID, "null ctx");
    if (!c->comm) return set_err(CSV_E_STATE, "csv_comm_init has not been called");
    if (!c->ran) return set_err(CSV_E_STATE, "csv_cluster has not been called");
    if (c->world > 1023) return set_err(CSV_E_INVALID, "world size");
    CU(cudaSetDevice(c->device));
    if (c->pad_cand == 0 && getenv("CUTESV_B200_GATHER_PAD")) {   // tests: start from a padding that is too small
        c->pad_cand = std::max<int64_t>(1, atoll(getenv("CUTESV_B200_GATHER_PAD")));
        c->pad_names = c->pad_cand;
    }
    if (c->pad_cand == 0) {
        // first call: agree on the padded message size (max over ranks, with head room); later calls reuse it and
        // carry their true counts in the header, so the steady state is pack -> ONE ncclAllGather -> merge, no host sync
        int64_t nc = 0, nn = 0;
        int rc = csv_result_counts(c, &nc, &nn);
        if (rc) return rc;
        CU(c->g_scratch.ensure((size_t)(c->world * GH_WORDS + 8) * 8, true));
        int64_t h[2] = {nc, nn};
        CU(cudaMemcpyAsync(c->g_scratch.p, h, 16, cudaMemcpyHostToDevice, c->stream));
        NC(g_nccl.AllReduce(c->g_scratch.p, c->g_scratch.p, 2, ncclInt64, ncclMax, c->comm, c->stream));
        CU(cudaMemcpyAsync(h, c->g_scratch.p, 16, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        c->pad_cand = h[0] + h[0] / 4 + 64;
        c->pad_names = h[1] + h[1] / 4 + 1024;
    }
    return gather_enqueue(c);
}

extern "C" int csv_gathered_counts(csv_ctx* c, int64_t* n_cand, int64_t* n_names) {
    if (!c) return set_err(CSV_E_INVALID, "null ctx");
    if (!c->gathered) return set_err(CSV_E_STATE, "csv_allgather has not been called");
    CU(cudaSetDevice(c->device));
    for (int attempt = 0; attempt < 4; attempt++) {
        // the local result must be valid too (input validation, table growth)
        int rc = finish(c);
        if (rc == 1) return set_err(CSV_E_STATE, "csv_allgather: the local result needs a rerun (call csv_result_counts before csv_allgather)");
        if (rc) return rc;
        CU(cudaStreamSynchronize(c->stream));
        bool over = false;
        int64_t tc = 0, tn = 0, mc = 0, mn = 0;
        for (int r = 0; r < c->world; r++) {
            const int64_t* h = c->h_gather + 4 * r;
            over |= h[2] != 0;
            tc += h[0]; tn += h[1];
            mc = std::max(mc, h[0]); mn = std::max(mn, h[1]);
        }
        if (!over) {
            if (c->h_gather[4 * c->world] == 3) {
                cudaMemsetAsync(c->g_scratch.as<int64_t>() + 4 * c->world, 0, 8, c->stream);
                return set_err(CSV_E_CUDA, "csv_allgather: a peer's message did not arrive (peer-to-peer gather timed out)");
            }
            if (c->h_gather[4 * c->world] != 0) {
                cudaMemsetAsync(c->g_scratch.as<int64_t>() + 4 * c->world, 0, 8, c->stream);
                return set_err(CSV_E_CUDA, "csv_allgather: merge failed (status %lld)", (long long)c->h_gather[4 * c->world]);
            }
            c->g_n_cand = tc; c->g_n_names = tn;
            if (n_cand) *n_cand = tc;
            if (n_names) *n_names = tn;
            return CSV_OK;
        }
        // some rank outgrew the padding: every rank sees the same headers, so every rank repeats with the same new size
        c->pad_cand = mc + mc / 4 + 64;
        c->pad_names = mn + mn / 4 + 1024;
        rc = gather_enqueue(c);
        if (rc) return rc;
    }
    return set_err(CSV_E_CUDA, "csv_allgather: message size did not converge");
}

extern "C" int csv_fetch_gathered(csv_ctx* c, csv_cand* cands, csv_geno* genos, int64_t cap_cand, int32_t* names, int64_t cap_names) {
    int64_t nc = 0, nn = 0;
    int rc = csv_gathered_counts(c, &nc, &nn);
    if (rc) return rc;
    if (nc > cap_cand || nn > cap_names) return set_err(CSV_E_CAPACITY, "need %lld candidates / %lld names", (long long)nc, (long long)nn);
    if (nc) {
        CU(cudaMemcpyAsync(cands, c->g_cand.p, (size_t)nc * sizeof(csv_cand), cudaMemcpyDeviceToHost, c->stream));
        CU(cudaMemcpyAsync(genos, c->g_geno.p, (size_t)nc * sizeof(csv_geno), cudaMemcpyDeviceToHost, c->stream));
    }
    if (nn) CU(cudaMemcpyAsync(names, c->g_names.p, (size_t)nn * 4, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return CSV_OK;
}

extern "C" int csv_gathered_device_ptrs(csv_ctx* c, const csv_cand** cands, const csv_geno** genos, const int32_t** names) {
    if (!c || !c->gathered) return set_err(CSV_E_STATE, "no gathered results");
    if (cands) *cands = c->g_cand.as<csv_cand>();
    if (genos) *genos = c->g_geno.as<csv_geno>();
    if (names) *names = c->g_names.as<int32_t>();
    return CSV_OK;
}
