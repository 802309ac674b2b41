// extract.cuh -- kernel (a): one warp per alignment record walks the packed CIGAR with a warp
// prefix scan over reference / query offsets (parse_read, cuteSV:606-655), lane 0 runs the
// streaming intra-read merge (generate_combine_sigs, cuteSV:515-575) and the split-read rule
// engine (cuteSV:50-513).  HBM-bound: 4 B per CIGAR op read once, coalesced.
#pragma once
#include "extract_core.h"

namespace csv {

struct ReadView {
    const int32_t *chrom, *ref_start, *ref_end, *flag, *mapq, *query_len, *read_id;
    const int64_t *cigar_off, *sa_off;
    int64_t n;
};

static constexpr int EX_THREADS = 256;

__global__ void __launch_bounds__(EX_THREADS) k_extract(ReadView R, const uint32_t* __restrict__ cigar, SaView sa, ExtractParams P,
                                                        ExtractOut O, int32_t rec_base) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * EX_THREADS + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * EX_THREADS) >> 5;
    for (int64_t rec = warp0; rec < R.n; rec += n_warps) {
        const int32_t flag = R.flag[rec];
        if (flag == 256 || flag == 272) continue;  // single_pipe, cuteSV:711
        const int32_t mapq = R.mapq[rec], qlen = R.query_len[rec], chrom = R.chrom[rec], rid = R.read_id[rec];
        const int32_t ref_start = R.ref_start[rec], ref_end = R.ref_end[rec];
        const bool mq_ok = mapq >= P.min_mapq;
        if (mq_ok && lane == 0) {  // reads_info_list row, cuteSV:729-733
            const uint32_t k = atomicAdd(O.n_rows, 1u);
            if (k < O.cap_rows) {
                O.rr_chrom[k] = chrom; O.rr_start[k] = ref_start; O.rr_end[k] = ref_end; O.rr_id[k] = rid;
                O.rr_prim[k] = (flag == 0 || flag == 16) ? 1 : 0;
            } else atomicOr(O.status, ST_LIST_OVERFLOW);
        }
        if (qlen < P.min_read_len) continue;  // parse_read, cuteSV:607
        ReadCtx RC;
        RC.rec = rec_base + (int32_t)rec; RC.chrom = chrom; RC.rid = rid; RC.qlen = qlen; RC.base_rc = 0;
        const int64_t c_lo = R.cigar_off[rec], c_hi = R.cigar_off[rec + 1];
        int32_t clip_l = 0, clip_r = 0;
        if (mq_ok && c_hi > c_lo) {
            const uint32_t first = cigar[c_lo], last = cigar[c_hi - 1];
            const int fop = first & 15, lop = last & 15;
            const int32_t hard_l = fop == OP_H ? (int32_t)(first >> 4) : 0;
            if (fop == OP_S || fop == OP_H) clip_l = (int32_t)(first >> 4);  // cuteSV:623-626,651-654
            if (lop == OP_S || lop == OP_H) clip_r = (int32_t)(last >> 4);
            MergeState S;
            S.reset();
            InsPiece open_pieces[MAX_OPEN_PIECES];
            int32_t ref = ref_start;
            int64_t q = -(int64_t)hard_l;
            for (int64_t base = c_lo; base < c_hi; base += 32) {
                const int64_t i = base + lane;
                const uint32_t cg = i < c_hi ? cigar[i] : 0u;
                const int op = (i < c_hi) ? (int)(cg & 15) : OP_P;
                const int32_t len = (i < c_hi) ? (int32_t)(cg >> 4) : 0;
                const int32_t radv = op_ref_change(op) ? len : 0;     // cuteSV:633-643
                const int32_t qadv = (op != OP_D) ? len : 0;           // cuteSV:631-632
                int32_t ir = radv, iq = qadv;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int32_t yr = __shfl_up_sync(0xffffffffu, ir, d);
                    const int32_t yq = __shfl_up_sync(0xffffffffu, iq, d);
                    if (lane >= d) { ir += yr; iq += yq; }
                }
                const int32_t sig_start = ref + ir - radv;
                const int64_t shift_after = q + iq;
                const bool qual = len >= P.min_siglength && (op == OP_I || op == OP_D) && i < c_hi;
                uint32_t mask = __ballot_sync(0xffffffffu, qual);
                while (mask) {
                    const int j = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int v_op = __shfl_sync(0xffffffffu, op, j);
                    const int32_t v_len = __shfl_sync(0xffffffffu, len, j);
                    const int32_t v_pos = __shfl_sync(0xffffffffu, sig_start, j);
                    const int64_t v_shift = __shfl_sync(0xffffffffu, shift_after, j);
                    if (lane == 0) {
                        if (v_op == OP_D) push_del(O, RC, P, S, v_pos, v_len);
                        else push_ins(O, RC, P, S, open_pieces, v_pos, v_len, v_shift - v_len, v_shift);
                    }
                }
                ref += __shfl_sync(0xffffffffu, ir, 31);
                q += __shfl_sync(0xffffffffu, iq, 31);
            }
            if (lane == 0) { flush_ins(O, RC, S, open_pieces); flush_del(O, RC, S); }
        }
        const int sig = detect_flag(flag);
        const int64_t s_lo = R.sa_off[rec], s_hi = R.sa_off[rec + 1];
        if ((sig == 1 || sig == 2) && s_hi > s_lo && lane == 0) {  // cuteSV:660-680
            SplitCtx C;
            C.O = &O; C.R = RC; C.P = P;
            C.R.base_rc = sig == 2 ? 1 : 0;
            Seg prim;
            if (sig == 1) { prim.rs = clip_l; prim.re = qlen - clip_r; }
            else { prim.rs = clip_r; prim.re = qlen - clip_l; }
            prim.fs = ref_start; prim.fe = ref_end; prim.chr = chrom; prim.strand = sig == 1 ? 0 : 1;
            organize_split_signal(C, mq_ok, prim, sa, s_lo, s_hi);
        }
        __syncwarp();
    }
}

}  // namespace csv
