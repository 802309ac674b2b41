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
            // 128 CIGAR ops per warp iteration: every lane owns 4 consecutive ops, so one warp scan
            // (5 shuffle steps per offset) is amortised over 128 ops
            // (the next 128 ops are requested before the current ones are scanned: with 24 resident warps per SM a
            //  single 512 B request per warp is far too little in flight for HBM latency)
            uint32_t nxt[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int64_t i = c_lo + lane * 4 + j;
                nxt[j] = i < c_hi ? __ldg(&cigar[i]) : (uint32_t)OP_P;
            }
            for (int64_t base = c_lo; base < c_hi; base += 128) {
                int op[4];
                int32_t len[4], radv[4], qadv[4];
                int32_t r_tot = 0, q_tot = 0;
                uint32_t cur[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    cur[j] = nxt[j];
                    const int64_t i = base + 128 + lane * 4 + j;
                    nxt[j] = i < c_hi ? __ldg(&cigar[i]) : (uint32_t)OP_P;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t cg = cur[j];
                    op[j] = (int)(cg & 15);
                    len[j] = (int32_t)(cg >> 4);
                    radv[j] = op_ref_change(op[j]) ? len[j] : 0;     // cuteSV:633-643
                    qadv[j] = (op[j] != OP_D) ? len[j] : 0;           // cuteSV:631-632
                    r_tot += radv[j];
                    q_tot += qadv[j];
                }
                int32_t ir = r_tot, iq = q_tot;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int32_t yr = __shfl_up_sync(0xffffffffu, ir, d);
                    const int32_t yq = __shfl_up_sync(0xffffffffu, iq, d);
                    if (lane >= d) { ir += yr; iq += yq; }
                }
                int32_t sig_start[4];
                int64_t shift_after[4];
                uint32_t qmask = 0;
                int32_t run_r = ref + ir - r_tot;
                int64_t run_q = q + iq - q_tot;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    sig_start[j] = run_r;
                    run_r += radv[j];
                    run_q += qadv[j];
                    shift_after[j] = run_q;
                    if (len[j] >= P.min_siglength && (op[j] == OP_I || op[j] == OP_D) && base + lane * 4 + j < c_hi) qmask |= 1u << j;
                }
                uint32_t lanes = __ballot_sync(0xffffffffu, qmask != 0);
                while (lanes) {  // qualifying ops are rare (~2 per read): serial hand-over to lane 0, in read order
                    const int L = __ffs(lanes) - 1;
                    lanes &= lanes - 1;
                    const uint32_t m4 = __shfl_sync(0xffffffffu, qmask, L);
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (!(m4 >> j & 1u)) continue;
                        const int v_op = __shfl_sync(0xffffffffu, op[j], L);
                        const int32_t v_len = __shfl_sync(0xffffffffu, len[j], L);
                        const int32_t v_pos = __shfl_sync(0xffffffffu, sig_start[j], L);
                        const int64_t v_shift = __shfl_sync(0xffffffffu, shift_after[j], L);
                        if (lane == 0) {
                            if (v_op == OP_D) push_del(O, RC, P, S, v_pos, v_len);
                            else push_ins(O, RC, P, S, open_pieces, v_pos, v_len, v_shift - v_len, v_shift);
                        }
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
