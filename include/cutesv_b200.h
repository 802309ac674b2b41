/*
 * cutesv_b200.h -- C-ABI of the B200-native cuteSV hot path
 * (signature extraction -> sort + chain-linkage clustering -> consensus -> genotype).
 *
 * The reference (tjiangHIT/cuteSV v2.1.4) is pure Python and has no FFI; every entry point
 * below cites the reference interface it replaces ("cuteSV:N" = src/cuteSV/cuteSV line N,
 * other files relative to src/cuteSV/).  The Python binding a maintainer would add is the
 * ctypes stub in cutesv_b200/_lib.py.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; csv_last_error() gives the text.
 *   - the caller owns every host buffer; the library owns all device memory inside a csv_ctx.
 *   - outputs use capacity + "needed" convention (CSV_E_CAPACITY -> retry with larger buffers).
 *   - one csv_ctx per GPU; calls on one ctx are not re-entrant; no callbacks, no exceptions.
 *   - there is NO CPU fallback: csv_create fails if no sm_100 device is usable.
 *
 * Columnar signature layout (int32 columns, one set per SV type; reference tuple types at
 * cuteSV:520-531 (INS/DEL), 235-239 (DUP), 55-60 (INV), 111-117 (TRA)):
 *
 *   type  chrom        a                  b        read_id   c
 *   DEL   contig id    pos                len      name rank -
 *   INS   contig id    2*pos (carries .5) len      name rank len(seq)
 *   INV   contig id    bp1                bp2      name rank strand: 0 "++", 1 "--"
 *   DUP   contig id    pos1               pos2     name rank -
 *   TRA   contig id 1  pos1               pos2     name rank chr2_id*4 + {A:0,B:1,C:2,D:3}
 *
 * contig id = rank of the contig name in Python string order, read_id = rank of the read name
 * in Python string order (the reference sorts tuples with string tie-breaks, cuteSV:764-801).
 *
 * INS rows that tie on (contig, int(pos), len, read_id) must come in the order of their sequence strings: the
 * reference's INS sort key is (chr, int(pos), len, name, seq) (cuteSV:774) and the library never sees the strings,
 * it breaks such ties by input order.  They need one read reporting two insertions of equal length at the same
 * position; the host converters (workdir.py, cli.py) put them in order.
 */
#ifndef CUTESV_B200_H
#define CUTESV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Emission order of the reference's clustering phase (cuteSV:1116-1189). */
enum { CSV_DEL = 0, CSV_INS = 1, CSV_INV = 2, CSV_DUP = 3, CSV_TRA = 4, CSV_NTYPES = 5 };

enum {
    CSV_OK = 0,
    CSV_E_INVALID = -1,  /* bad argument */
    CSV_E_CUDA = -2,     /* CUDA runtime error */
    CSV_E_CAPACITY = -3, /* output buffer too small; sizes are reported back */
    CSV_E_NODEVICE = -4, /* no usable sm_100 device: the product path never falls back to CPU */
    CSV_E_INPUT = -5,    /* device-side validation of the inputs failed (see csv_last_error) */
    CSV_E_STATE = -6     /* call sequence error */
};

/* Flags of csv_cand.flags */
enum {
    CSV_F_NO_READS = 1, /* call_gt(): contig absent from the reads table -> the reference drops
                           every candidate of that contig (resolveINDEL.py:443-444) */
    CSV_F_GT_HOST = 2   /* genotype must be completed on the host (TRA: BAM-order dependent,
                           resolveTRA.py:260-309) */
};

/* POD mirror of the reference flags that reach the hot path (cuteSV_Description.py:53-263;
 * argument wiring at cuteSV:1058-1076 and 1116-1189). */
typedef struct csv_params {
    int32_t min_support;        /* -s: read_count of every resolution_* */
    int32_t min_support_allele; /* minimum_support_reads = min(min_support, 5), cuteSV:1124,1141 */
    int32_t min_size;           /* -l: sv_size */
    int32_t max_size;           /* -L: MaxSize (-1 = unlimited) */
    int32_t bias_del, bias_ins, bias_inv, bias_dup, bias_tra; /* --max_cluster_bias_* */
    int32_t genotype;           /* --genotype: "action" */
    int32_t gt_round;           /* --gt_round (TRA host genotyper only) */
    int32_t gt_bias_ins;        /* constant 1000 of resolveINDEL.py:312 */
    double ratio_del, ratio_ins; /* --diff_ratio_merging_DEL / _INS: threshold_gloab */
    double ratio_tra;            /* --diff_ratio_filtering_TRA: overlap_size */
    double remain_reads_ratio;   /* --remain_reads_ratio */
    /* extraction (cuteSV:606, 697) */
    int32_t min_mapq, max_split_parts, min_read_len, min_siglength;
    int32_t merge_del_threshold, merge_ins_threshold;
    int32_t reserved[2];
} csv_params;

/* One SV type's signature columns (host or device pointers, depending on the call). */
typedef struct csv_sig_cols {
    int64_t n;
    const int32_t* chrom;
    const int32_t* a;
    const int32_t* b;
    const int32_t* read_id;
    const int32_t* c; /* may be NULL for DEL / DUP */
} csv_sig_cols;

/* reads_info_list rows (cuteSV:729-733): (start, end, is_primary, name, chr). */
typedef struct csv_reads_cols {
    int64_t n;
    const int32_t* chrom;
    const int32_t* start;
    const int32_t* end;
    const int32_t* read_id;
    const uint8_t* is_primary;
} csv_reads_cols;

/* Candidate record, 64 B.  One per row returned by resolution_* (resolveINDEL.py:197-205,
 * 408-417; resolveDUP.py:114-118; resolveINV.py:136-143; resolveTRA.py:171-182). */
typedef struct csv_cand {
    int32_t svtype;     /* CSV_* */
    int32_t chrom;      /* contig id (TRA: chr1) */
    int32_t pos;        /* DEL/INS breakpoint, DUP bp1, INV bp1, TRA pos1 */
    int32_t len;        /* DEL: -len (as printed), INS: len, DUP: bp2-bp1, INV: inv_len, TRA: 0 */
    int32_t support;    /* RE / DV */
    int32_t cipos;      /* INDEL: x of "-x,x" */
    int32_t cilen;
    int32_t search_pos; /* INDEL: centre of the genotyping window (resolveINDEL.py:204,415) */
    int32_t pos2;       /* DUP bp2, INV bp2, TRA pos2 */
    int32_t aux;        /* INS: input index of the signature whose seq is the ALT (host slices
                           seq[:len]); INV: strand; TRA: chr2_id*4+type */
    int32_t names_off;  /* slice of the names buffer: supporting read ids in reference order */
    int32_t names_cnt;
    int32_t cluster;    /* ordinal of the chain-linkage cluster this row came from */
    int32_t flags;      /* CSV_F_* */
    int32_t reserved[2]; /* [1]: source rank of a record returned by csv_fetch_gathered */
} csv_cand;

/* assign_gt()/cal_GL() result (cuteSV_genotype.py:33-56,161-173), 40 B. */
typedef struct csv_geno {
    int32_t dr;    /* -1 when genotyping is off or deferred to the host */
    int32_t dv;
    int32_t gt;    /* 0 "0/0", 1 "0/1", 2 "1/1", -1 "./." */
    int32_t pl[3];
    int32_t gq;
    int32_t status; /* 0 filled, 1 not computed */
    double qual;
} csv_geno;

typedef struct csv_ctx csv_ctx;

/* Stages timed with CUDA events on the ctx stream when profiling is on. */
enum {
    CSV_ST_H2D = 0,
    CSV_ST_KEYS,      /* key building + validation */
    CSV_ST_SORT,      /* radix sort passes */
    CSV_ST_SEGMENT,   /* chain-linkage boundary votes + kept-cluster compaction */
    CSV_ST_CLUSTER,   /* per-cluster consensus kernels */
    CSV_ST_ORDER,     /* candidate ordering */
    CSV_ST_GENOTYPE,  /* window binning + reads pass + cal_GL */
    CSV_ST_D2H,
    CSV_ST_EXTRACT,   /* CIGAR / SA walk */
    CSV_ST_COUNT
};

const char* csv_last_error(void);
int csv_version(void);

/* Fill *p with the reference defaults (cuteSV_Description.py:78-262). */
int csv_default_params(csv_params* p);

/* device: CUDA ordinal.  stream: a cudaStream_t to run on (NULL = library-owned stream). */
int csv_create(int device, void* stream, csv_ctx** out);
int csv_destroy(csv_ctx* ctx);
int csv_set_params(csv_ctx* ctx, const csv_params* p);
/* Contig table: id = rank of the name in Python string order; lens from the BAM header
 * (cuteSV:1029).  Needed to linearise (contig, pos) into one sortable coordinate. */
int csv_set_contigs(csv_ctx* ctx, int32_t n_contigs, const int64_t* contig_len);

/* Pinned host memory helpers (caller-owned buffers stay caller-owned). */
int csv_host_alloc(void** p, size_t bytes);
int csv_host_free(void* p);
int csv_host_register(void* p, size_t bytes);
int csv_host_unregister(void* p);

/* Replaces the pickle transport <work_dir>/<TYPE>.pickle + reads.pickle (cuteSV:817-857):
 * async H2D copy of the columns onto the ctx stream.  n == 0 clears the type. */
int csv_upload_sigs(csv_ctx* ctx, int svtype, const csv_sig_cols* host_cols);
int csv_upload_reads(csv_ctx* ctx, const csv_reads_cols* host_cols);
/* The same with rows GROUPED BY CONTIG, as the reference itself holds them (one list per chromosome:
 * the <TYPE>.pickle / reads.pickle files are dicts keyed by chr, cuteSV:817-857): all rows of contig
 * id 0, then id 1, ...; host_cols->chrom is ignored (may be NULL) and contig_off[k] .. contig_off[k+1]
 * (n_contigs + 1 entries, contig_off[0] == 0, contig_off[n_contigs] == n) is the row range of contig k.
 * Saves the 4-byte contig column on the PCIe link; the column is rebuilt on the device. */
int csv_upload_sigs_grouped(csv_ctx* ctx, int svtype, const csv_sig_cols* host_cols, const int64_t* contig_off);
int csv_upload_reads_grouped(csv_ctx* ctx, const csv_reads_cols* host_cols, const int64_t* contig_off);

/* Optional input of the TRA genotyper: ALL alignment records (no mapq filter, every flag) in BAM
 * order, i.e. coordinate-sorted per contig, contigs ascending by id.  is_primary = flag in (0, 16).
 * The reference re-opens the BAM per TRA candidate and iterates bam.fetch() with an early exit
 * (call_gt resolveTRA.py:260-309, count_coverage cuteSV_genotype.py:72-93); with this table the same
 * scan runs on the device.  Without it TRA rows keep CSV_F_GT_HOST.  n == 0 clears the table.
 * Precondition (as for the reads table): one primary record per read name. */
int csv_upload_alignments(csv_ctx* ctx, const csv_reads_cols* aln);

/* Replaces process_process_sigs_type (sort + dedup, cuteSV:750-857) and the whole clustering
 * phase Pool(run_del|run_ins|run_inv|run_dup|run_tra) (cuteSV:1113-1199) including call_gt /
 * overlap_cover / assign_gt / cal_GL (cuteSV_genotype.py:33-173) for every contig at once.
 * Asynchronous on the ctx stream; operates on the device-resident inputs.
 * type_mask: bit t set = run SV type t. */
int csv_cluster(csv_ctx* ctx, uint32_t type_mask);

/* Blocks until csv_cluster finished; reports the result sizes. */
int csv_result_counts(csv_ctx* ctx, int64_t* n_cand, int64_t* n_names);
/* D2H of the results.  Order: svtype ascending (DEL, INS, INV, DUP, TRA), then contig id, then
 * the reference's emission order inside one resolution_* call. */
int csv_fetch(csv_ctx* ctx, csv_cand* cands, csv_geno* genos, int64_t cap_cand, int32_t* names,
              int64_t cap_names);
/* Device pointers of the result buffers (for an NCCL all-gather of candidate records). */
int csv_result_device_ptrs(csv_ctx* ctx, const csv_cand** cands, const csv_geno** genos,
                           const int32_t** names);

/* The reference-facing one-shot call: host columns in, host rows out (H2D + kernels + D2H).
 * sigs[t] may have n == 0.  On CSV_E_CAPACITY *n_cand / *n_names hold the needed sizes. */
int csv_cluster_host(csv_ctx* ctx, const csv_sig_cols sigs[CSV_NTYPES], const csv_reads_cols* reads,
                     uint32_t type_mask, csv_cand* cands, csv_geno* genos, int64_t cap_cand,
                     int32_t* names, int64_t cap_names, int64_t* n_cand, int64_t* n_names);

/* csv_cluster_host over grouped inputs (see csv_upload_sigs_grouped); sig_off[t] may be NULL when sigs[t].n == 0. */
int csv_cluster_host_grouped(csv_ctx* ctx, const csv_sig_cols sigs[CSV_NTYPES], const int64_t* const sig_off[CSV_NTYPES],
                             const csv_reads_cols* reads, const int64_t* reads_off, uint32_t type_mask, csv_cand* cands,
                             csv_geno* genos, int64_t cap_cand, int32_t* names, int64_t cap_names, int64_t* n_cand,
                             int64_t* n_names);

/* cal_GL(c0=DR, c1=DV) (cuteSV_genotype.py:33-56) for n pairs, evaluated on the device. */
int csv_cal_gl(csv_ctx* ctx, const int32_t* c0, const int32_t* c1, int64_t n, csv_geno* out);

/* ---- signature extraction: parse_read / generate_combine_sigs / organize_split_signal /
 * analysis_split_read (cuteSV:50-681) over a packet of decoded alignment records ---- */

/* Per-record header (pysam fields the reference reads at cuteSV:606-680). */
typedef struct csv_read_cols {
    int64_t n;
    const int32_t* chrom;     /* contig id */
    const int32_t* ref_start; /* read.reference_start */
    const int32_t* ref_end;   /* read.reference_end */
    const int32_t* flag;      /* read.flag */
    const int32_t* mapq;
    const int32_t* query_len; /* read.query_length */
    const int32_t* read_id;   /* name rank */
    const int64_t* cigar_off; /* n+1 offsets into cigar[] */
    const int64_t* sa_off;    /* n+1 offsets into the SA segment table */
} csv_read_cols;

/* SA-tag entries reduced on the host with acquire_clip_pos semantics (cuteSV:466-513). */
typedef struct csv_sa_cols {
    int64_t n;
    const int32_t* chrom;      /* contig id of the SA entry */
    const int32_t* pos0;       /* int(seq[1]) - 1 */
    const int32_t* strand;     /* 0 '+', 1 '-' */
    const int32_t* mapq;
    const int32_t* first_clip; /* leading S length */
    const int32_t* last_clip;  /* trailing S length */
    const int32_t* ref_span;   /* sum of M, D, =, X */
} csv_sa_cols;

/* Extracted signatures and reads-table rows REPLACE the device-resident inputs of csv_cluster()
 * (one packet per call); counts per type are reported.  cigar[] is BAM-native u32 = len << 4 | op.
 * Replaces Pool#1 (single_pipe/parse_read per window, cuteSV:1058-1076) without the pickle files.
 * INS sequences are not materialised on the device: every INS signature carries len(seq) in
 * column c plus a list of "pieces" (Python slices of a record's query sequence, or of its
 * reverse complement) from which the host rebuilds the string when it needs it. */
int csv_extract(csv_ctx* ctx, const csv_read_cols* reads, const uint32_t* cigar, int64_t n_cigar,
                const csv_sa_cols* sa, int64_t counts[CSV_NTYPES], int64_t* n_read_rows);
/* Append mode (cuteSV:734-739: every task's signatures are appended to the per-type lists; the reference then
 * concatenates the lists of all tasks, cuteSV:750-762): like csv_extract, but the signatures, INS piece descriptors
 * and reads-table rows of this packet are APPENDED to the device-resident inputs of csv_cluster, so a BAM-to-VCF run
 * never moves a signature column across PCIe.  The record index in an INS piece is the packet-local index plus the
 * number of records of all earlier packets of the accumulation.  counts / n_read_rows report the totals so far.
 * csv_extract_reset (or a plain csv_extract / csv_upload_*) starts a new accumulation. */
int csv_extract_append(csv_ctx* ctx, const csv_read_cols* reads, const uint32_t* cigar, int64_t n_cigar,
                       const csv_sa_cols* sa, int64_t counts[CSV_NTYPES], int64_t* n_read_rows);
int csv_extract_reset(csv_ctx* ctx);
/* Records whose split-read analysis was skipped because they carry more than 64 qualifying segments (only reachable
 * with --max_split_parts -1; their CIGAR signatures are taken).  The reference has no such limit: a documented,
 * counted deviation instead of a failed run. */
int64_t csv_extract_skipped(csv_ctx* ctx);
/* read_id of every device-resident signature / reads-table row: id -> rank[id].  The CLI numbers read names in
 * first-seen order while it decodes and learns their ranks in Python string order at the end (cuteSV:764-801). */
int csv_remap_read_ids(csv_ctx* ctx, const int32_t* rank, int64_t n_rank);
/* Swap rows pairs[2k] <-> pairs[2k+1] of the device-resident INS signatures (columns and piece descriptors), in
 * sequence.  INS rows that tie on (contig, int(pos), len, read) must be in the order of their sequence strings
 * (the reference's sort key ends with the sequence, cuteSV:774); the host, which owns the strings, fixes the few
 * ties of device-extracted rows with this call. */
int csv_swap_ins_rows(csv_ctx* ctx, const int64_t* pairs, int64_t n_pairs);
/* Slices [first, first + count) of the extracted columns of one type / of the piece table (append mode: the rows a
 * packet added). */
int csv_fetch_sigs_range(csv_ctx* ctx, int svtype, int64_t first, int64_t count, int32_t* chrom, int32_t* a, int32_t* b,
                         int32_t* read_id, int32_t* c, int32_t* piece_off, int32_t* piece_cnt);
int csv_fetch_pieces_range(csv_ctx* ctx, int64_t first, int64_t count, int32_t* pieces4);
/* D2H of the extracted signature columns of one type (parity tests, .sigs dumps, host ALT
 * strings).  piece_off / piece_cnt (INS only, may be NULL): slice of the piece table. */
int csv_fetch_sigs(csv_ctx* ctx, int svtype, int64_t cap, int32_t* chrom, int32_t* a, int32_t* b,
                   int32_t* read_id, int32_t* c, int32_t* piece_off, int32_t* piece_cnt);
/* Piece table: 4 int32 per piece = (record index, slice start, slice stop, reverse-complement flag),
 * Python slice semantics (negative indices allowed).  *n_pieces reports the table size.
 * flag == 2 marks a signature that merged more CIGAR insertions than the device buffers (64): `slice start` is then
 * the reference position of the merged group and the host rebuilds the string by walking that record's CIGAR
 * (cutesv_b200/packing.py merged_ins_from_cigar); position, length and len(seq) of the signature are complete. */
int csv_fetch_pieces(csv_ctx* ctx, int64_t cap, int32_t* pieces4, int64_t* n_pieces);
int csv_fetch_read_rows(csv_ctx* ctx, int64_t cap, int32_t* chrom, int32_t* start, int32_t* end,
                        int32_t* read_id, uint8_t* is_primary);

/* Profiling: per-stage device milliseconds of the last csv_cluster / csv_extract call. */
int csv_set_profiling(csv_ctx* ctx, int on);
int csv_stage_ms(csv_ctx* ctx, float ms[CSV_ST_COUNT]);
/* Per-kernel totals of the last profiled call(s): with profiling on, every kernel launch sits between its own
 * pair of CUDA events on its launching stream; one text line per kernel, "name<TAB>launches<TAB>total_ms".
 * Returns the buffer size the text needs (incl. NUL). */
int64_t csv_kernel_times(csv_ctx* ctx, char* buf, int64_t cap);
/* SV types are independent until the final ordering (the reference runs them as separate Pool#3
 * tasks, cuteSV:1113-1199); by default each type's kernel chain runs on its own stream ("lane").
 * on = 0 serialises the types on the ctx stream, e.g. to time every kernel alone. */
int csv_set_lanes(csv_ctx* ctx, int on);
/* Number of kernels the library launched since the ctx was created (kernels inside a replayed CUDA graph count). */
int64_t csv_launch_count(csv_ctx* ctx);
/* csv_cluster calls served by replaying a captured CUDA graph.  A call whose inputs are already device resident
 * replays the graph of its (type_mask, sizes, params) from its third occurrence on; CUTESV_B200_GRAPHS=0 disables. */
int64_t csv_graph_replays(csv_ctx* ctx);

/* ---- multi-GPU: contigs sharded over ranks, one csv_ctx per GPU / process -------------------------------------
 * Every resolution_* call of the reference is keyed by (svtype, chr) and reads only that contig's signatures and
 * reads-table rows (cuteSV:1116-1189; resolveINDEL.py:52-54,445-447), so contigs are independent units: each rank
 * runs the whole pipeline on its contigs with no data-path collective, and ONE all-gather of the final
 * records assembles the result on every rank in the single-GPU order.  (The reference's Pool(threads).map_async
 * over (type, chr) tasks, cuteSV:1113-1199, is the CPU counterpart.) */

/* owned[k] != 0: contig k belongs to this ctx's shard (NULL = all).  Contigs outside the shard take no room in the
 * linear coordinate, so histogram / bin tables scale with the shard; a signature or reads-table row on such a
 * contig is an input error (CSV_E_INPUT).  Contig ids stay global.  Call after csv_set_contigs. */
int csv_set_shard(csv_ctx* ctx, const uint8_t* owned);
/* ncclGetUniqueId: rank 0 calls it and ships the bytes (>= 128) to the other ranks by any means. */
int csv_comm_unique_id(void* id, size_t bytes);
/* ncclCommInitRank on the ctx's device.  Collective: every rank calls it with the same id. */
int csv_comm_init(csv_ctx* ctx, const void* id, int rank, int world);
int csv_comm_destroy(csv_ctx* ctx);
/* After csv_cluster, asynchronous on the ctx stream, collective: packs this rank's records (csv_cand, csv_geno,
 * supporting read ids) into one padded message, gathers the messages of all ranks, then merges them on the device
 * into the single-GPU order (svtype, contig id, emission order).  The gather itself is either ONE ncclAllGather over
 * NVLink or (default when every rank could map every rank's mail box through CUDA IPC) ONE kernel that packs the
 * records straight into the peers' mail boxes over NVLink and releases an arrival flag (no staging copy); csv_set_gather(ctx, 0) or
 * CUTESV_B200_GATHER=nccl selects NCCL, csv_gather_mode() tells which one the last gather used (1 peer-to-peer).  csv_cand.reserved[1] of a
 * gathered record is its source rank (csv_cand.aux of an INS row indexes THAT rank's INS signatures).  The padded
 * message size is agreed once (first call: one count all-reduce) and re-agreed only when a rank outgrows it. */
int csv_allgather(csv_ctx* ctx);
int csv_set_gather(csv_ctx* ctx, int peer_to_peer);
int csv_gather_mode(csv_ctx* ctx);
/* Blocks until the gather finished; total sizes over all ranks. */
int csv_gathered_counts(csv_ctx* ctx, int64_t* n_cand, int64_t* n_names);
int csv_fetch_gathered(csv_ctx* ctx, csv_cand* cands, csv_geno* genos, int64_t cap_cand, int32_t* names, int64_t cap_names);
int csv_gathered_device_ptrs(csv_ctx* ctx, const csv_cand** cands, const csv_geno** genos, const int32_t** names);
/* Device counters of the last finished csv_cluster call, 32 words: [0] status, [1] candidates,
 * [2] names, [3] max support, [4..8] kept clusters per type, [9..13] CTA-class clusters,
 * [14..18] global-scratch-class clusters, [19] (read, window) pairs, [20..24] sorted-domain size per
 * type, [25..29] signatures inside kept clusters per type. */
int csv_debug_counters(csv_ctx* ctx, uint32_t out[32]);
/* Duration (ms) and element count of the last radix scatter pass launches (roofline probe). */
int csv_sort_probe(csv_ctx* ctx, float* ms_total, int64_t* bytes_total, int32_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* CUTESV_B200_H */
