"""Drop-in for the reference's cuteSV_resolveTRA (resolveTRA.py:30-104,257-258).

Clustering runs on the GPU.  TRA genotyping (call_gt, resolveTRA.py:260-309) depends on the BAM
iteration order and an early-exit counter (count_coverage, cuteSV_genotype.py:72-93), so with
action=True it is completed on the host from the BAM, exactly where the reference does it."""
from . import _abi, cuteSV_genotype
from ._resolve_common import resolve_one


def resolution_TRA(path, chr_1, read_count, overlap_size, max_cluster_bias, bam_path, action, gt_round, sigs_index):
    p = _abi.default_params(min_support=read_count, ratio_tra=overlap_size, bias_tra=max_cluster_bias, genotype=0, gt_round=gt_round)
    chrom, rows_ = resolve_one(path, chr_1, "TRA", p, sigs_index, False, want_reads=False)
    if action:
        for r in rows_:
            names = set(r[11].split(",")) if r[11] else set()
            dv, dr, gt, gl, gq, qual = call_gt(bam_path, int(r[2]), int(r[4]), r[0], r[3], names, max_cluster_bias, gt_round)
            r[6], r[7], r[8], r[9], r[10] = str(dr), str(gt), str(gl), str(gq), str(qual)
    return (chrom, rows_)


def call_gt(bam_path, pos_1, pos_2, chr_1, chr_2, read_id_list, max_cluster_bias, gt_round):
    """Host restatement of resolveTRA.py:260-309 (needs pysam for the BAM)."""
    import pysam
    bam = pysam.AlignmentFile(bam_path)
    try:
        query = set()
        up_bound = cuteSV_genotype.threshold_ref_count(len(read_id_list))
        status = cuteSV_genotype.count_coverage(chr_1, max(int(pos_1) - max_cluster_bias, 0),
                                                min(int(pos_1) + max_cluster_bias, bam.get_reference_length(chr_1)), bam, query,
                                                up_bound, gt_round)
        if status == -1:
            return len(read_id_list), ".", "./.", ".,.,.", ".", "."
        if status != 1:
            cuteSV_genotype.count_coverage(chr_2, max(int(pos_2) - max_cluster_bias, 0),
                                           min(int(pos_2) + max_cluster_bias, bam.get_reference_length(chr_2)), bam, query, up_bound,
                                           gt_round)
        dr = sum(1 for q in query if q not in read_id_list)
        gt, gl, gq, qual = cuteSV_genotype.cal_GL(dr, len(read_id_list))
        return len(read_id_list), dr, gt, gl, gq, qual
    finally:
        bam.close()


def run_tra(args):
    return resolution_TRA(*args)
