"""Drop-in for the reference's cuteSV_resolveINDEL (resolveINDEL.py:17-108, 222-317, 435-439):
same names, same argument meaning, same returned rows -- computed by the CUDA path."""
from . import _abi
from ._resolve_common import resolve_one


def resolution_DEL(path, chr, svtype, read_count, threshold_gloab, max_cluster_bias, minimum_support_reads, bam_path, action,
                   gt_round, remain_reads_ratio, sigs_index):
    p = _abi.default_params(min_support=read_count, min_support_allele=minimum_support_reads, ratio_del=threshold_gloab,
                            bias_del=max_cluster_bias, genotype=1 if action else 0, gt_round=gt_round,
                            remain_reads_ratio=remain_reads_ratio)
    return resolve_one(path, chr, "DEL", p, sigs_index, action)


def resolution_INS(path, chr, svtype, read_count, threshold_gloab, max_cluster_bias, minimum_support_reads, bam_path, action,
                   gt_round, remain_reads_ratio, sigs_index):
    p = _abi.default_params(min_support=read_count, min_support_allele=minimum_support_reads, ratio_ins=threshold_gloab,
                            bias_ins=max_cluster_bias, genotype=1 if action else 0, gt_round=gt_round,
                            remain_reads_ratio=remain_reads_ratio)
    return resolve_one(path, chr, "INS", p, sigs_index, action)


def run_del(args):
    return resolution_DEL(*args)


def run_ins(args):
    return resolution_INS(*args)
