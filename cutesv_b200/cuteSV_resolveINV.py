"""Drop-in for the reference's cuteSV_resolveINV (resolveINV.py:6-99,205-206)."""
from . import _abi
from ._resolve_common import resolve_one


def resolution_INV(path, chr, svtype, read_count, max_cluster_bias, sv_size, bam_path, action, MaxSize, gt_round, sigs_index):
    p = _abi.default_params(min_support=read_count, bias_inv=max_cluster_bias, min_size=sv_size, max_size=MaxSize,
                            genotype=1 if action else 0, gt_round=gt_round)
    return resolve_one(path, chr, "INV", p, sigs_index, action)


def run_inv(args):
    return resolution_INV(*args)
