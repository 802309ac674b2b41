"""CPU: the drop-in resolution_* entry points (work-dir pickles -> columns -> one cluster call -> the reference's rows)
with the kernels replaced by the pipeline emulator, against the rows the REAL reference produced (tests/golden).
The -m gpu twin is tests/test_gpu_dropin.py."""
import pytest

import golden_util
from cutesv_b200 import cuteSV_resolveDUP, cuteSV_resolveINDEL, cuteSV_resolveINV, cuteSV_resolveTRA, runtime, workdir
from emul_engine import EmulEngine
from oracle import compare
from test_gpu_dropin import _tuples


@pytest.mark.parametrize("name", ["adv001", "adv034", "adv144", "cfg3_s0p004", "sweep05", "sweep11"])
def test_resolution_entry_points_cpu(tmp_path, name):
    runtime.set_engine(EmulEngine())
    try:
        case = golden_util.load_case(name)
        p = case["params"]
        path = str(tmp_path) + "/"
        idx = workdir.write_workdir(path, _tuples(case))
        action = bool(p.genotype)
        got = {}
        for chrom in idx["DEL"]:
            got[("DEL", chrom)] = cuteSV_resolveINDEL.run_del((path, chrom, "DEL", p.min_support, p.ratio_del, p.bias_del, p.min_support_allele,
                                                               "", action, p.gt_round, p.remain_reads_ratio, idx))[1]
        for chrom in idx["INS"]:
            got[("INS", chrom)] = cuteSV_resolveINDEL.run_ins((path, chrom, "INS", p.min_support, p.ratio_ins, p.bias_ins, p.min_support_allele,
                                                               "", action, p.gt_round, p.remain_reads_ratio, idx))[1]
        for chrom in idx["INV"]:
            got[("INV", chrom)] = cuteSV_resolveINV.run_inv((path, chrom, "INV", p.min_support, p.bias_inv, p.min_size, "", action, p.max_size,
                                                             p.gt_round, idx))[1]
        for chrom in idx["DUP"]:
            got[("DUP", chrom)] = cuteSV_resolveDUP.run_dup((path, chrom, p.min_support, p.bias_dup, p.min_size, "", action, p.max_size, p.gt_round,
                                                             idx))[1]
        for chrom in idx["TRA"]:
            got[("TRA", chrom)] = cuteSV_resolveTRA.run_tra((path, chrom, p.min_support, p.ratio_tra, p.bias_tra, "", False, p.gt_round, idx))[1]
        got = {k: v for k, v in got.items() if v}
        d = compare.diff_rows(case["rows"], got)
        assert not d, "\n".join(d[:4])
    finally:
        runtime.set_engine(None)
