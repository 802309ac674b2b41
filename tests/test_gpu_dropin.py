"""-m gpu: the reference-facing drop-in entry points (resolution_DEL/INS/INV/DUP/TRA over the
reference's work-dir pickles, cal_GL) against the rows the REAL reference produced (tests/golden)."""
import json
import os

import pytest

import golden_util
from cutesv_b200 import (cuteSV_genotype, cuteSV_resolveDUP, cuteSV_resolveINDEL, cuteSV_resolveINV, cuteSV_resolveTRA, runtime,
                         synth, workdir)
from oracle import compare

pytestmark = pytest.mark.gpu


def _tuples(case):
    """Golden columnar inputs -> the reference's tuple lists (same conversion as the generator used)."""
    names = case["names"]
    out = {}
    s = case["sigs"].get("DEL")
    out["DEL"] = [(int(s["a"][i]), int(s["b"][i]), synth.read_name(int(s["read_id"][i])), "DEL", names[int(s["chrom"][i])])
                  for i in range(len(s["chrom"]))] if s else []
    s = case["sigs"].get("INS")
    seq = golden_util.ins_seq_fn(s) if s else None
    out["INS"] = [((int(s["a"][i]) // 2 if int(s["a"][i]) % 2 == 0 else int(s["a"][i]) / 2), int(s["b"][i]), synth.read_name(int(s["read_id"][i])),
                   seq(i), "INS", names[int(s["chrom"][i])]) for i in range(len(s["chrom"]))] if s else []
    s = case["sigs"].get("DUP")
    out["DUP"] = [(int(s["a"][i]), int(s["b"][i]), synth.read_name(int(s["read_id"][i])), "DUP", names[int(s["chrom"][i])])
                  for i in range(len(s["chrom"]))] if s else []
    s = case["sigs"].get("INV")
    out["INV"] = [("++" if int(s["c"][i]) == 0 else "--", int(s["a"][i]), int(s["b"][i]), synth.read_name(int(s["read_id"][i])), "INV",
                   names[int(s["chrom"][i])]) for i in range(len(s["chrom"]))] if s else []
    s = case["sigs"].get("TRA")
    out["TRA"] = [("ABCD"[int(s["c"][i]) & 3], int(s["a"][i]), names[int(s["c"][i]) >> 2], int(s["b"][i]), synth.read_name(int(s["read_id"][i])),
                   "TRA", names[int(s["chrom"][i])]) for i in range(len(s["chrom"]))] if s else []
    r = case["reads"]
    out["reads"] = [(int(r["start"][i]), int(r["end"][i]), int(r["is_primary"][i]), synth.read_name(int(r["read_id"][i])), names[int(r["chrom"][i])])
                    for i in range(len(r["chrom"]))]
    return out


@pytest.mark.parametrize("name", ["adv001", "adv034", "adv144", "cfg3_s0p004"])
def test_resolution_entry_points(engine, tmp_path, name):
    runtime.set_engine(engine)
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


def test_cal_gl_dropin(engine):
    runtime.set_engine(engine)
    tab = json.load(open(os.path.join(golden_util.GOLDEN, "cal_gl.json")))
    for c0, c1, g, pl, gq, qual in tab[::97] + tab[-6:]:
        r = cuteSV_genotype.cal_GL(c0, c1)
        assert (r[0], r[1], int(r[2]), str(r[3])) == (g, pl, gq, qual), (c0, c1)
