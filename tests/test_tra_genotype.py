"""TRA genotyping from a packed all-alignments table: call_gt / count_coverage
(resolveTRA.py:260-309, cuteSV_genotype.py:62-93) against rows the REAL reference produced by
re-opening a (fake-pysam) BAM per candidate (tests/golden/tragt_*.json)."""
import json
import os

import numpy as np
import pytest

import emul_lib
import golden_util
from cutesv_b200 import _abi, rows, synth
from oracle import compare, oracle_lib

CASES = ["adv034", "adv144", "cfg3_s0p004"]


def _aln(reads):
    order = np.lexsort((np.arange(len(reads["chrom"])), reads["start"], reads["chrom"]))
    return {k: v[order] for k, v in reads.items()}


def _golden(name):
    d = json.load(open(os.path.join(golden_util.GOLDEN, "tragt_%s.json" % name)))
    return {tuple(k.split("|")): v for k, v in d.items()}


def _tra_rows(case, res):
    cands, genos, names = res
    got = rows.records_to_rows(cands, genos, names, case["names"], synth.read_name, None, True)
    return {k: v for k, v in got.items() if v and k[0] == "TRA"}


@pytest.mark.parametrize("name", CASES)
def test_oracle_and_emulator_match_reference(name):
    case = golden_util.load_case(name)
    aln = _aln(case["reads"])
    tra_only = {"TRA": case["sigs"]["TRA"]}
    for fn in (oracle_lib.cluster, emul_lib.cluster):
        res = fn(case["params"], case["lens"], tra_only, case["reads"], aln=aln)
        d = compare.diff_rows(_golden(name), _tra_rows(case, res))
        assert not d, "\n".join(d)
        assert not (res[0]["flags"] & _abi.CSV_F_GT_HOST).any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_matches_reference(engine, name):
    case = golden_util.load_case(name)
    engine.set_params(case["params"])
    engine.set_contigs(case["lens"])
    engine.upload_alignments(_aln(case["reads"]))
    try:
        res = engine.cluster({"TRA": case["sigs"]["TRA"]}, case["reads"])
        d = compare.diff_rows(_golden(name), _tra_rows(case, res))
        assert not d, "\n".join(d)
        assert not (res[0]["flags"] & _abi.CSV_F_GT_HOST).any()
    finally:
        engine.upload_alignments(None)
    # without the table the rows stay flagged for the host genotyper
    res = engine.cluster({"TRA": case["sigs"]["TRA"]}, case["reads"])
    assert (res[0]["flags"] & _abi.CSV_F_GT_HOST).all()


@pytest.mark.gpu
def test_cuda_matches_oracle_all_types_config3(engine):
    cfg = synth.make_config(3, 0.05)
    p = _abi.default_params(**cfg["params"])
    aln = _aln(cfg["reads"])
    engine.set_params(p)
    engine.set_contigs(cfg["lens"])
    engine.upload_alignments(aln)
    try:
        from oracle import compare_records
        got = engine.cluster(cfg["sigs"], cfg["reads"])
        ref = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"], n_threads=8, aln=aln)
        d = compare_records.diff_records(ref, got)
        assert not d, "\n".join(d[:3])
        tra = got[0]["svtype"] == _abi.CSV_TRA
        assert tra.sum() > 0 and (got[1]["status"][tra] != 1).all()
    finally:
        engine.upload_alignments(None)


@pytest.mark.gpu
def test_unsorted_alignment_table_is_rejected(engine):
    from cutesv_b200._lib import CuteSVError
    case = golden_util.load_case("adv034")
    engine.set_contigs(case["lens"])
    bad = {k: v[::-1].copy() for k, v in _aln(case["reads"]).items()}
    with pytest.raises(CuteSVError):
        engine.upload_alignments(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_dropin_resolution_tra_with_bam(engine, tmp_path, monkeypatch, name):
    """resolution_TRA(action=True, bam_path=...) : pysam only decodes records around the breakpoints, the
    device genotypes.  Same rows as the reference re-opening the same (fake) BAM."""
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(ROOT, "tests", "fake_pysam"))
    sys.modules.pop("pysam", None)
    from cutesv_b200 import cuteSV_resolveTRA, runtime, workdir
    from oracle import ref_harness
    runtime.set_engine(engine)
    case = golden_util.load_case(name)
    bam = str(tmp_path / "aln.bam")
    ref_harness.write_fake_bam(bam, _aln(case["reads"]), case["names"], case["lens"], synth.read_name)
    s = case["sigs"]["TRA"]
    tuples = [("ABCD"[int(s["c"][i]) & 3], int(s["a"][i]), case["names"][int(s["c"][i]) >> 2], int(s["b"][i]), synth.read_name(int(s["read_id"][i])),
               "TRA", case["names"][int(s["chrom"][i])]) for i in range(len(s["chrom"]))]
    path = str(tmp_path) + "/"
    idx = workdir.write_workdir(path, {"TRA": tuples})
    p = case["params"]
    got = {}
    for chrom in idx["TRA"]:
        r = cuteSV_resolveTRA.run_tra((path, chrom, p.min_support, p.ratio_tra, p.bias_tra, bam, True, p.gt_round, idx))[1]
        if r:
            got[("TRA", chrom)] = r
    d = compare.diff_rows(_golden(name), got)
    assert not d, "\n".join(d)
    sys.modules.pop("pysam", None)
