"""CPU: the extraction logic (cutesv_b200/csrc/extract_core.h through the emulator) against golden
tuples produced by the REAL reference's parse_read, and against the reference itself when present."""
import json
import os

import numpy as np
import pytest

import emul_lib
import golden_util
from cutesv_b200 import _abi, packing, synth
from oracle import compare_extract, ref_harness


EXTRACT_GOLDENS = ["extract_s0", "extract_s1", "extract_s2", "extract_s3", "extract_s4", "extract_s5", "extract_s6",
                   "extract_l0", "extract_l1", "extract_l2", "extract_l3", "extract_l4"]


def _run(seed, n, p, kind="short"):
    reads, names, lens = synth.synth_alignments_long(seed, n) if kind == "long" else synth.synth_alignments(seed, n)
    rnames = sorted(set(r.query_name for r in reads))
    rid = {nm: i for i, nm in enumerate(rnames)}
    cid = {nm: i for i, nm in enumerate(names)}
    pk = packing.pack_alignments(reads, cid, rid)
    ex = emul_lib.extract(p, pk)
    cigar_of = lambda rec: (pk["cigar"][pk["cigar_off"][rec]:pk["cigar_off"][rec + 1]], int(pk["ref_start"][rec]))
    return reads, compare_extract.tuples_from_columns(ex, names, rnames, lambda rec: reads[rec].query_sequence, cigar_of,
                                                      (p.min_siglength, p.merge_ins_threshold)), ex


@pytest.mark.parametrize("name", EXTRACT_GOLDENS)
def test_emulator_matches_reference_golden(name):
    """short packets: every flag / strand branch at scale; long ones (extract_l*): BASELINE config-5-shaped records
    (>= 10^4 CIGAR ops, clips on both ends, 2-6 SA segments in every strand pattern, MaxSize -1, chains of > 64 merged
    insertions whose sequence the host rebuilds from the CIGAR)."""
    meta = json.load(open(os.path.join(golden_util.GOLDEN, name + ".json")))
    p = _abi.default_params(**meta["params"])
    reads, (gc, gr), ex = _run(meta["seed"], meta["n_reads"], p, meta.get("kind", "short"))
    if name in ("extract_l0", "extract_l1", "extract_l4"):
        assert (ex["pieces"][:, 3] == 2).any(), "the chained-insertion spill path should be exercised"
    ref_c = {k: [tuple(t) for t in v] for k, v in meta["candidate"].items()}
    ref_r = [tuple(t) for t in meta["rows"]]
    assert not compare_extract.diff_extract(ref_c, ref_r, gc, gr)


@pytest.mark.skipif(not ref_harness.available(), reason="reference not present (GPU box)")
@pytest.mark.parametrize("seed", range(500, 520))
def test_emulator_matches_live_reference(seed):
    rng = np.random.default_rng(seed)
    p = _abi.default_params(min_size=int(rng.choice([30, 50, 10])), max_size=int(rng.choice([-1, 100000, 2000])),
                            min_mapq=int(rng.choice([20, 0, 30])), max_split_parts=int(rng.choice([7, -1, 2, 3])),
                            min_read_len=int(rng.choice([500, 100])), min_siglength=int(rng.choice([10, 30])),
                            merge_del_threshold=int(rng.choice([0, 500])), merge_ins_threshold=int(rng.choice([100, 500, 0])))
    reads, (gc, gr), _ = _run(seed, 120, p)
    ref_c, ref_r = ref_harness.run_parse_reads(reads, p)
    assert not compare_extract.diff_extract(ref_c, ref_r, gc, gr)


def test_acquire_clip_pos():
    assert packing.acquire_clip_pos("10S100M5D20M3S") == (10, 3, 125)
    assert packing.acquire_clip_pos("10H100M") == (0, 0, 100)
    assert packing.acquire_clip_pos("5=2X3I4S") == (0, 4, 7)
