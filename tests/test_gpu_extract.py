"""-m gpu: kernel (a) (CIGAR walk + split-read engine) through the C-ABI against the emulator of the
same logic (which tests/test_extract_cpu.py pins against the REAL reference's parse_read) and
against golden tuples from the reference; then extract -> cluster end to end on the device."""
import collections
import json
import os

import numpy as np
import pytest

import emul_lib
import golden_util
from cutesv_b200 import _abi, packing, synth
from oracle import compare_extract, compare_records, oracle_lib

pytestmark = pytest.mark.gpu


def _packet(seed, n=300):
    reads, names, lens = synth.synth_alignments(seed, n)
    rnames = sorted(set(r.query_name for r in reads))
    rid = {nm: i for i, nm in enumerate(rnames)}
    cid = {nm: i for i, nm in enumerate(names)}
    return reads, names, lens, rnames, packing.pack_alignments(reads, cid, rid)


def _canon(ex):
    out = {}
    for t, cols in ex["sigs"].items():
        rows = list(zip(*[cols[k].tolist() for k in ("chrom", "a", "b", "read_id", "c")]))
        if t in ("DEL", "DUP"):
            rows = [r[:4] for r in rows]
        out[t] = collections.Counter(rows)
    r = ex["rows"]
    out["rows"] = collections.Counter(zip(r["chrom"].tolist(), r["start"].tolist(), r["end"].tolist(), r["read_id"].tolist(), r["is_primary"].tolist()))
    return out


@pytest.mark.parametrize("seed", range(12))
def test_extract_matches_emulator(engine, seed):
    reads, names, lens, rnames, pk = _packet(seed)
    rng = np.random.default_rng(seed)
    p = _abi.default_params(min_size=int(rng.choice([30, 50, 10])), max_size=int(rng.choice([-1, 100000, 2000])),
                            min_mapq=int(rng.choice([20, 0, 30])), max_split_parts=int(rng.choice([7, -1, 2, 3])),
                            min_read_len=int(rng.choice([500, 100])), min_siglength=int(rng.choice([10, 30])),
                            merge_del_threshold=int(rng.choice([0, 500])), merge_ins_threshold=int(rng.choice([100, 500, 0])))
    engine.set_params(p)
    engine.set_contigs(lens)
    engine.extract(pk)
    got = engine.fetch_extracted()
    ref = emul_lib.extract(p, pk)
    assert _canon(got) == _canon(ref)
    # INS sequences rebuilt from the piece table
    gc, gr = compare_extract.tuples_from_columns(got, names, rnames, lambda rec: reads[rec].query_sequence)
    rc, rr = compare_extract.tuples_from_columns(ref, names, rnames, lambda rec: reads[rec].query_sequence)
    assert not compare_extract.diff_extract(rc, rr, gc, gr)


@pytest.mark.parametrize("name", ["extract_s0", "extract_s1", "extract_s2"])
def test_extract_matches_reference_golden(engine, name):
    meta = json.load(open(os.path.join(golden_util.GOLDEN, name + ".json")))
    reads, names, lens, rnames, pk = _packet(meta["seed"], meta["n_reads"])
    p = _abi.default_params(**meta["params"])
    engine.set_params(p)
    engine.set_contigs(lens)
    engine.extract(pk)
    got = engine.fetch_extracted()
    gc, gr = compare_extract.tuples_from_columns(got, names, rnames, lambda rec: reads[rec].query_sequence)
    ref_c = {k: [tuple(t) for t in v] for k, v in meta["candidate"].items()}
    ref_r = [tuple(t) for t in meta["rows"]]
    assert not compare_extract.diff_extract(ref_c, ref_r, gc, gr)


def test_extract_then_cluster_on_device(engine):
    """Signatures never leave the GPU between extraction and clustering."""
    reads, names, lens, rnames, pk = _packet(5, 2500)
    p = _abi.default_params(min_support=2, genotype=1, min_mapq=0, min_read_len=100)
    engine.set_params(p)
    engine.set_contigs(lens)
    engine.extract(pk)
    ex = engine.fetch_extracted()
    engine.cluster_device(0x1F)
    got = engine.fetch()
    sigs = {t: dict(v) for t, v in ex["sigs"].items()}
    sigs["DEL"]["c"] = None
    sigs["DUP"]["c"] = None
    ref = oracle_lib.cluster(p, lens, sigs, ex["rows"])
    d = compare_records.diff_records(ref, got)
    assert not d, "\n".join(d[:3])
    assert len(got[0]) > 0
