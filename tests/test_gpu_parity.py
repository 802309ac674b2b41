"""-m gpu: the CUDA path through the C-ABI against the oracle (bit-exact records)."""
import numpy as np
import pytest

from cutesv_b200 import _abi, synth
from oracle import compare_records, oracle_lib

pytestmark = pytest.mark.gpu


def _run(engine, cfg, **over):
    kw = dict(cfg["params"])
    kw.update(over)
    p = _abi.default_params(**kw)
    engine.set_params(p)
    engine.set_contigs(cfg["lens"])
    got = engine.cluster(cfg["sigs"], cfg["reads"])
    ref = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"], n_threads=8)
    d = compare_records.diff_records(ref, got)
    assert not d, "\n".join(d[:5])
    return len(ref[0])


@pytest.mark.parametrize("seed", range(40))
def test_adversarial(engine, seed):
    _run(engine, synth.adversarial(seed))


@pytest.mark.parametrize("cid,scale", [(2, 0.004), (3, 0.01), (5, 0.002), (2, 0.05), (3, 0.1)])
def test_configs(engine, cid, scale):
    n = _run(engine, synth.make_config(cid, scale))
    assert n > 0


def test_config2_no_genotype(engine):
    _run(engine, synth.make_config(2, 0.01), genotype=0)


def test_remain_ratio(engine):
    _run(engine, synth.make_config(2, 0.01), remain_reads_ratio=0.6)


def test_pileup_big_clusters(engine):
    """Clusters larger than the warp / CTA shared-memory classes (global-scratch team)."""
    rng = np.random.default_rng(7)
    names, lens = synth.contigs(0.01)
    n_reads = 6000
    reads = dict(chrom=np.zeros(n_reads, np.int32), start=np.full(n_reads, 1000, np.int32) + rng.integers(0, 500, n_reads).astype(np.int32),
                 end=np.full(n_reads, 60000, np.int32), read_id=np.arange(n_reads, dtype=np.int32),
                 is_primary=np.ones(n_reads, np.uint8))
    sigs = {}
    for name, n in (("DEL", 5000), ("INS", 700)):
        pos = 20000 + rng.integers(0, 300, n)
        ln = np.where(rng.random(n) < 0.5, 300, 900) + rng.integers(-20, 20, n)
        rid = rng.integers(0, n_reads, n).astype(np.int32)
        if name == "DEL":
            sigs[name] = dict(chrom=np.zeros(n, np.int32), a=pos.astype(np.int32), b=ln.astype(np.int32), read_id=rid, c=None)
        else:
            sigs[name] = dict(chrom=np.zeros(n, np.int32), a=(2 * pos).astype(np.int32), b=ln.astype(np.int32), read_id=rid,
                              c=ln.astype(np.int32))
    cfg = dict(names=names, lens=lens, reads=reads, sigs=sigs, params=dict(min_support=10, genotype=1))
    assert _run(engine, cfg) > 0


def test_cal_gl_full_domain(engine):
    c0, c1 = np.meshgrid(np.arange(0, 160), np.arange(0, 160), indexing="ij")
    c0, c1 = c0.ravel(), c1.ravel()
    keep = (c0 + c1) > 0
    c0, c1 = c0[keep], c1[keep]
    got = engine.cal_gl(c0, c1)
    for i in range(len(c0)):
        ref = oracle_lib.cal_gl(int(c0[i]), int(c1[i]))
        g = got[i]
        assert (g["gt"], tuple(g["pl"]), g["gq"], g["qual"]) == (ref["gt"], tuple(ref["pl"]), ref["gq"], ref["qual"]), (c0[i], c1[i])


def test_pair_buffer_overflow_and_no_prefilter(monkeypatch):
    """Debug knobs: a tiny (read, window) pair buffer forces the inline-test overflow path of the reads
    pass; disabling the density filter sorts every signature.  Results must not change."""
    from cutesv_b200.engine import Engine
    cfg = synth.make_config(2, 0.05)
    p = _abi.default_params(**cfg["params"])
    ref = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"], n_threads=8)
    for env in ({"CUTESV_B200_PAIR_CAP": "1000"}, {"CUTESV_B200_NO_PREFILTER": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        e = Engine(0, params=p, contig_lens=cfg["lens"])
        for _ in range(2):  # the second run sees the first run's stale buffers
            got = e.cluster(cfg["sigs"], cfg["reads"])
            d = compare_records.diff_records(ref, got)
            assert not d, "\n".join(d[:3])
        e.close()
        for k in env:
            monkeypatch.delenv(k)


def test_edge_cases(engine):
    """Empty and degenerate inputs (the reference returns (chr, []) / no rows)."""
    names, lens = synth.contigs(0.001, 3)
    p = _abi.default_params(min_support=2, genotype=1)
    engine.set_params(p)
    engine.set_contigs(lens)
    empty_reads = dict(chrom=np.zeros(0, np.int32), start=np.zeros(0, np.int32), end=np.zeros(0, np.int32), read_id=np.zeros(0, np.int32),
                       is_primary=np.zeros(0, np.uint8))
    got = engine.cluster({}, empty_reads)
    assert len(got[0]) == 0 and len(got[2]) == 0
    # one signature, then identical signatures from distinct reads, no reads table at all -> CSV_F_NO_READS
    one = dict(chrom=np.array([1], np.int32), a=np.array([500], np.int32), b=np.array([60], np.int32), read_id=np.array([0], np.int32), c=None)
    assert len(engine.cluster({"DEL": one}, empty_reads)[0]) == 0
    same = dict(chrom=np.full(5, 1, np.int32), a=np.full(5, 500, np.int32), b=np.full(5, 60, np.int32), read_id=np.arange(5, dtype=np.int32), c=None)
    cfg = dict(lens=lens, sigs={"DEL": same}, reads=empty_reads, params=dict(min_support=2, genotype=1))
    _run(engine, cfg)
    c, g, n = engine.cluster({"DEL": same}, empty_reads)
    assert len(c) == 1 and c[0]["support"] == 5 and (c[0]["flags"] & _abi.CSV_F_NO_READS)
    # exact duplicates collapse: five copies of one tuple are ONE signature < min_support
    dup = dict(chrom=np.full(5, 1, np.int32), a=np.full(5, 500, np.int32), b=np.full(5, 60, np.int32), read_id=np.zeros(5, np.int32), c=None)
    assert len(engine.cluster({"DEL": dup}, empty_reads)[0]) == 0


def test_genome_larger_than_4gbp_uses_64bit_keys(engine):
    """Total linear length >= 2^32 switches the radix sort to 64-bit keys (12 items per thread)."""
    lens = np.array([2000000000, 2000000000, 1500000000], dtype=np.int64)
    rng = np.random.default_rng(3)
    n_reads = 4000
    chrom = rng.integers(0, 3, n_reads).astype(np.int32)
    start = (rng.random(n_reads) * (lens[chrom] - 30000)).astype(np.int64)
    reads = dict(chrom=chrom, start=start.astype(np.int32), end=(start + 20000).astype(np.int32), read_id=np.arange(n_reads, dtype=np.int32),
                 is_primary=np.ones(n_reads, np.uint8))
    sigs = {}
    for name in ("DEL", "INS"):
        k = rng.integers(0, n_reads, 30000)
        pos = start[k // 30 * 30] + 5000 + rng.integers(0, 40, 30000)   # groups of 30 reads share a locus neighbourhood
        ln = (100 + rng.integers(0, 10, 30000)).astype(np.int32)
        a = pos if name == "DEL" else 2 * pos
        ok = a < 2 ** 31 - 1
        sigs[name] = dict(chrom=chrom[k // 30 * 30][ok], a=a[ok].astype(np.int32), b=ln[ok], read_id=k[ok].astype(np.int32),
                          c=ln[ok] if name == "INS" else None)
    cfg = dict(lens=lens, sigs=sigs, reads=reads, params=dict(min_support=5, genotype=1))
    assert _run(engine, cfg) > 10


def _expand(g):
    """Grouped dict -> plain dict in the same (grouped) row order; `aux` of an INS row indexes these rows."""
    out = {k: v for k, v in g.items() if k != "contig_off"}
    out["chrom"] = np.repeat(np.arange(len(g["contig_off"]) - 1, dtype=np.int32), np.diff(g["contig_off"])).astype(np.int32)
    return out


def _grouped(cfg):
    """(grouped sigs, grouped reads, the same rows as plain columns)."""
    nc = len(cfg["lens"])
    gs = {k: _abi.group_by_contig(v, nc) for k, v in cfg["sigs"].items()}
    gr = _abi.group_by_contig(cfg["reads"], nc)
    return gs, gr, {k: _expand(v) for k, v in gs.items()}, _expand(gr)


@pytest.mark.parametrize("seed", [0, 3, 7, 11, 19])
def test_grouped_inputs_adversarial(engine, seed):
    """csv_cluster_host_grouped (rows grouped by contig + offsets, no contig column) == oracle on the plain columns."""
    cfg = synth.adversarial(seed)
    p = _abi.default_params(**cfg["params"])
    engine.set_params(p)
    engine.set_contigs(cfg["lens"])
    gs, gr, ps, pr = _grouped(cfg)
    got = engine.cluster(gs, gr, grouped=True)
    ref = oracle_lib.cluster(p, cfg["lens"], ps, pr, n_threads=8)
    d = compare_records.diff_records(ref, got)
    assert not d, "\n".join(d[:5])
    # grouping is stable inside a contig: apart from `aux` (a row index) the records equal those of the original order
    ref0 = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"], n_threads=8)
    for f in ("svtype", "chrom", "pos", "len", "support", "cipos", "cilen", "names_cnt"):
        assert np.array_equal(ref0[0][f], got[0][f]), f


@pytest.mark.parametrize("cid,scale", [(2, 0.02), (3, 0.05)])
def test_grouped_inputs_configs(engine, cid, scale):
    cfg = synth.make_config(cid, scale)
    p = _abi.default_params(**cfg["params"])
    engine.set_params(p)
    engine.set_contigs(cfg["lens"])
    gs, gr, ps, pr = _grouped(cfg)
    got = engine.cluster(gs, gr, grouped=True)
    plain = engine.cluster(ps, pr)
    assert not compare_records.diff_records(plain, got)
    ref = oracle_lib.cluster(p, cfg["lens"], ps, pr, n_threads=8)
    assert not compare_records.diff_records(ref, got)


def test_grouped_bad_offsets(engine):
    cfg = synth.make_config(2, 0.004)
    engine.set_params(_abi.default_params(**cfg["params"]))
    engine.set_contigs(cfg["lens"])
    gs, gr, _, _ = _grouped(cfg)
    bad = dict(gs["DEL"])
    off = bad["contig_off"].copy()
    off[-1] -= 1
    bad["contig_off"] = off
    with pytest.raises(RuntimeError):
        engine.cluster({"DEL": bad, "INS": gs["INS"]}, gr, grouped=True)
    off = gs["DEL"]["contig_off"].copy()
    off[1], off[2] = off[2] + 1, off[1]
    bad["contig_off"] = off
    with pytest.raises(RuntimeError):
        engine.cluster({"DEL": bad, "INS": gs["INS"]}, gr, grouped=True)
    got = engine.cluster(gs, gr, grouped=True)   # the ctx still works afterwards
    assert len(got[0]) > 0


def test_lanes_off_same_records(engine):
    """Per-type stream lanes on / off give identical records (order, genotypes, supporting reads)."""
    cfg = synth.make_config(3, 0.05)
    engine.set_params(_abi.default_params(**cfg["params"]))
    engine.set_contigs(cfg["lens"])
    a = engine.cluster(cfg["sigs"], cfg["reads"])
    engine.set_lanes(False)
    try:
        b = engine.cluster(cfg["sigs"], cfg["reads"])
    finally:
        engine.set_lanes(True)
    assert not compare_records.diff_records(a, b)


@pytest.mark.parametrize("seed", range(24))
def test_parameter_sweep(engine, seed):
    """Random flag settings (synth.random_params) x adversarial inputs; the first 12 settings are also pinned against
    the real reference (tests/golden/sweep*.json)."""
    _run(engine, synth.adversarial(3000 + seed), **synth.random_params(seed))
