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
