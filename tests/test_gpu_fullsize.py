"""-m gpu: parity at BASELINE.json's full sizes.  The oracle (C, all host cores) finishes the whole
30x ONT workload in a few seconds on the GPU box, so full-size parity is checked directly
(bit-exact records), plus size-independent properties of the result."""
import numpy as np
import pytest

from cutesv_b200 import _abi, synth
from oracle import compare_records, oracle_lib

pytestmark = pytest.mark.gpu


def _check_properties(cfg, res):
    cands, genos, names = res
    # emission order: svtype, contig id, then non-decreasing cluster order
    key = cands["svtype"].astype(np.int64) * (1 << 32) + cands["chrom"].astype(np.int64)
    assert (np.diff(key) >= 0).all()
    # every supporting read list is a slice of the names buffer; INDEL supports are unique reads
    assert (cands["names_off"] >= 0).all() and (cands["names_off"] + cands["names_cnt"] <= len(names)).all()
    assert (cands["support"] == cands["names_cnt"]).all()
    indel = cands[cands["svtype"] <= 1]
    for c in indel[:: max(len(indel) // 200, 1)]:
        ids = names[c["names_off"]: c["names_off"] + c["names_cnt"]]
        assert len(np.unique(ids)) == len(ids)
    # genotype records: DV == support, likelihoods are phred-scaled (non-negative), DR >= 0
    ok = genos["status"] == 0
    assert (genos["dv"][ok] == cands["support"][ok]).all()
    assert (genos["pl"][ok] >= 0).all() and (genos["dr"][ok] >= 0).all() and (genos["gt"][ok] >= 0).all()


@pytest.mark.parametrize("cid,scale", [(2, 1.0), (3, 1.0), (5, 0.25)])
def test_full_size_parity(engine, cid, scale):
    cfg = synth.make_config(cid, scale)
    p = _abi.default_params(**cfg["params"])
    engine.set_params(p)
    engine.set_contigs(cfg["lens"])
    got = engine.cluster(cfg["sigs"], cfg["reads"])
    _check_properties(cfg, got)
    again = engine.cluster(cfg["sigs"], cfg["reads"])  # idempotent / deterministic (names_off is a layout detail)
    assert not compare_records.diff_records(got, again) and got[1].tobytes() == again[1].tobytes()
    ref = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"], n_threads=0)
    d = compare_records.diff_records(ref, got)
    assert not d, "\n".join(d[:3])
    assert len(got[0]) > 1000
