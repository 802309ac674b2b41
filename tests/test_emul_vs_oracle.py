"""CPU: the kernels' per-cluster templates (cutesv_b200/csrc/core.h, one-thread team) and the
pipeline structure (linear keys, binned genotype pass) against the oracle."""
import pytest

import emul_lib
import golden_util
from cutesv_b200 import _abi, synth
from oracle import compare, compare_records, oracle_lib


def _check(cfg, **over):
    kw = dict(cfg["params"])
    kw.update(over)
    p = _abi.default_params(**kw)
    ref = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"])
    got = emul_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"])
    d = compare_records.diff_records(ref, got)
    assert not d, "\n".join(d[:3])


@pytest.mark.parametrize("seed", range(60))
def test_adversarial(seed):
    _check(synth.adversarial(1000 + seed))


@pytest.mark.parametrize("cid,scale", [(2, 0.004), (3, 0.01), (5, 0.002)])
def test_configs(cid, scale):
    _check(synth.make_config(cid, scale))


def test_remain_ratio_and_no_genotype():
    _check(synth.make_config(2, 0.004), remain_reads_ratio=0.6)
    _check(synth.make_config(3, 0.004), genotype=0)


@pytest.mark.parametrize("name", golden_util.case_names())
def test_emulator_matches_reference_rows(name):
    case = golden_util.load_case(name)
    res = emul_lib.cluster(case["params"], case["lens"], case["sigs"], case["reads"])
    d = compare.diff_rows(case["rows"], golden_util.to_rows(case, res))
    assert not d, "\n".join(d)


@pytest.mark.parametrize("seed", range(24))
def test_parameter_sweep(seed):
    """Random flag settings x adversarial inputs (ties, duplicates, pile-ups, half positions)."""
    _check(synth.adversarial(3000 + seed), **synth.random_params(seed))
