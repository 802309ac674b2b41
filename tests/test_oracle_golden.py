"""CPU: the oracle (C restatement) against the rows the REAL reference produced (tests/golden)."""
import json
import os

import numpy as np
import pytest

import golden_util
from oracle import compare, oracle_lib


@pytest.mark.parametrize("name", golden_util.case_names())
def test_oracle_matches_reference_rows(name):
    case = golden_util.load_case(name)
    res = oracle_lib.cluster(case["params"], case["lens"], case["sigs"], case["reads"])
    d = compare.diff_rows(case["rows"], golden_util.to_rows(case, res))
    assert not d, "\n".join(d)


def test_cal_gl_golden():
    tab = json.load(open(os.path.join(golden_util.GOLDEN, "cal_gl.json")))
    gt = {"0/0": 0, "0/1": 1, "1/1": 2}
    for c0, c1, g, pl, gq, qual in tab:
        r = oracle_lib.cal_gl(c0, c1)
        assert (int(r["gt"]), "%d,%d,%d" % tuple(r["pl"]), int(r["gq"]), str(float(r["qual"]))) == (gt[g], pl, gq, qual), (c0, c1)


def test_cipos_known_answers():
    kat = json.load(open(os.path.join(golden_util.GOLDEN, "cipos_kat.json")))
    L = oracle_lib.lib()
    for k in kat:
        v = np.array(k["v"], dtype=np.int32)
        std = oracle_lib.np_std(v)
        assert std.hex() == k["std_hex"], len(v)
        x = L.csvo_cal_cipos(std, len(v))
        assert "-%d,%d" % (x, x) == k["cipos"]


def test_pow_half_matches_python():
    L = oracle_lib.lib()
    for n in list(range(1, 5000)) + [2921, 3541, 65535, 1 << 20]:
        assert L.csvo_pow_half(n) == n ** 0.5
