"""Loads tests/golden fixtures (inputs + the REAL reference's rows, see oracle/gen_golden.py)."""
import json
import os

import numpy as np

from cutesv_b200 import _abi, rows, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def case_names():
    return json.load(open(os.path.join(GOLDEN, "index.json")))


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
    sigs = {}
    for t in _abi.TYPE_NAMES:
        if "sig_%s_chrom" % t in z:
            sigs[t] = {k: (z["sig_%s_%s" % (t, k)] if "sig_%s_%s" % (t, k) in z else None) for k in ("chrom", "a", "b", "read_id", "c")}
    reads = {k: z["reads_" + k] for k in ("chrom", "start", "end", "read_id", "is_primary")}
    p = _abi.default_params(**meta["params"])
    exp = {}
    for k, v in meta["rows"].items():
        t, c = k.split("|")
        exp[(t, c)] = v
    return dict(lens=z["lens"], sigs=sigs, reads=reads, params=p, names=meta["names"], rows=exp)


def ins_seq_fn(ins_cols):
    """The synthetic INS sequence generator of oracle/ref_harness.py (content-keyed ACGT rotation)."""
    def f(i):
        key = int(ins_cols["a"][i]) + int(ins_cols["read_id"][i]) + int(ins_cols["b"][i])
        n = int(ins_cols["c"][i])
        pat = "ACGT"
        k = key % 4
        return ((pat[k:] + pat[:k]) * (n // 4 + 1))[:n]
    return f


def to_rows(case, result):
    """(cands, genos, names) -> reference rows; TRA rows are compared un-genotyped (the golden
    reference run used action=False for TRA, whose call_gt needs a BAM)."""
    cands, genos, names = result
    ins = case["sigs"].get("INS")
    got = rows.records_to_rows(cands, genos, names, case["names"], synth.read_name, ins_seq_fn(ins) if ins is not None else None,
                               bool(case["params"].genotype))
    return {k: v for k, v in got.items() if v}
