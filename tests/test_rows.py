"""rows.records_to_rows (column-wise) == rows.records_to_rows_slow (record_to_row per candidate) on oracle results that hold every SV type,
genotyped and not, incl. candidates dropped for contigs without reads-table rows."""
import numpy as np
import pytest

from cutesv_b200 import _abi, rows, synth
from oracle import oracle_lib

import golden_util


def _both(cfg, genotype):
    params = dict(cfg["params"])
    params["genotype"] = genotype
    p = _abi.default_params(**params)
    r = cfg["reads"]
    order = np.lexsort((np.arange(len(r["chrom"])), r["start"], r["chrom"]))
    aln = {k: v[order] for k, v in r.items()} if "TRA" in cfg["sigs"] else None
    cands, genos, names = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"], aln=aln)[:3]
    chrom_names = ["c%02d" % i for i in range(len(cfg["lens"]))]
    ins = cfg["sigs"].get("INS")
    seq = golden_util.ins_seq_fn(ins) if ins is not None and ins.get("c") is not None else (lambda i: "ACGT" * 4000)
    a = rows.records_to_rows(cands, genos, names, chrom_names, synth.read_name, seq, bool(genotype))
    b = rows.records_to_rows_slow(cands, genos, names, chrom_names, synth.read_name, seq, bool(genotype))
    return a, b, len(cands)


@pytest.mark.parametrize("genotype", [1, 0])
def test_rows_columnwise_equals_per_record(genotype):
    total = 0
    for cfg in [synth.make_config(3, 0.01)] + [synth.adversarial(s) for s in (1, 5, 13, 34)]:
        a, b, n = _both(cfg, genotype)
        assert a == b
        assert list(a.keys()) == list(b.keys())   # emission order of the (type, contig) groups too
        total += n
    assert total > 200
