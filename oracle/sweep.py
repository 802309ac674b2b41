"""Offline parity sweeps over fresh seeds (authoring container: the live reference at /root/reference is imported).
    python -m oracle.sweep cluster LO HI     kernel-template emulator (tests/emul) vs the C oracle, adversarial inputs x {preset, random flags}
    python -m oracle.sweep reference LO HI   the C oracle vs the live reference's rows on the same inputs
    python -m oracle.sweep extract LO HI [long]   emulator of kernel (a) vs the live reference's parse_read
Prints every mismatch and a summary line.  Two mismatch classes of `reference` are documented deviations:
INS rows tying on (contig, int(pos), len, read) in arbitrary input order, and the reference's own KeyError in overlap_cover
on a zero-width window (zero-length DUP, only with -l 0)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))

import numpy as np  # noqa: E402

from cutesv_b200 import _abi, packing, synth  # noqa: E402
from oracle import compare, compare_extract, compare_records, oracle_lib, ref_harness  # noqa: E402


def _cases(seed):
    for k in range(2):
        cfg = synth.adversarial(seed, max_sigs=160)
        if k:
            cfg["params"] = dict(cfg["params"], **synth.random_params(seed))
        yield k, cfg, _abi.default_params(**cfg["params"])


def cluster(seed):
    import emul_lib
    for k, cfg, p in _cases(seed):
        ref = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"])
        got = emul_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"])
        yield k, compare_records.diff_records(ref, got)


def reference(seed):
    import golden_util
    for k, cfg, p in _cases(seed):
        rows = ref_harness.run_reference(cfg["sigs"], cfg["reads"], cfg["names"], synth.read_name, p)
        exp = {kk: v for kk, v in rows.items() if v}
        res = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"])
        yield k, compare.diff_rows(exp, golden_util.to_rows(dict(sigs=cfg["sigs"], names=cfg["names"], params=p), res))


def extract(seed, kind="short"):
    import emul_lib
    rng = np.random.default_rng(seed)
    p = _abi.default_params(min_size=int(rng.choice([30, 50, 10])), max_size=int(rng.choice([-1, 100000, 2000])),
                            min_mapq=int(rng.choice([20, 0, 30])), max_split_parts=int(rng.choice([7, -1, 2, 3])),
                            min_read_len=int(rng.choice([500, 100])), min_siglength=int(rng.choice([10, 30])),
                            merge_del_threshold=int(rng.choice([0, 500])), merge_ins_threshold=int(rng.choice([100, 500, 0])))
    reads, names, _ = synth.synth_alignments_long(seed, 6) if kind == "long" else synth.synth_alignments(seed, 120)
    rnames = sorted(set(r.query_name for r in reads))
    pk = packing.pack_alignments(reads, {nm: i for i, nm in enumerate(names)}, {nm: i for i, nm in enumerate(rnames)})
    ex = emul_lib.extract(p, pk)
    cigar_of = lambda rec: (pk["cigar"][pk["cigar_off"][rec]:pk["cigar_off"][rec + 1]], int(pk["ref_start"][rec]))  # noqa: E731
    gc, gr = compare_extract.tuples_from_columns(ex, names, rnames, lambda rec: reads[rec].query_sequence, cigar_of,
                                                 (p.min_siglength, p.merge_ins_threshold))
    ref_c, ref_r = ref_harness.run_parse_reads(reads, p)
    yield 0, compare_extract.diff_extract(ref_c, ref_r, gc, gr)


def main():
    mode, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    extra = sys.argv[4:]
    fn = {"cluster": cluster, "reference": reference, "extract": extract}[mode]
    bad = n = 0
    t0 = time.time()
    for seed in range(lo, hi):
        try:
            results = list(fn(seed, *extra))
        except Exception as e:   # the reference itself raises on some degenerate inputs
            results = [(-1, ["EXC %r" % (e,)])]
        for k, d in results:
            n += 1
            if d:
                bad += 1
                print("MISMATCH seed %d variant %d: %s" % (seed, k, str(d)[:400]))
                sys.stdout.flush()
    print("%s sweep [%d, %d): %d cases, %d mismatches, %.0f s" % (mode, lo, hi, n, bad, time.time() - t0))


if __name__ == "__main__":
    main()
