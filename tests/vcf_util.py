"""Shared helpers of the VCF parity tests."""
import numpy as np


def synthetic_reference(names, lens, seed=7):
    rng = np.random.default_rng(seed)
    alphabet = np.array(list("ACGTACGTACGTNRYK"))
    return {n: "".join(alphabet[rng.integers(0, len(alphabet), int(l) + 64)]) for n, l in zip(names, lens)}


def rows_by_chrom(rows):
    """{(type, chrom): rows} -> {chrom: rows in DEL, INS, INV, DUP, TRA order} (cuteSV:1191-1199)."""
    out = {}
    for t in ("DEL", "INS", "INV", "DUP", "TRA"):
        for (tt, chrom), r in sorted(rows.items()):
            if tt == t:
                out.setdefault(chrom, []).extend([list(x) for x in r])
    return out


def normalise_rnames(lines):
    """DUP / BND RNAMES come from Python set iteration in the reference (hash-seed dependent, resolveDUP.py:82,96;
    resolveTRA.py:182): compare those lists as sets."""
    out = []
    for ln in lines:
        if "RNAMES=" in ln and ("SVTYPE=DUP" in ln or "SVTYPE=BND" in ln):
            head, rest = ln.split("RNAMES=", 1)
            names, tail = rest.split(";", 1)
            ln = head + "RNAMES=" + ",".join(sorted(names.split(","))) + ";" + tail
        out.append(ln)
    return out
