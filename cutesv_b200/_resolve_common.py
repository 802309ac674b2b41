"""Shared body of the drop-in resolution_* functions.

The reference issues one resolution_* call per (svtype, contig) (cuteSV:1113-1199, ~125 calls for a human genome), each of
which re-reads its pickle slice.  Here the FIRST call for a (work dir, svtype, parameters) converts every contig of that type
(column-wise, no per-tuple loop) and runs ONE C-ABI call over all of them; the rows are cached and the other contigs' calls
slice the cache.  A call for a single contig of an uncached work dir costs the same as before."""
import os
from collections import OrderedDict

import numpy as np

from . import _abi, rows, runtime, workdir

_CACHE = OrderedDict()
_CACHE_MAX = 8
N_BATCHED_CALLS = 0     # csv_cluster calls issued by the drop-ins (tests)


def _fingerprint(path, svtype, sigs_index):
    out = []
    for t in (svtype, "reads"):
        fn = "%s%s.pickle" % (path, t)
        try:
            st = os.stat(fn)
            out.append((fn, st.st_size, st.st_mtime_ns))
        except OSError:
            out.append((fn, -1, -1))
    out.append(tuple(sorted(sigs_index.get(svtype, {}).items())))
    return tuple(out)


def _params_key(p):
    return bytes(p)


def clear_cache():
    _CACHE.clear()


def _run_type(path, svtype, params, sigs_index, action, want_reads):
    """All contigs of one SV type of a work dir in ONE csv_cluster call -> {chrom: rows}."""
    global N_BATCHED_CALLS
    chroms_t = list(sigs_index[svtype])
    has_reads = action and want_reads
    # call_gt: `if chr not in sigs_index["reads"]: return []` -> those contigs produce no rows
    run_chroms = [c for c in chroms_t if (not has_reads) or c in sigs_index["reads"]]
    if not run_chroms:
        return {}
    seqs = []
    for c in run_chroms:
        seqs.extend(workdir.load_slice(path, svtype, c, sigs_index))
    reads_rows = []
    if has_reads:
        for c in run_chroms:
            reads_rows.extend(workdir.load_slice(path, "reads", c, sigs_index))
    name_field = {"DEL": 2, "INS": 2, "DUP": 2, "INV": 3, "TRA": 4}[svtype]
    name_id, names = workdir.name_index((seqs, name_field), (reads_rows, 3))
    chroms = sorted(set(run_chroms + ([t[2] for t in seqs] if svtype == "TRA" else [])))
    chrom_id = {c: i for i, c in enumerate(chroms)}
    cols = workdir.tuples_to_columns(svtype, seqs, chrom_id, name_id)
    reads = workdir.reads_to_columns(reads_rows, chrom_id, name_id) if reads_rows else None
    hi = 1
    for arr in (cols["a"], cols["b"]):
        if len(arr):
            hi = max(hi, int(arr.max()))
    if reads is not None and len(reads["end"]):
        hi = max(hi, int(reads["end"].max()))
    eng = runtime.get_engine()
    eng.set_params(params)
    eng.set_contigs(np.full(len(chroms), hi + 2, dtype=np.int64))
    cands, genos, nbuf = eng.cluster({svtype: cols}, reads, type_mask=1 << _abi.TYPE_IDS[svtype])
    N_BATCHED_CALLS += 1
    ins_seq = (lambda i: seqs[i][3]) if svtype == "INS" else None
    out = rows.records_to_rows(cands, genos, nbuf, chroms, lambda i: names[i], ins_seq, bool(action))
    return {c: out.get((svtype, c), []) for c in run_chroms}


def resolve_one(path, chrom, svtype, params, sigs_index, action, want_reads=True):
    """Returns (chrom, rows) like the reference's resolution_* (cuteSV:1116-1189 call sites)."""
    if chrom not in sigs_index[svtype]:
        return (chrom, [])
    key = (os.path.abspath(path) + ("/" if path.endswith("/") else ""), svtype, _params_key(params), bool(action), bool(want_reads),
           _fingerprint(path, svtype, sigs_index))
    hit = _CACHE.get(key)
    if hit is None:
        hit = _run_type(path, svtype, params, sigs_index, action, want_reads)
        _CACHE[key] = hit
        while len(_CACHE) > _CACHE_MAX:
            _CACHE.popitem(last=False)
    else:
        _CACHE.move_to_end(key)
    return (chrom, hit.get(chrom, []))
