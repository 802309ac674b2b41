"""Shared body of the drop-in resolution_* functions: one (svtype, contig) slice of the reference's
work dir -> columns -> ONE C-ABI call -> the reference's rows."""
import numpy as np

from . import _abi, rows, runtime, workdir


def resolve_one(path, chrom, svtype, params, sigs_index, action, want_reads=True):
    """Returns (chrom, rows) like the reference's resolution_* (cuteSV:1116-1189 call sites)."""
    if chrom not in sigs_index[svtype]:
        return (chrom, [])
    seqs = workdir.load_slice(path, svtype, chrom, sigs_index)
    reads_rows = []
    if action and want_reads:
        if chrom not in sigs_index["reads"]:
            return (chrom, [])  # call_gt: `if chr not in sigs_index["reads"]: return []`
        reads_rows = workdir.load_slice(path, "reads", chrom, sigs_index)
    name_field = {"DEL": 2, "INS": 2, "DUP": 2, "INV": 3, "TRA": 4}[svtype]
    name_id, names = workdir.name_index((seqs, name_field), (reads_rows, 3))
    chroms = sorted(set([chrom] + ([t[2] for t in seqs] if svtype == "TRA" else [])))
    chrom_id = {c: i for i, c in enumerate(chroms)}
    cols = workdir.tuples_to_columns(svtype, seqs, chrom_id, name_id)
    reads = workdir.reads_to_columns(reads_rows, chrom_id, name_id) if reads_rows else None
    hi = 1
    for arr in (cols["a"], cols["b"]):
        if len(arr):
            hi = max(hi, int(arr.max()))
    if svtype == "INS" and len(cols["a"]):
        pass
    if reads is not None and len(reads["end"]):
        hi = max(hi, int(reads["end"].max()))
    eng = runtime.get_engine()
    eng.set_params(params)
    eng.set_contigs(np.full(len(chroms), hi + 2, dtype=np.int64))
    cands, genos, nbuf = eng.cluster({svtype: cols}, reads, type_mask=1 << _abi.TYPE_IDS[svtype])
    ins_seq = (lambda i: seqs[i][3]) if svtype == "INS" else None
    out = rows.records_to_rows(cands, genos, nbuf, chroms, lambda i: names[i], ins_seq, bool(action))
    return (chrom, out.get((svtype, chrom), []))
