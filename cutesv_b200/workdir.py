"""The reference's work-dir transport (cuteSV:817-857): <TYPE>.pickle = concatenated pickled
per-contig lists of signature tuples, sigindex = {type: {chr: byte offset}}.  Readers for the
drop-in resolution_* entry points, a writer for --retain_work_dir compatibility and tests, and the
tuple <-> column conversion (names -> ranks in Python string order)."""
import pickle

import numpy as np

from . import _abi

TYPES = ("DEL", "INS", "DUP", "INV", "TRA")
_TRA = {"A": 0, "B": 1, "C": 2, "D": 3}


def sort_key(svtype):
    """Sort keys of process_process_sigs_type (cuteSV:764,774,783,792,801)."""
    if svtype == "DEL":
        return lambda x: (x[-1], int(x[0]), x[1], x[2])
    if svtype == "INS":
        return lambda x: (x[-1], int(x[0]), x[1], x[2], x[3])
    if svtype == "DUP":
        return lambda x: (x[-1], int(x[0]), int(x[1]), x[2])
    if svtype == "INV":
        return lambda x: (x[-1], x[0], int(x[1]), x[2], x[3])
    if svtype == "TRA":
        return lambda x: (x[-1], x[2], x[0], int(x[1]), x[3], x[4], x[5])
    return lambda x: (x[-1])


def write_type(path, svtype, tuples):
    """Write <path><svtype>.pickle in the reference layout; returns (index, reads_count)."""
    cand = sorted(tuples, key=sort_key(svtype))
    if svtype != "reads":  # remove_duplicates_sorted, cuteSV:958-969
        dedup = []
        for t in cand:
            if not dedup or dedup[-1] != t:
                dedup.append(t)
        cand = dedup
    index, counts = {}, {}
    with open("%s%s.pickle" % (path, svtype), "wb") as f:
        start = 0
        i = 0
        while i < len(cand):
            j = i
            while j < len(cand) and cand[j][-1] == cand[i][-1]:
                j += 1
            blob = pickle.dumps(cand[i:j])
            f.write(blob)
            index[cand[i][-1]] = start
            counts[cand[i][-1]] = j - i
            start += len(blob)
            i = j
    return index, counts


def write_workdir(path, tuples_by_type):
    """tuples_by_type: {"DEL": [...], ..., "reads": [...]} -> sigs_index (also pickled as sigindex.pickle)."""
    sigs_index = {}
    for t in TYPES + ("reads",):
        idx, cnt = write_type(path, t, tuples_by_type.get(t, []))
        sigs_index[t] = idx
        if t == "reads":
            sigs_index["reads_count"] = cnt
    with open("%ssigindex.pickle" % path, "wb") as f:
        pickle.dump(sigs_index, f)
    return sigs_index


def load_slice(path, svtype, chrom, sigs_index):
    with open("%s%s.pickle" % (path, svtype), "rb") as f:
        f.seek(sigs_index[svtype][chrom])
        return pickle.load(f)


def name_index(*tuple_lists_and_fields):
    """Rank of every read name in Python string order over several (list, field index) pairs."""
    names = set()
    for lst, k in tuple_lists_and_fields:
        for t in lst:
            names.add(t[k])
    ordered = sorted(names)
    return {n: i for i, n in enumerate(ordered)}, ordered


def tuples_to_columns(svtype, tuples, chrom_id, name_id):
    n = len(tuples)
    cols = dict(chrom=np.zeros(n, np.int32), a=np.zeros(n, np.int32), b=np.zeros(n, np.int32), read_id=np.zeros(n, np.int32),
                c=np.zeros(n, np.int32) if svtype in ("INS", "INV", "TRA") else None)
    for i, t in enumerate(tuples):
        cols["chrom"][i] = chrom_id[t[-1]]
        if svtype == "DEL" or svtype == "DUP":
            cols["a"][i], cols["b"][i], cols["read_id"][i] = int(t[0]), int(t[1]), name_id[t[2]]
        elif svtype == "INS":
            cols["a"][i] = int(round(float(t[0]) * 2))  # positions from split reads can be x.5 (cuteSV:228,244)
            cols["b"][i], cols["read_id"][i], cols["c"][i] = int(t[1]), name_id[t[2]], len(t[3])
        elif svtype == "INV":
            cols["c"][i] = 0 if t[0] == "++" else 1
            cols["a"][i], cols["b"][i], cols["read_id"][i] = int(t[1]), int(t[2]), name_id[t[3]]
        elif svtype == "TRA":
            cols["a"][i], cols["b"][i], cols["read_id"][i] = int(t[1]), int(t[3]), name_id[t[4]]
            cols["c"][i] = chrom_id[t[2]] * 4 + _TRA[t[0]]
    return cols


def reads_to_columns(rows, chrom_id, name_id):
    n = len(rows)
    out = dict(chrom=np.zeros(n, np.int32), start=np.zeros(n, np.int32), end=np.zeros(n, np.int32), read_id=np.zeros(n, np.int32),
               is_primary=np.zeros(n, np.uint8))
    for i, r in enumerate(rows):
        out["start"][i], out["end"][i], out["is_primary"][i] = r[0], r[1], r[2]
        out["read_id"][i] = name_id[r[3]]
        out["chrom"][i] = chrom_id[r[4]]
    return out
