"""ctypes / numpy mirrors of include/cutesv_b200.h (struct layouts only, no library loading).

Shared by the product binding (cutesv_b200/_lib.py) and by the test-only oracle binding
(the oracle wrapper under oracle/); the header is the single source of truth for the layouts.
"""
import ctypes as C

import numpy as np

CSV_DEL, CSV_INS, CSV_INV, CSV_DUP, CSV_TRA, CSV_NTYPES = 0, 1, 2, 3, 4, 5
TYPE_NAMES = ("DEL", "INS", "INV", "DUP", "TRA")
TYPE_IDS = {n: i for i, n in enumerate(TYPE_NAMES)}

CSV_OK, CSV_E_INVALID, CSV_E_CUDA, CSV_E_CAPACITY, CSV_E_NODEVICE, CSV_E_INPUT, CSV_E_STATE = (
    0, -1, -2, -3, -4, -5, -6)
CSV_F_NO_READS, CSV_F_GT_HOST = 1, 2

STAGES = ("h2d", "keys", "sort", "segment", "cluster", "order", "genotype", "d2h", "extract")
CSV_ST_COUNT = len(STAGES)


class csv_params(C.Structure):
    _fields_ = [
        ("min_support", C.c_int32), ("min_support_allele", C.c_int32),
        ("min_size", C.c_int32), ("max_size", C.c_int32),
        ("bias_del", C.c_int32), ("bias_ins", C.c_int32), ("bias_inv", C.c_int32),
        ("bias_dup", C.c_int32), ("bias_tra", C.c_int32),
        ("genotype", C.c_int32), ("gt_round", C.c_int32), ("gt_bias_ins", C.c_int32),
        ("ratio_del", C.c_double), ("ratio_ins", C.c_double), ("ratio_tra", C.c_double),
        ("remain_reads_ratio", C.c_double),
        ("min_mapq", C.c_int32), ("max_split_parts", C.c_int32), ("min_read_len", C.c_int32),
        ("min_siglength", C.c_int32), ("merge_del_threshold", C.c_int32),
        ("merge_ins_threshold", C.c_int32), ("reserved", C.c_int32 * 2),
    ]


_I32P = C.POINTER(C.c_int32)
_I64P = C.POINTER(C.c_int64)
_U8P = C.POINTER(C.c_uint8)
_U32P = C.POINTER(C.c_uint32)


class csv_sig_cols(C.Structure):
    _fields_ = [("n", C.c_int64), ("chrom", _I32P), ("a", _I32P), ("b", _I32P),
                ("read_id", _I32P), ("c", _I32P)]


class csv_reads_cols(C.Structure):
    _fields_ = [("n", C.c_int64), ("chrom", _I32P), ("start", _I32P), ("end", _I32P),
                ("read_id", _I32P), ("is_primary", _U8P)]


class csv_read_cols(C.Structure):
    _fields_ = [("n", C.c_int64), ("chrom", _I32P), ("ref_start", _I32P), ("ref_end", _I32P),
                ("flag", _I32P), ("mapq", _I32P), ("query_len", _I32P), ("read_id", _I32P),
                ("cigar_off", _I64P), ("sa_off", _I64P)]


class csv_sa_cols(C.Structure):
    _fields_ = [("n", C.c_int64), ("chrom", _I32P), ("pos0", _I32P), ("strand", _I32P),
                ("mapq", _I32P), ("first_clip", _I32P), ("last_clip", _I32P), ("ref_span", _I32P)]


CAND_DTYPE = np.dtype([
    ("svtype", "<i4"), ("chrom", "<i4"), ("pos", "<i4"), ("len", "<i4"), ("support", "<i4"),
    ("cipos", "<i4"), ("cilen", "<i4"), ("search_pos", "<i4"), ("pos2", "<i4"), ("aux", "<i4"),
    ("names_off", "<i4"), ("names_cnt", "<i4"), ("cluster", "<i4"), ("flags", "<i4"),
    ("reserved", "<i4", (2,)),
])
GENO_DTYPE = np.dtype([
    ("dr", "<i4"), ("dv", "<i4"), ("gt", "<i4"), ("pl", "<i4", (3,)), ("gq", "<i4"),
    ("status", "<i4"), ("qual", "<f8"),
])
assert CAND_DTYPE.itemsize == 64 and GENO_DTYPE.itemsize == 40


def i32(a):
    """Contiguous int32 view/copy of `a` (None passes through)."""
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a, ctype=C.c_int32):
    if a is None:
        return C.cast(None, C.POINTER(ctype))
    return a.ctypes.data_as(C.POINTER(ctype))


def make_sig_cols(cols):
    """cols: dict(chrom, a, b, read_id[, c]) of int32 arrays (or None) -> (struct, keepalive)."""
    if cols is None or len(cols["chrom"]) == 0:
        return csv_sig_cols(0, None, None, None, None, None), ()
    keep = tuple(i32(cols.get(k)) for k in ("chrom", "a", "b", "read_id", "c"))
    n = len(keep[0])
    for k in keep:
        assert k is None or len(k) == n
    return csv_sig_cols(n, *[ptr(k) for k in keep]), keep


def make_sig_cols_grouped(cols):
    """cols: dict(contig_off, a, b, read_id[, c]) (rows grouped by contig) -> (struct, offsets array, keepalive)."""
    if cols is None or len(cols["a"]) == 0:
        return csv_sig_cols(0, None, None, None, None, None), None, ()
    keep = tuple(i32(cols.get(k)) for k in ("a", "b", "read_id", "c"))
    off = np.ascontiguousarray(cols["contig_off"], dtype=np.int64)
    n = len(keep[0])
    for k in keep:
        assert k is None or len(k) == n
    return csv_sig_cols(n, None, *[ptr(k) for k in keep]), off, keep + (off,)


def make_reads_cols_grouped(reads):
    if reads is None or len(reads["start"]) == 0:
        return csv_reads_cols(0, None, None, None, None, None), None, ()
    keep = [i32(reads[k]) for k in ("start", "end", "read_id")]
    prim = np.ascontiguousarray(reads["is_primary"], dtype=np.uint8)
    off = np.ascontiguousarray(reads["contig_off"], dtype=np.int64)
    s = csv_reads_cols(len(keep[0]), None, *[ptr(k) for k in keep], ptr(prim, C.c_uint8))
    return s, off, tuple(keep) + (prim, off)


def group_by_contig(cols, n_contigs):
    """Stable regrouping of a column dict by its `chrom` column: returns the dict without `chrom`, with
    `contig_off` (n_contigs + 1 row offsets).  The order inside a contig is kept, so every tie-break on
    the input order gives the same result as for the ungrouped columns."""
    chrom = np.asarray(cols["chrom"])
    order = np.argsort(chrom, kind="stable")
    out = {k: (None if v is None else np.ascontiguousarray(np.asarray(v)[order])) for k, v in cols.items() if k != "chrom"}
    off = np.zeros(n_contigs + 1, dtype=np.int64)
    np.cumsum(np.bincount(chrom, minlength=n_contigs)[:n_contigs], out=off[1:])
    out["contig_off"] = off
    return out


def make_reads_cols(reads):
    """reads: dict(chrom, start, end, read_id, is_primary) -> (struct, keepalive)."""
    if reads is None or len(reads["chrom"]) == 0:
        return csv_reads_cols(0, None, None, None, None, None), ()
    keep = [i32(reads[k]) for k in ("chrom", "start", "end", "read_id")]
    prim = np.ascontiguousarray(reads["is_primary"], dtype=np.uint8)
    n = len(keep[0])
    s = csv_reads_cols(n, *[ptr(k) for k in keep], ptr(prim, C.c_uint8))
    return s, tuple(keep) + (prim,)


def default_params(**kw):
    """Reference defaults (cuteSV_Description.py:78-262; wiring cuteSV:1116-1189)."""
    p = csv_params()
    p.min_support = 10
    p.min_size = 30
    p.max_size = 100000
    p.bias_del, p.bias_ins, p.bias_inv, p.bias_dup, p.bias_tra = 200, 100, 500, 500, 50
    p.genotype = 0
    p.gt_round = 500
    p.gt_bias_ins = 1000
    p.ratio_del, p.ratio_ins, p.ratio_tra = 0.5, 0.3, 0.6
    p.remain_reads_ratio = 1.0
    p.min_mapq, p.max_split_parts, p.min_read_len, p.min_siglength = 20, 7, 500, 10
    p.merge_del_threshold, p.merge_ins_threshold = 0, 100
    explicit_allele = "min_support_allele" in kw
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError("unknown parameter %r" % k)
        setattr(p, k, v)
    if not explicit_allele:
        p.min_support_allele = min(p.min_support, 5)  # cuteSV:1124,1141
    return p
