"""ctypes binding of the test-only pipeline emulator (tests/emul/emul.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

from cutesv_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcutesv_emul.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(_HERE, "emul.cpp"), os.path.join(_HERE, "../../cutesv_b200/csrc/core.h"),
            os.path.join(_HERE, "../../cutesv_b200/csrc/host_tables.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                               "-o", _SO, srcs[0]])
    return _SO


def cluster(params, lens, sigs, reads, type_mask=0x1F, aln=None):
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.emul_cluster.restype = C.c_int
    arr = (_abi.csv_sig_cols * _abi.CSV_NTYPES)()
    keep = []
    total = 0
    for t, name in enumerate(_abi.TYPE_NAMES):
        s, k = _abi.make_sig_cols(sigs.get(name))
        arr[t] = s
        keep.append(k)
        total += s.n
    rc_struct, rk = _abi.make_reads_cols(reads)
    aln_struct, ak = _abi.make_reads_cols(aln)
    lens = np.ascontiguousarray(lens, dtype=np.int64)
    cap = max(2 * total, 16)
    cands = np.zeros(cap, dtype=_abi.CAND_DTYPE)
    genos = np.zeros(cap, dtype=_abi.GENO_DTYPE)
    names = np.zeros(cap, dtype=np.int32)
    nc, nn = C.c_int64(0), C.c_int64(0)
    rc = _lib.emul_cluster(C.byref(params), C.c_int32(len(lens)), lens.ctypes.data_as(C.POINTER(C.c_int64)), arr,
                           C.byref(rc_struct), C.c_uint32(type_mask), cands.ctypes.data_as(C.c_void_p),
                           genos.ctypes.data_as(C.c_void_p), C.c_int64(cap), _abi.ptr(names), C.c_int64(cap),
                           C.byref(nc), C.byref(nn), C.byref(aln_struct))
    if rc != 0:
        raise RuntimeError("emulator failed: %d" % rc)
    return cands[:nc.value].copy(), genos[:nc.value].copy(), names[:nn.value].copy()


def extract(params, packed):
    """Extraction emulator on a cutesv_b200.packing.pack_alignments() packet.
    Returns dict(sigs={type: cols}, piece_off, piece_cnt, pieces, rows)."""
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.emul_cluster.restype = C.c_int
    n = len(packed["chrom"])
    rc_ = _abi.csv_read_cols(n, *[_abi.ptr(packed[k]) for k in ("chrom", "ref_start", "ref_end", "flag", "mapq", "query_len", "read_id")],
                             packed["cigar_off"].ctypes.data_as(C.POINTER(C.c_int64)), packed["sa_off"].ctypes.data_as(C.POINTER(C.c_int64)))
    sa = packed["sa"]
    sa_ = _abi.csv_sa_cols(len(sa["chrom"]), *[_abi.ptr(sa[k]) for k in ("chrom", "pos0", "strand", "mapq", "first_clip", "last_clip", "ref_span")])
    cap = 4 * len(packed["cigar"]) + 64 * n + 1024
    cols = np.zeros((5, 5, cap), dtype=np.int32)
    poff = np.zeros(cap, dtype=np.int32)
    pcnt = np.zeros(cap, dtype=np.int32)
    pieces = np.zeros((2 * cap, 4), dtype=np.int32)
    rr = np.zeros((4, n + 1), dtype=np.int32)
    rp = np.zeros(n + 1, dtype=np.uint8)
    counts = (C.c_int64 * 8)()
    cig = np.ascontiguousarray(packed["cigar"], dtype=np.uint32)
    r = _lib.emul_extract(C.byref(params), C.byref(rc_), cig.ctypes.data_as(C.POINTER(C.c_uint32)), C.byref(sa_), _abi.ptr(cols), C.c_int64(cap),
                          _abi.ptr(poff), _abi.ptr(pcnt), _abi.ptr(pieces), C.c_int64(2 * cap), _abi.ptr(rr), rp.ctypes.data_as(C.POINTER(C.c_uint8)),
                          C.c_int64(n + 1), counts)
    assert r == 0 and (counts[7] & ~1024) == 0, (r, counts[7])   # 1024 = ST_SKIPPED: a record with more than 64 segments, not fatal
    sigs = {}
    for t, name in enumerate(_abi.TYPE_NAMES):
        k = counts[t]
        sigs[name] = dict(chrom=cols[t, 0, :k].copy(), a=cols[t, 1, :k].copy(), b=cols[t, 2, :k].copy(), read_id=cols[t, 3, :k].copy(),
                          c=cols[t, 4, :k].copy())
    nr = counts[6]
    rows = dict(chrom=rr[0, :nr].copy(), start=rr[1, :nr].copy(), end=rr[2, :nr].copy(), read_id=rr[3, :nr].copy(), is_primary=rp[:nr].copy())
    return dict(sigs=sigs, piece_off=poff[:counts[1]].copy(), piece_cnt=pcnt[:counts[1]].copy(), pieces=pieces[:counts[5]].copy(), rows=rows)
