"""ctypes binding of oracle/libcutesv_oracle.so (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under cutesv_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from cutesv_b200 import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcutesv_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "cutesv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.csvo_cluster.restype = C.c_int
        _lib.csvo_cal_gl.restype = None
        _lib.csvo_pow_half.restype = C.c_double
        _lib.csvo_pow_half.argtypes = [C.c_int]
        _lib.csvo_np_std_i32.restype = C.c_double
        _lib.csvo_np_std_i32.argtypes = [C.POINTER(C.c_int32), C.c_int64]
        _lib.csvo_cal_cipos.restype = C.c_int
        _lib.csvo_cal_cipos.argtypes = [C.c_double, C.c_int]
    return _lib


def cal_gl(c0, c1):
    g = np.zeros(1, dtype=_abi.GENO_DTYPE)
    lib().csvo_cal_gl(C.c_int(int(c0)), C.c_int(int(c1)), g.ctypes.data_as(C.c_void_p))
    return g[0]


def np_std(v):
    v = np.ascontiguousarray(v, dtype=np.int32)
    return lib().csvo_np_std_i32(_abi.ptr(v), len(v))


def cluster(params, lens, sigs, reads, type_mask=0x1F, n_threads=1, aln=None):
    """sigs: {type_name: cols}.  Returns (cands, genos, names) numpy arrays."""
    L = lib()
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
    cap_c, cap_n = max(2 * (total // max(min(params.min_support_allele, params.min_support), 1)), 16), max(total, 16)
    while True:
        cands = np.zeros(cap_c, dtype=_abi.CAND_DTYPE)
        genos = np.zeros(cap_c, dtype=_abi.GENO_DTYPE)
        names = np.zeros(cap_n, dtype=np.int32)
        nc, nn = C.c_int64(0), C.c_int64(0)
        rc = L.csvo_cluster(C.byref(params), C.c_int32(len(lens)), lens.ctypes.data_as(C.POINTER(C.c_int64)),
                            arr, C.byref(rc_struct), C.c_uint32(type_mask),
                            cands.ctypes.data_as(C.c_void_p), genos.ctypes.data_as(C.c_void_p), C.c_int64(cap_c),
                            _abi.ptr(names), C.c_int64(cap_n), C.byref(nc), C.byref(nn), C.c_int(n_threads), C.byref(aln_struct))
        if rc == _abi.CSV_E_CAPACITY:
            cap_c, cap_n = max(nc.value, 16), max(nn.value, 16)
            continue
        if rc != 0:
            raise RuntimeError("oracle failed: %d" % rc)
        return cands[:nc.value].copy(), genos[:nc.value].copy(), names[:nn.value].copy()
