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


def cluster(params, lens, sigs, reads, type_mask=0x1F):
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
    lens = np.ascontiguousarray(lens, dtype=np.int64)
    cap = max(total, 16)
    cands = np.zeros(cap, dtype=_abi.CAND_DTYPE)
    genos = np.zeros(cap, dtype=_abi.GENO_DTYPE)
    names = np.zeros(cap, dtype=np.int32)
    nc, nn = C.c_int64(0), C.c_int64(0)
    rc = _lib.emul_cluster(C.byref(params), C.c_int32(len(lens)), lens.ctypes.data_as(C.POINTER(C.c_int64)), arr,
                           C.byref(rc_struct), C.c_uint32(type_mask), cands.ctypes.data_as(C.c_void_p),
                           genos.ctypes.data_as(C.c_void_p), C.c_int64(cap), _abi.ptr(names), C.c_int64(cap),
                           C.byref(nc), C.byref(nn))
    if rc != 0:
        raise RuntimeError("emulator failed: %d" % rc)
    return cands[:nc.value].copy(), genos[:nc.value].copy(), names[:nn.value].copy()
