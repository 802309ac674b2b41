"""CPU: the C-ABI library loads and exports every symbol include/cutesv_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

from cutesv_b200 import _abi, _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "cutesv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(csv_[a-z_0-9]+)\s*\(", src)))


def test_library_is_built_and_exports_header():
    if build.needs_build():
        build.build()
    L = ctypes.CDLL(_lib.so_path())
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "missing export %s" % s
    assert set(syms) == set(_lib.EXPORTS), set(syms) ^ set(_lib.EXPORTS)


def test_struct_layouts():
    assert ctypes.sizeof(_abi.csv_params) == 112
    assert _abi.CAND_DTYPE.itemsize == 64 and _abi.GENO_DTYPE.itemsize == 40
    assert ctypes.sizeof(_abi.csv_sig_cols) == 48 and ctypes.sizeof(_abi.csv_reads_cols) == 48


def test_no_cpu_fallback_without_gpu():
    """Without a usable device csv_create must fail loudly (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    h = ctypes.c_void_p()
    rc = L.csv_create(0, None, ctypes.byref(h))
    assert rc == _abi.CSV_E_NODEVICE
    assert b"no CPU fallback" in L.csv_last_error() or b"CUDA" in L.csv_last_error()


def test_product_does_not_import_oracle():
    """Nothing under cutesv_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "cutesv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inl")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
                assert "emul_lib" not in txt, f


def test_group_by_contig_is_stable_and_complete():
    """Grouped host inputs (csv_*_grouped): stable regrouping by contig + row offsets."""
    import numpy as np
    from cutesv_b200 import _abi, synth
    cfg = synth.adversarial(5)
    nc = len(cfg["lens"])
    for cols in list(cfg["sigs"].values()) + [cfg["reads"]]:
        g = _abi.group_by_contig(cols, nc)
        off = g["contig_off"]
        assert off[0] == 0 and off[-1] == len(cols["chrom"]) and np.all(np.diff(off) >= 0) and "chrom" not in g
        order = np.argsort(cols["chrom"], kind="stable")
        for k, v in cols.items():
            if k == "chrom" or v is None:
                continue
            assert np.array_equal(g[k], np.asarray(v)[order]), k
        for c in range(nc):   # rows of contig c, in their original relative order
            assert np.all(np.asarray(cols["chrom"])[order][off[c]:off[c + 1]] == c)
