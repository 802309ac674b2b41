"""Builds libcutesv_b200.so in-tree with nvcc for sm_100a and the host-only BAM decoder libcutesv_bam.so
(no JIT cache: the .so files travel with the repo)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "cutesv_b200.cu")
SO = os.path.join(HERE, "libcutesv_b200.so")
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared"]


def sources():
    d = os.path.join(HERE, "csrc")
    out = [os.path.join(d, f) for f in sorted(os.listdir(d)) if f != "bam_reader.cpp"]  # host-only library, see bamio.py
    out.append(os.path.join(HERE, "..", "include", "cutesv_b200.h"))
    return out


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    from . import bamio
    bamio.build(force=force)   # libcutesv_bam.so (g++, zlib): the native BAM decoder
    if not force and not needs_build():
        return SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO, SRC]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    print(build(force=True, verbose=True))
