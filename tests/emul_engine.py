"""Test-only stand-in for cutesv_b200.engine.Engine built on the CPU pipeline emulator (tests/emul):
lets the CLI shell (host code: BAM decoding, packet filtering, id ranking, rows, VCF) run end to end
without a GPU.  The emulator instantiates the same per-cluster templates (csrc/core.h,
csrc/extract_core.h) the kernels do; it is never used by the product."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import emul_lib  # noqa: E402


class EmulEngine(object):
    def __init__(self):
        self.params = None
        self.lens = None
        self.aln = None
        self._ex = None

    def set_params(self, params):
        self.params = params

    def set_contigs(self, lens):
        self.lens = np.ascontiguousarray(lens, dtype=np.int64)

    def extract(self, packed):
        pk = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in packed.items()}
        pk["sa"] = {k: np.ascontiguousarray(v, dtype=np.int32) for k, v in packed["sa"].items()}
        self._ex = emul_lib.extract(self.params, pk)

    def fetch_extracted(self):
        return self._ex

    def upload_alignments(self, aln):
        self.aln = aln

    def cluster(self, sigs, reads, type_mask=0x1F):
        return emul_lib.cluster(self.params, self.lens, sigs, reads, type_mask, aln=self.aln)
