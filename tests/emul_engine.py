"""Test-only stand-in for cutesv_b200.engine.Engine built on the CPU pipeline emulator (tests/emul):
lets the CLI shell (host code: BAM decoding, packet filtering, id ranking, rows, VCF) run end to end
without a GPU.  The emulator instantiates the same per-cluster templates (csrc/core.h,
csrc/extract_core.h) the kernels do; it is never used by the product.

Implements the part of the Engine interface the CLI drives: append-mode extraction with the "device-resident"
signature columns kept in numpy arrays, id remapping, INS row swaps, cluster_device / fetch."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul"))
import emul_lib  # noqa: E402

from cutesv_b200 import _abi  # noqa: E402

_COLS = ("chrom", "a", "b", "read_id", "c")
_ROWS = ("chrom", "start", "end", "read_id", "is_primary")


class EmulEngine(object):
    def __init__(self):
        self.params = None
        self.lens = None
        self.aln = None
        self._ex = None
        self.extract_reset()

    def set_params(self, params):
        self.params = params

    def set_contigs(self, lens):
        self.lens = np.ascontiguousarray(lens, dtype=np.int64)

    def extract_reset(self):
        self.sigs = {t: {k: np.zeros(0, np.int32) for k in _COLS} for t in _abi.TYPE_NAMES}
        self.rows = {k: np.zeros(0, np.uint8 if k == "is_primary" else np.int32) for k in _ROWS}
        self.piece_off = np.zeros(0, np.int32)
        self.piece_cnt = np.zeros(0, np.int32)
        self.pieces = np.zeros((0, 4), np.int32)
        self.n_records = 0
        self._res = None

    def extract(self, packed, append=False):
        if not append:
            self.extract_reset()
        pk = {k: (np.ascontiguousarray(v) if isinstance(v, np.ndarray) else v) for k, v in packed.items()}
        pk["sa"] = {k: np.ascontiguousarray(v, dtype=np.int32) for k, v in packed["sa"].items()}
        ex = emul_lib.extract(self.params, pk)
        self._ex = ex
        first = {t: len(self.sigs[t]["chrom"]) for t in _abi.TYPE_NAMES}
        first_pieces, first_rows = len(self.pieces), len(self.rows["chrom"])
        for t in _abi.TYPE_NAMES:
            for k in _COLS:
                self.sigs[t][k] = np.concatenate([self.sigs[t][k], ex["sigs"][t][k]])
        for k in _ROWS:
            self.rows[k] = np.concatenate([self.rows[k], ex["rows"][k]])
        pieces = ex["pieces"].copy()
        if len(pieces):
            pieces[:, 0] += self.n_records
        self.piece_off = np.concatenate([self.piece_off, ex["piece_off"] + first_pieces])
        self.piece_cnt = np.concatenate([self.piece_cnt, ex["piece_cnt"]])
        self.pieces = np.concatenate([self.pieces, pieces]) if len(pieces) else self.pieces
        self.n_records += len(pk["chrom"])
        return dict(counts={t: len(self.sigs[t]["chrom"]) for t in _abi.TYPE_NAMES}, n_rows=len(self.rows["chrom"]), first=first,
                    first_rows=first_rows, first_pieces=first_pieces, n_pieces=len(self.pieces))

    def fetch_extracted(self):
        return self._ex

    def fetch_ins_pieces(self, first_sig, n_sig, first_piece, n_piece):
        return (self.piece_off[first_sig:first_sig + n_sig] - first_piece, self.piece_cnt[first_sig:first_sig + n_sig],
                self.pieces[first_piece:first_piece + n_piece])

    def fetch_sig_cols(self, name, cols=_COLS):
        return {k: (self.sigs[name][k].copy() if k in cols else None) for k in _COLS}

    def fetch_read_rows(self):
        return {k: v.copy() for k, v in self.rows.items()}

    def remap_read_ids(self, rank):
        rank = np.asarray(rank, dtype=np.int32)
        for t in _abi.TYPE_NAMES:
            if len(self.sigs[t]["read_id"]):
                self.sigs[t]["read_id"] = rank[self.sigs[t]["read_id"]]
        if len(self.rows["read_id"]):
            self.rows["read_id"] = rank[self.rows["read_id"]]

    def swap_ins_rows(self, pairs):
        for i, j in np.asarray(pairs, dtype=np.int64).reshape(-1, 2):
            for k in _COLS:
                a = self.sigs["INS"][k]
                a[i], a[j] = a[j], a[i]
            for a in (self.piece_off, self.piece_cnt):
                a[i], a[j] = a[j], a[i]

    def upload_alignments(self, aln):
        self.aln = aln

    def cluster_device(self, type_mask=0x1F):
        sigs = {t: dict(v, c=(None if t in ("DEL", "DUP") else v["c"])) for t, v in self.sigs.items()}
        self._res = emul_lib.cluster(self.params, self.lens, sigs, self.rows, type_mask, aln=self.aln)

    def fetch(self):
        return self._res

    def cluster(self, sigs, reads, type_mask=0x1F):
        return emul_lib.cluster(self.params, self.lens, sigs, reads, type_mask, aln=self.aln)
