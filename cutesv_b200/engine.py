"""Engine: one csv_ctx (one GPU) driven from Python.  Host side of the drop-in boundary."""
import ctypes as C

import numpy as np

from . import _abi, _lib


class Engine(object):
    """Owns a csv_ctx.  stream: optional cudaStream_t handle (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device=0, stream=None, params=None, contig_lens=None):
        self.L = _lib.lib()
        h = C.c_void_p()
        _lib.check(self.L.csv_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.h = h
        self.device = int(device)
        self.params = None
        self._keep = None
        self.set_params(params if params is not None else _abi.default_params())
        if contig_lens is not None:
            self.set_contigs(contig_lens)

    def close(self):
        if getattr(self, "h", None):
            self.L.csv_destroy(self.h)
            self.h = None
        for p in getattr(self, "_pinned", []):
            self.L.csv_host_free(p)
        self._pinned = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params):
        self.params = params
        _lib.check(self.L.csv_set_params(self.h, C.byref(params)))

    def set_contigs(self, lens):
        lens = np.ascontiguousarray(lens, dtype=np.int64)
        self.n_contigs = len(lens)
        _lib.check(self.L.csv_set_contigs(self.h, C.c_int32(len(lens)), lens.ctypes.data_as(C.POINTER(C.c_int64))))

    def set_profiling(self, on):
        _lib.check(self.L.csv_set_profiling(self.h, int(bool(on))))

    def set_lanes(self, on):
        """Per-SV-type stream lanes (default on); off serialises the types on the ctx stream."""
        _lib.check(self.L.csv_set_lanes(self.h, int(bool(on))))

    # -- device-resident path (bench `value`): upload once, cluster many times --

    def upload(self, sigs, reads, grouped=False):
        """Asynchronous H2D of the inputs of cluster_device().  grouped=True: dicts from _abi.group_by_contig (rows grouped by
        contig + `contig_off` instead of the contig column; csv_upload_*_grouped)."""
        keep = []
        for t, name in enumerate(_abi.TYPE_NAMES):
            if grouped:
                s, off, k = _abi.make_sig_cols_grouped(sigs.get(name))
                keep.append((k, off))
                if off is not None:
                    _lib.check(self.L.csv_upload_sigs_grouped(self.h, t, C.byref(s), off.ctypes.data_as(C.POINTER(C.c_int64))))
                else:
                    _lib.check(self.L.csv_upload_sigs(self.h, t, C.byref(s)))
            else:
                s, k = _abi.make_sig_cols(sigs.get(name))
                keep.append(k)
                _lib.check(self.L.csv_upload_sigs(self.h, t, C.byref(s)))
        if grouped:
            r, r_off, rk = _abi.make_reads_cols_grouped(reads)
            keep.append((rk, r_off))
            if r_off is not None:
                _lib.check(self.L.csv_upload_reads_grouped(self.h, C.byref(r), r_off.ctypes.data_as(C.POINTER(C.c_int64))))
            else:
                _lib.check(self.L.csv_upload_reads(self.h, C.byref(r)))
        else:
            r, rk = _abi.make_reads_cols(reads)
            keep.append(rk)
            _lib.check(self.L.csv_upload_reads(self.h, C.byref(r)))
        self._keep = keep  # host buffers must outlive the async copies

    def upload_alignments(self, aln):
        """ALL alignment records in BAM order (dict like the reads table, is_primary = flag in (0, 16)): enables the
        device TRA genotyper (call_gt, resolveTRA.py:260-309).  None / empty clears the table."""
        r, keep = _abi.make_reads_cols(aln)
        _lib.check(self.L.csv_upload_alignments(self.h, C.byref(r)))
        self._keep_aln = keep

    def cluster_device(self, type_mask=0x1F):
        _lib.check(self.L.csv_cluster(self.h, C.c_uint32(type_mask)))

    def counts(self):
        nc, nn = C.c_int64(0), C.c_int64(0)
        _lib.check(self.L.csv_result_counts(self.h, C.byref(nc), C.byref(nn)))
        return nc.value, nn.value

    def fetch(self):
        nc, nn = self.counts()
        cands = np.zeros(max(nc, 1), dtype=_abi.CAND_DTYPE)
        genos = np.zeros(max(nc, 1), dtype=_abi.GENO_DTYPE)
        names = np.zeros(max(nn, 1), dtype=np.int32)
        _lib.check(self.L.csv_fetch(self.h, cands.ctypes.data_as(C.c_void_p), genos.ctypes.data_as(C.c_void_p), C.c_int64(len(cands)),
                                    _abi.ptr(names), C.c_int64(len(names))))
        return cands[:nc], genos[:nc], names[:nn]

    # -- the reference-facing one-shot call: host columns in, host rows out --
    def cluster(self, sigs, reads, type_mask=0x1F, out=None, grouped=False):
        """sigs: {type_name: dict(chrom,a,b,read_id[,c])}; reads: dict(chrom,start,end,read_id,is_primary).
        grouped=True: every dict holds rows grouped by contig and `contig_off` instead of `chrom`
        (_abi.group_by_contig; csv_cluster_host_grouped).
        Returns (cands, genos, names) numpy arrays in the reference's emission order."""
        arr = (_abi.csv_sig_cols * _abi.CSV_NTYPES)()
        offs = (C.POINTER(C.c_int64) * _abi.CSV_NTYPES)()
        keep = []
        total = 0
        for t, name in enumerate(_abi.TYPE_NAMES):
            if grouped:
                s, off, k = _abi.make_sig_cols_grouped(sigs.get(name))
                if off is not None:
                    offs[t] = off.ctypes.data_as(C.POINTER(C.c_int64))
            else:
                s, k = _abi.make_sig_cols(sigs.get(name))
            arr[t] = s
            keep.append(k)
            total += s.n
        if grouped:
            r, r_off, rk = _abi.make_reads_cols_grouped(reads)
        else:
            r, rk = _abi.make_reads_cols(reads)
        if out is None:
            cap_c = max(2 * (total // max(min(self.params.min_support_allele, self.params.min_support), 1)) + 16, 16)
            cap_n = total + 16
            cands = np.zeros(cap_c, dtype=_abi.CAND_DTYPE)
            genos = np.zeros(cap_c, dtype=_abi.GENO_DTYPE)
            names = np.zeros(cap_n, dtype=np.int32)
        else:
            cands, genos, names = out
        nc, nn = C.c_int64(0), C.c_int64(0)
        tail = (C.c_uint32(type_mask), cands.ctypes.data_as(C.c_void_p), genos.ctypes.data_as(C.c_void_p), C.c_int64(len(cands)),
                _abi.ptr(names), C.c_int64(len(names)), C.byref(nc), C.byref(nn))
        if grouped:
            _lib.check(self.L.csv_cluster_host_grouped(self.h, arr, offs, C.byref(r),
                                                       None if r_off is None else r_off.ctypes.data_as(C.POINTER(C.c_int64)), *tail))
        else:
            _lib.check(self.L.csv_cluster_host(self.h, arr, C.byref(r), *tail))
        return cands[:nc.value], genos[:nc.value], names[:nn.value]

    def cal_gl(self, c0, c1):
        c0 = np.ascontiguousarray(c0, dtype=np.int32)
        c1 = np.ascontiguousarray(c1, dtype=np.int32)
        out = np.zeros(len(c0), dtype=_abi.GENO_DTYPE)
        _lib.check(self.L.csv_cal_gl(self.h, _abi.ptr(c0), _abi.ptr(c1), C.c_int64(len(c0)), out.ctypes.data_as(C.c_void_p)))
        return out

    def stage_ms(self):
        ms = (C.c_float * _abi.CSV_ST_COUNT)()
        _lib.check(self.L.csv_stage_ms(self.h, ms))
        return {name: float(ms[i]) for i, name in enumerate(_abi.STAGES)}

    def sort_probe(self):
        ms, b, n = C.c_float(0), C.c_int64(0), C.c_int32(0)
        _lib.check(self.L.csv_sort_probe(self.h, C.byref(ms), C.byref(b), C.byref(n)))
        return dict(ms=float(ms.value), bytes=int(b.value), launches=int(n.value))

    def counters(self):
        out = (C.c_uint32 * 32)()
        _lib.check(self.L.csv_debug_counters(self.h, out))
        v = list(out)
        t = _abi.TYPE_NAMES
        return dict(status=v[0], n_cand=v[1], n_names=v[2], max_support=v[3], kept=dict(zip(t, v[4:9])), big=dict(zip(t, v[9:14])),
                    giant=dict(zip(t, v[14:19])), pairs=v[19], domain=dict(zip(t, v[20:25])), members=dict(zip(t, v[25:30])), small_path=v[30])

    def kernel_times(self):
        """{kernel name: (launches, total ms)} of the calls made while profiling was on (collected by fetch())."""
        need = int(self.L.csv_kernel_times(self.h, None, 0))
        buf = C.create_string_buffer(need + 16)
        self.L.csv_kernel_times(self.h, buf, need + 16)
        out = {}
        for ln in buf.value.decode().splitlines():
            nm, n, ms = ln.split("\t")
            out[nm] = (int(n), float(ms))
        return out

    def launch_count(self):
        return int(self.L.csv_launch_count(self.h))

    def graph_replays(self):
        return int(self.L.csv_graph_replays(self.h))

    # -- multi-GPU: contig shards + one NCCL all-gather of the final records --

    def set_shard(self, owned):
        """owned: bool/uint8 mask over contig ids (None = all contigs)."""
        if owned is None:
            _lib.check(self.L.csv_set_shard(self.h, None))
            return
        m = np.ascontiguousarray(owned, dtype=np.uint8)
        assert len(m) == self.n_contigs
        _lib.check(self.L.csv_set_shard(self.h, m.ctypes.data_as(C.POINTER(C.c_uint8))))

    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        _lib.check(self.L.csv_comm_unique_id(buf, 128))
        return buf.raw

    def comm_init(self, uid, rank, world):
        buf = C.create_string_buffer(bytes(uid), 128)
        _lib.check(self.L.csv_comm_init(self.h, buf, int(rank), int(world)))
        self.rank, self.world = int(rank), int(world)

    def comm_destroy(self):
        _lib.check(self.L.csv_comm_destroy(self.h))

    def allgather(self):
        """Asynchronous, collective: pack + ONE ncclAllGather + device merge of the records of the last cluster_device()."""
        _lib.check(self.L.csv_allgather(self.h))

    def set_gather(self, peer_to_peer):
        """True: store into the peers' mail boxes over NVLink (CUDA IPC); False: ncclAllGather."""
        _lib.check(self.L.csv_set_gather(self.h, 1 if peer_to_peer else 0))

    def gather_mode(self):
        return "peer-to-peer" if self.L.csv_gather_mode(self.h) else "nccl"

    def gathered_counts(self):
        nc, nn = C.c_int64(0), C.c_int64(0)
        _lib.check(self.L.csv_gathered_counts(self.h, C.byref(nc), C.byref(nn)))
        return nc.value, nn.value

    def fetch_gathered(self, out=None):
        nc, nn = self.gathered_counts()
        if out is not None:
            cands, genos, names = out
        else:
            cands = np.zeros(max(nc, 1), dtype=_abi.CAND_DTYPE)
            genos = np.zeros(max(nc, 1), dtype=_abi.GENO_DTYPE)
            names = np.zeros(max(nn, 1), dtype=np.int32)
        _lib.check(self.L.csv_fetch_gathered(self.h, cands.ctypes.data_as(C.c_void_p), genos.ctypes.data_as(C.c_void_p), C.c_int64(len(cands)),
                                             _abi.ptr(names), C.c_int64(len(names))))
        return cands[:nc], genos[:nc], names[:nn]

    def device_ptrs(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(self.L.csv_result_device_ptrs(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value


def _pin_packet_method(self, packed):
    """Copy of an alignment packet in page-locked host memory (csv_host_alloc), so that csv_extract's H2D copies run at PCIe
    speed and asynchronously.  The buffers are freed by close()."""
    if not hasattr(self, "_pinned"):
        self._pinned = []

    def pin(a):
        a = np.ascontiguousarray(a)
        p = C.c_void_p()
        _lib.check(self.L.csv_host_alloc(C.byref(p), C.c_size_t(max(a.nbytes, 1))))
        self._pinned.append(p)
        buf = (C.c_char * max(a.nbytes, 1)).from_address(p.value)
        out = np.frombuffer(buf, dtype=a.dtype, count=a.size).reshape(a.shape)
        out[...] = a
        return out

    out = {}
    for k, v in packed.items():
        if k == "sa":
            out[k] = {kk: pin(np.asarray(vv, dtype=np.int32)) for kk, vv in v.items()}
        elif k in ("cigar_off", "sa_off"):
            out[k] = pin(np.asarray(v, dtype=np.int64))
        elif k == "cigar":
            out[k] = pin(np.asarray(v, dtype=np.uint32))
        elif isinstance(v, np.ndarray) and v.dtype.kind in "iu" and k in ("chrom", "ref_start", "ref_end", "flag", "mapq", "query_len", "read_id"):
            out[k] = pin(np.asarray(v, dtype=np.int32))
        else:
            out[k] = v
    return out


def _extract_method(self, packed, append=False):
    """csv_extract on a packing.pack_alignments() packet.  The extracted signatures and reads rows
    stay device-resident as the inputs of cluster_device(); returns dict(counts, n_rows).
    append=True (csv_extract_append): this packet's output is appended to what earlier packets left on the device;
    counts / n_rows are the totals so far and `first` holds the totals before this packet."""
    n = len(packed["chrom"])
    keep = [np.ascontiguousarray(packed[k], dtype=np.int32) for k in ("chrom", "ref_start", "ref_end", "flag", "mapq", "query_len", "read_id")]
    co = np.ascontiguousarray(packed["cigar_off"], dtype=np.int64)
    so = np.ascontiguousarray(packed["sa_off"], dtype=np.int64)
    rc_ = _abi.csv_read_cols(n, *[_abi.ptr(k) for k in keep], co.ctypes.data_as(C.POINTER(C.c_int64)), so.ctypes.data_as(C.POINTER(C.c_int64)))
    sa = {k: np.ascontiguousarray(v, dtype=np.int32) for k, v in packed["sa"].items()}
    sa_ = _abi.csv_sa_cols(len(sa["chrom"]), *[_abi.ptr(sa[k]) for k in ("chrom", "pos0", "strand", "mapq", "first_clip", "last_clip", "ref_span")])
    cig = np.ascontiguousarray(packed["cigar"], dtype=np.uint32)
    counts = (C.c_int64 * _abi.CSV_NTYPES)()
    n_rows = C.c_int64(0)
    first = list(getattr(self, "_ex_counts", [0] * _abi.CSV_NTYPES)) if (append and getattr(self, "_ex_appending", False)) else [0] * _abi.CSV_NTYPES
    first_rows = getattr(self, "_ex_rows", 0) if (append and getattr(self, "_ex_appending", False)) else 0
    first_pieces = getattr(self, "_ex_pieces", 0) if (append and getattr(self, "_ex_appending", False)) else 0
    fn = self.L.csv_extract_append if append else self.L.csv_extract
    _lib.check(fn(self.h, C.byref(rc_), cig.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_int64(len(cig)), C.byref(sa_), counts, C.byref(n_rows)))
    self._ex_counts = [int(x) for x in counts]
    self._ex_rows = int(n_rows.value)
    self._ex_appending = bool(append)
    npz = C.c_int64(0)
    _lib.check(self.L.csv_fetch_pieces(self.h, C.c_int64(0), None, C.byref(npz)))
    self._ex_pieces = int(npz.value)
    return dict(counts={name: int(counts[t]) for t, name in enumerate(_abi.TYPE_NAMES)}, n_rows=int(n_rows.value),
                first={name: first[t] for t, name in enumerate(_abi.TYPE_NAMES)}, first_rows=first_rows, first_pieces=first_pieces,
                n_pieces=self._ex_pieces)


def _extract_skipped_method(self):
    """Records whose split-read analysis was skipped (more than 64 qualifying segments, only with max_split_parts -1)."""
    return int(self.L.csv_extract_skipped(self.h))


def _extract_reset_method(self):
    _lib.check(self.L.csv_extract_reset(self.h))
    self._ex_counts = [0] * _abi.CSV_NTYPES
    self._ex_rows = 0
    self._ex_pieces = 0
    self._ex_appending = False


def _fetch_ins_pieces_method(self, first_sig, n_sig, first_piece, n_piece):
    """Piece descriptors of INS signatures [first_sig, first_sig + n_sig) and pieces [first_piece, first_piece + n_piece):
    what the host needs to rebuild the sequences of the rows ONE packet appended.  piece_off is re-based to the slice."""
    po = np.zeros(max(n_sig, 1), dtype=np.int32)
    pc = np.zeros(max(n_sig, 1), dtype=np.int32)
    if n_sig:
        _lib.check(self.L.csv_fetch_sigs_range(self.h, _abi.CSV_INS, C.c_int64(first_sig), C.c_int64(n_sig), None, None, None, None, None,
                                               _abi.ptr(po), _abi.ptr(pc)))
    pieces = np.zeros((max(n_piece, 1), 4), dtype=np.int32)
    if n_piece:
        _lib.check(self.L.csv_fetch_pieces_range(self.h, C.c_int64(first_piece), C.c_int64(n_piece), _abi.ptr(pieces)))
    return po[:n_sig] - first_piece, pc[:n_sig], pieces[:n_piece]


def _fetch_sig_cols_method(self, name, cols=("chrom", "a", "b", "read_id", "c")):
    """D2H of whole columns of the device-resident signatures of one type."""
    t = _abi.TYPE_IDS[name]
    k = self._ex_counts[t]
    out = {c: (np.zeros(max(k, 1), dtype=np.int32) if c in cols else None) for c in ("chrom", "a", "b", "read_id", "c")}
    if k:
        _lib.check(self.L.csv_fetch_sigs_range(self.h, t, C.c_int64(0), C.c_int64(k), *[(_abi.ptr(out[c]) if out[c] is not None else None)
                                                                                          for c in ("chrom", "a", "b", "read_id", "c")], None, None))
    return {c: (v[:k] if v is not None else None) for c, v in out.items()}


def _fetch_read_rows_method(self):
    """D2H of the device-resident reads table (reads_info_list rows, cuteSV:729-733)."""
    nr = self._ex_rows
    rows = {k: np.zeros(max(nr, 1), dtype=np.int32) for k in ("chrom", "start", "end", "read_id")}
    prim = np.zeros(max(nr, 1), dtype=np.uint8)
    _lib.check(self.L.csv_fetch_read_rows(self.h, C.c_int64(max(nr, 1)), _abi.ptr(rows["chrom"]), _abi.ptr(rows["start"]), _abi.ptr(rows["end"]),
                                          _abi.ptr(rows["read_id"]), prim.ctypes.data_as(C.POINTER(C.c_uint8))))
    rows = {k: v[:nr] for k, v in rows.items()}
    rows["is_primary"] = prim[:nr]
    return rows


def _remap_read_ids_method(self, rank):
    rank = np.ascontiguousarray(rank, dtype=np.int32)
    _lib.check(self.L.csv_remap_read_ids(self.h, _abi.ptr(rank), C.c_int64(len(rank))))


def _swap_ins_rows_method(self, pairs):
    pairs = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
    if len(pairs):
        _lib.check(self.L.csv_swap_ins_rows(self.h, pairs.ctypes.data_as(C.POINTER(C.c_int64)), C.c_int64(len(pairs))))


def _fetch_extracted_method(self):
    """D2H of everything csv_extract produced (parity tests / host ALT strings)."""
    sigs = {}
    poff = pcnt = None
    for t, name in enumerate(_abi.TYPE_NAMES):
        k = self._ex_counts[t]
        cols = {c: np.zeros(max(k, 1), dtype=np.int32) for c in ("chrom", "a", "b", "read_id", "c")}
        po = np.zeros(max(k, 1), dtype=np.int32)
        pc = np.zeros(max(k, 1), dtype=np.int32)
        _lib.check(self.L.csv_fetch_sigs(self.h, t, C.c_int64(max(k, 1)), _abi.ptr(cols["chrom"]), _abi.ptr(cols["a"]), _abi.ptr(cols["b"]),
                                         _abi.ptr(cols["read_id"]), _abi.ptr(cols["c"]), _abi.ptr(po), _abi.ptr(pc)))
        sigs[name] = {c: v[:k] for c, v in cols.items()}
        if name == "INS":
            poff, pcnt = po[:k], pc[:k]
    npz = C.c_int64(0)
    _lib.check(self.L.csv_fetch_pieces(self.h, C.c_int64(0), None, C.byref(npz)))
    pieces = np.zeros((max(npz.value, 1), 4), dtype=np.int32)
    _lib.check(self.L.csv_fetch_pieces(self.h, C.c_int64(len(pieces)), _abi.ptr(pieces), C.byref(npz)))
    nr = self._ex_rows
    rows = {k: np.zeros(max(nr, 1), dtype=np.int32) for k in ("chrom", "start", "end", "read_id")}
    prim = np.zeros(max(nr, 1), dtype=np.uint8)
    _lib.check(self.L.csv_fetch_read_rows(self.h, C.c_int64(max(nr, 1)), _abi.ptr(rows["chrom"]), _abi.ptr(rows["start"]), _abi.ptr(rows["end"]),
                                          _abi.ptr(rows["read_id"]), prim.ctypes.data_as(C.POINTER(C.c_uint8))))
    rows = {k: v[:nr] for k, v in rows.items()}
    rows["is_primary"] = prim[:nr]
    return dict(sigs=sigs, piece_off=poff, piece_cnt=pcnt, pieces=pieces[:npz.value], rows=rows)


Engine.pin_packet = _pin_packet_method
Engine.extract_reset = _extract_reset_method
Engine.extract_skipped = _extract_skipped_method
Engine.fetch_ins_pieces = _fetch_ins_pieces_method
Engine.fetch_sig_cols = _fetch_sig_cols_method
Engine.remap_read_ids = _remap_read_ids_method
Engine.fetch_read_rows = _fetch_read_rows_method
Engine.swap_ins_rows = _swap_ins_rows_method
Engine.extract = _extract_method
Engine.fetch_extracted = _fetch_extracted_method
