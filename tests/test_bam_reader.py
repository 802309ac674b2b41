"""Native BAM decoder (cutesv_b200/bamio.py) against the per-read Python packer on the same records."""
import os

import numpy as np
import pytest

import bam_writer
from cutesv_b200 import bamio, packing, synth


@pytest.fixture(scope="module", autouse=True)
def _built():
    bamio.build()


def _sorted_reads(seed, n_reads, with_seq=True):
    reads, names, lens = synth.synth_alignments(seed, n_reads=n_reads, with_seq=with_seq)
    contigs = list(zip(names, (int(x) for x in lens)))
    order = {n: i for i, n in enumerate(names)}
    reads.sort(key=lambda r: (order[r.reference_name], r.reference_start))
    return reads, contigs


def _read_all(path, chrom_id, chunk, tune=None, **kw):
    rd = bamio.BamReader(path, **kw)
    if tune:
        rd.tune(*tune)
    rd.set_chrom_ids(chrom_id)
    packets = []
    while True:
        pk = rd.next_packet(chunk)
        if pk is None:
            break
        packets.append(pk)
    return rd, packets


@pytest.mark.parametrize("seed,chunk,via_cg", [(1, 1000, False), (2, 37, False), (3, 64, True)])
def test_packet_equals_python_packer(tmp_path, seed, chunk, via_cg):
    reads, contigs = _sorted_reads(seed, 300)
    path = str(tmp_path / "a.bam")
    bam_writer.write_bam(path, contigs, reads, long_cigar_via_cg=via_cg, block_bytes=9000, extra_unmapped=3)
    chrom_names = sorted(n for n, _ in contigs)
    chrom_id = {n: i for i, n in enumerate(chrom_names)}
    rd, packets = _read_all(path, chrom_id, chunk, threads=3)
    assert rd.references == [n for n, _ in contigs] and rd.lengths == [l for _, l in contigs]
    names = rd.names()
    o = 0
    for pk in packets:
        n = len(pk["chrom"])
        assert n <= chunk
        sub = reads[o:o + n]
        want = packing.pack_alignments(sub, chrom_id, {nm: i for i, nm in enumerate(names)})
        for k in ("chrom", "ref_start", "ref_end", "flag", "mapq", "query_len", "read_id", "cigar_off", "sa_off", "cigar"):
            assert np.array_equal(pk[k], want[k]), k
        for k in want["sa"]:
            assert np.array_equal(pk["sa"][k], want["sa"][k]), k
        for i in (0, n // 2, n - 1):
            assert bamio.decode_seq(pk, i) == sub[i].query_sequence
        o += n
    assert o == len(reads)  # the unplaced records at the end are not returned
    rank, uniq = packing.name_ranks(names)
    assert np.array_equal(rd.name_ranks()[:len(names)], rank) and sorted(names) == uniq
    assert rd.index_statistics() == [(nm, sum(1 for r in reads if r.reference_name == nm)) for nm, _ in contigs]
    rd.close()


@pytest.mark.parametrize("tune", [(1, -1), (3, 64), (2, 0), (7, 5000)])
def test_records_straddling_chunks(tmp_path, tune):
    """Tiny chunks (1-7 BGZF blocks of 1.5 KB) and tiny / zero headroom: every record straddles chunk borders and the
    oversized-carry path is taken; the packets must not change."""
    reads, contigs = _sorted_reads(9, 200)
    path = str(tmp_path / "s.bam")
    bam_writer.write_bam(path, contigs, reads, block_bytes=1500, extra_unmapped=2)
    chrom_id = {n: i for i, n in enumerate(sorted(n for n, _ in contigs))}
    rd0, ref = _read_all(path, chrom_id, 1000)
    rd1, got = _read_all(path, chrom_id, 61, tune=tune, threads=3)
    assert rd0.names() == rd1.names()
    cat = lambda ps, k: np.concatenate([p[k] for p in ps])
    for k in ("chrom", "ref_start", "ref_end", "flag", "mapq", "query_len", "read_id", "cigar", "seq4"):
        assert np.array_equal(cat(ref, k), cat(got, k)), k
    for k in ref[0]["sa"]:
        assert np.array_equal(np.concatenate([p["sa"][k] for p in ref]), np.concatenate([p["sa"][k] for p in got])), k
    assert sum(len(p["chrom"]) for p in got) == len(reads)
    rd0.close(); rd1.close()


def test_subset_packet(tmp_path):
    reads, contigs = _sorted_reads(5, 120)
    path = str(tmp_path / "b.bam")
    bam_writer.write_bam(path, contigs, reads)
    chrom_id = {n: i for i, n in enumerate(sorted(n for n, _ in contigs))}
    rd, packets = _read_all(path, chrom_id, 1000)
    pk = packets[0]
    keep = np.flatnonzero((pk["flag"] != 256) & (pk["flag"] != 272))
    sub = bamio.subset_packet(pk, keep)
    names = rd.names()
    want = packing.pack_alignments([reads[i] for i in keep], chrom_id, {nm: i for i, nm in enumerate(names)})
    for k in ("chrom", "ref_start", "ref_end", "flag", "mapq", "query_len", "read_id", "cigar_off", "sa_off", "cigar"):
        assert np.array_equal(sub[k], want[k]), k
    for k in want["sa"]:
        assert np.array_equal(sub["sa"][k], want["sa"][k]), k
    for j in range(len(keep)):   # the subset reads the bases of its parent packet through (seq_lo, seq_hi)
        assert bamio.decode_seq(sub, j) == reads[keep[j]].query_sequence
    assert bamio.subset_packet(pk, np.arange(len(pk["chrom"]))) is pk   # nothing dropped: no copy
    rd.close()


def test_errors(tmp_path):
    bad = tmp_path / "bad.bam"
    bad.write_bytes(b"this is not a bam file at all, just text")
    assert not bamio.is_bam(str(bad))
    with pytest.raises(IOError):
        bamio.BamReader(str(bad))
    with pytest.raises(IOError):
        bamio.BamReader(str(tmp_path / "missing.bam"))
    reads, contigs = _sorted_reads(7, 50)
    path = str(tmp_path / "t.bam")
    bam_writer.write_bam(path, contigs, reads)
    assert bamio.is_bam(path)
    data = open(path, "rb").read()
    trunc = tmp_path / "trunc.bam"
    trunc.write_bytes(data[:len(data) // 2])
    with pytest.raises(IOError, match="truncated"):
        rd = bamio.BamReader(str(trunc))
        while rd.next_packet(1000) is not None:
            pass
    os.remove(path + ".bai")
    rd = bamio.BamReader(path)
    with pytest.raises(IOError):
        rd.index_statistics()
    rd.close()


def test_close_before_eof_and_bgzf_non_bam(tmp_path):
    """Destroying a reader while a read-ahead batch is still inflating, and opening a BGZF file that is not BAM
    (a bgzipped text file): both used to free the batch under the inflate workers."""
    import gzip
    import struct
    import zlib
    reads, contigs = _sorted_reads(5, 1500, with_seq=True)
    path = str(tmp_path / "big.bam")
    bam_writer.write_bam(path, contigs, reads, block_bytes=4000)
    chrom_id = {n: i for i, n in enumerate(sorted(n for n, _ in contigs))}
    for _ in range(20):
        rd = bamio.BamReader(path, threads=4)
        rd.tune(4, -1)          # small batches: several are still ahead of the parser at close()
        rd.set_chrom_ids(chrom_id)
        assert rd.next_packet(10) is not None
        rd.close()
    # a BGZF container whose payload is text (e.g. a .vcf.gz): many blocks so that the read-ahead is busy when open fails
    def bgzf_block(data):
        comp = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = comp.compress(data) + comp.flush()
        head = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(body) + 25)
        return head + body + struct.pack("<II", zlib.crc32(data), len(data))
    txt = str(tmp_path / "x.vcf.gz")
    with open(txt, "wb") as f:
        for i in range(300):
            f.write(bgzf_block((b"##line %d of a text file\n" % i) * 500))
        f.write(bgzf_block(b""))
    assert bamio.is_bam(txt)   # BGZF magic only
    for _ in range(10):
        with pytest.raises(Exception):
            bamio.BamReader(txt, threads=4)
