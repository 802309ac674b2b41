"""Test-only BAM/BAI writer (SAM/BAM spec, BGZF via zlib raw deflate): turns duck-typed read objects
into a real .bam so the native decoder (cutesv_b200/bamio.py) is exercised on the on-disk format."""
import struct
import zlib

_SEQ_CODE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def _bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25  # header 18 + trailer 8 - 1
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", bsize) + body
            + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def _reg2bin(beg, end):
    end -= 1
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return base + (beg >> shift)
    return 0


def _record(r, ref_id, long_cigar_via_cg):
    name = r.query_name.encode() + b"\0"
    cig = [(ln << 4) | op for op, ln in r.cigartuples]
    seq = r.query_sequence or ""
    l_seq = len(seq) if seq else int(r.query_length)
    if seq:
        codes = [_SEQ_CODE.get(c.upper(), 15) for c in seq]
    else:
        codes = [15] * l_seq  # 'N' padding keeps l_seq == query_length for sequence-less synthetic reads
    if len(codes) & 1:
        codes.append(0)
    packed = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))
    aux = b""
    real = cig
    if long_cigar_via_cg:
        span = sum(ln for op, ln in r.cigartuples if op in (0, 2, 3, 7, 8))
        aux += b"CGBI" + struct.pack("<I", len(cig)) + struct.pack("<%dI" % len(cig), *cig)
        real = [(l_seq << 4) | 4, (span << 4) | 3]
    for tag in r.get_tags():
        if isinstance(tag[1], str):
            aux += tag[0].encode() + b"Z" + tag[1].encode() + b"\0"
        else:
            aux += tag[0].encode() + b"i" + struct.pack("<i", int(tag[1]))
    body = struct.pack("<iiBBHHHiiii", ref_id, r.reference_start, len(name), r.mapq,
                       _reg2bin(r.reference_start, max(r.reference_end, r.reference_start + 1)), len(real), r.flag, l_seq, -1, -1, 0)
    body += name + struct.pack("<%dI" % len(real), *real) + packed + b"\xff" * l_seq + aux
    return struct.pack("<i", len(body)) + body


def write_bam(path, contigs, reads, long_cigar_via_cg=False, block_bytes=60000, extra_unmapped=0):
    """contigs: [(name, length)] in header order; reads: objects sorted by (header index, start).
    Also writes <path>.bai holding only the per-contig pseudo-bins (mapped counts)."""
    index = {n: i for i, (n, _) in enumerate(contigs)}
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % c for c in contigs)
    stream = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(contigs)))
    for n, ln in contigs:
        stream += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", ln)
    mapped = [0] * len(contigs)
    for r in reads:
        stream += _record(r, index[r.reference_name], long_cigar_via_cg)
        mapped[index[r.reference_name]] += 1
    for k in range(extra_unmapped):  # unplaced records at the end, as a sorted BAM has them
        nm = ("unmapped%d" % k).encode() + b"\0"
        body = struct.pack("<iiBBHHHiiii", -1, -1, len(nm), 0, 4680, 0, 4, 4, -1, -1, 0) + nm + b"\x11\x11" + b"\xff" * 4
        stream += struct.pack("<i", len(body)) + body
    with open(path, "wb") as f:
        for o in range(0, len(stream), block_bytes):
            f.write(_bgzf_block(bytes(stream[o:o + block_bytes])))
        f.write(_bgzf_block(b""))  # EOF marker
    with open(path + ".bai", "wb") as f:
        f.write(b"BAI\1" + struct.pack("<i", len(contigs)))
        for m in mapped:
            f.write(struct.pack("<i", 1) + struct.pack("<Ii", 37450, 2) + struct.pack("<QQQQ", 0, 0, m, 0) + struct.pack("<i", 0))
        f.write(struct.pack("<Q", extra_unmapped))
