"""Host-side parity sink: candidate rows -> VCF records.

Restates generate_output (cuteSV_genotype.py:242-467), the header (cuteSV_Description.py:265-305)
and the serial SVID numbering (cuteSV:1208-1237).  Pure string formatting + FASTA lookups: it stays on
the host by design; it consumes exactly the rows the reference's resolution_* return.
"""
import os
import time

VERSION = "2.1.4"

_ALT = (("INS", "Insertion of novel sequence relative to the reference"), ("DEL", "Deletion relative to the reference"),
        ("DUP", "Region of elevated copy number relative to the reference"), ("INV", "Inversion of reference sequence"),
        ("BND", "Breakend of translocation"))
_INFO = (("PRECISE", "0", "Flag", "Precise structural variant"), ("IMPRECISE", "0", "Flag", "Imprecise structural variant"),
         ("SVTYPE", "1", "String", "Type of structural variant"),
         ("SVLEN", "1", "Integer", "Difference in length between REF and ALT alleles"),
         ("CHR2", "1", "String", "Chromosome for END coordinate in case of a translocation"),
         ("END", "1", "Integer", "End position of the variant described in this record"),
         ("CIPOS", "2", "Integer", "Confidence interval around POS for imprecise variants"),
         ("CILEN", "2", "Integer", "Confidence interval around inserted/deleted material between breakends"),
         ("RE", "1", "Integer", "Number of read support this record"),
         ("STRAND", "A", "String", "Strand orientation of the adjacency in BEDPE format (DEL:+-, DUP:-+, INV:++/--)"),
         ("RNAMES", ".", "String", "Supporting read names of SVs (comma separated)"), ("AF", "A", "Float", "Allele Frequency."))
_FORMAT = (("GT", "1", "String", "Genotype"), ("DR", "1", "Integer", "# High-quality reference reads"),
           ("DV", "1", "Integer", "# High-quality variant reads"),
           ("PL", "G", "Integer", "# Phred-scaled genotype likelihoods rounded to the closest integer"),
           ("GQ", "1", "Integer", "# Genotype quality"))
_IUPAC = str.maketrans("RYSWKMBDHV", "ACCAGACAAA")  # cuteSV_genotype.py:262
_RECORD = "{CHR}\t{POS}\t{ID}\t{REF}\t{ALT}\t{QUAL}\t{PASS}\t{INFO}\tGT:DR:DV:PL:GQ\t{GT}:{DR}:{RE}:{PL}:{GQ}\n"


def header_lines(contig_info, sample, argv, date=None):
    out = ["##fileformat=VCFv4.2", "##source=cuteSV-%s" % VERSION,
           "##fileDate=%s" % (date if date is not None else time.strftime("%Y-%m-%d %H:%M:%S %w-%Z", time.localtime()))]
    out += ["##contig=<ID=%s,length=%d>" % (c[0], c[1]) for c in contig_info]
    out += ['##ALT=<ID=%s,Description="%s">' % a for a in _ALT]
    out += ['##INFO=<ID=%s,Number=%s,Type=%s,Description="%s">' % i for i in _INFO]
    out.append('##FILTER=<ID=q5,Description="Quality below 5">')
    out += ['##FORMAT=<ID=%s,Number=%s,Type=%s,Description="%s">' % f for f in _FORMAT]
    out.append('##CommandLine="cuteSV %s"' % " ".join(argv))
    out.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s" % sample)
    return out


def _af(re_, dr):
    try:
        return ";AF=" + str(round(int(re_) / (int(re_) + int(dr)), 4))
    except Exception:
        return ";AF=."


def _filter(qual):
    if qual == "." or qual is None:
        return "PASS"
    return "PASS" if float(qual) >= 5.0 else "q5"


def format_records(rows, ref_seq, opts):
    """One contig's rows (DEL, INS, INV, DUP, TRA rows in that order) -> [(svtype_for_id, line)].

    opts: dict(genotype, max_size, min_size, report_readid, ignore_sequence).  ref_seq: the contig's
    reference string.  Mirrors generate_output line by line in behaviour (cuteSV_genotype.py:252-458)."""
    rows = sorted(rows, key=lambda x: int(x[2]))  # stable
    action = opts["genotype"]
    max_size, min_size = opts["max_size"], opts["min_size"]
    out = []
    for r in rows:
        kind = r[1]
        if kind in ("DEL", "INS"):
            size = abs(int(float(r[3])))
            if (size > max_size and max_size != -1) or size < min_size:
                continue
            pos = int(r[2])
            end = pos if kind == "INS" else pos + size
            info = "%s;SVTYPE=%s;SVLEN=%s;END=%d;CIPOS=%s;CILEN=%s;RE=%s%s" % (
                "IMPRECISE" if r[8] == "0/0" else "PRECISE", kind, r[3], end, r[5], r[6], r[4],
                (";RNAMES=" + r[12]) if opts["report_readid"] else "")
            if action:
                info += _af(r[4], r[7])
            if kind == "DEL":
                info += ";STRAND=+-"
            if opts["ignore_sequence"]:
                ref, alt = "N", "<%s>" % kind
            elif kind == "INS":
                ref = ref_seq[max(pos - 1, 0)]
                alt = ref + r[13]
            else:
                ref = ref_seq[max(pos - 1, 0):pos - int(r[3])]
                alt = ref_seq[max(pos - 1, 0)]
            out.append((kind, _RECORD.format(CHR=r[0], POS=str(pos), ID="cuteSV.%s.<SVID>" % kind, REF=ref.translate(_IUPAC), ALT=alt,
                                             QUAL=r[11], PASS=_filter(r[11]), INFO=info, GT=r[8], DR=r[7], RE=r[4], PL=r[9], GQ=r[10])))
        elif kind == "DUP":
            size = abs(int(float(r[3])))
            if size > max_size and max_size != -1:
                continue
            pos = int(r[2])
            info = "%s;SVTYPE=DUP;SVLEN=%s;END=%d;RE=%s;STRAND=-+%s" % ("IMPRECISE" if r[6] == "0/0" else "PRECISE", r[3], pos + 1 + size, r[4],
                                                                       (";RNAMES=" + r[10]) if opts["report_readid"] else "")
            if action:
                info += _af(r[4], r[5])
            out.append(("DUP", _RECORD.format(CHR=r[0], POS=str(pos + 1), ID="cuteSV.DUP.<SVID>", REF=ref_seq[pos].translate(_IUPAC), ALT="<DUP>",
                                              QUAL=r[9], PASS=_filter(r[9]) if r[9] != "." else "PASS", INFO=info, GT=r[6], DR=r[5], RE=r[4],
                                              PL=r[7], GQ=r[8])))
        elif kind == "INV":
            size = abs(int(float(r[3])))
            if size > max_size and max_size != -1:
                continue
            if r[7] == "++":
                pos = int(r[2])
                ref_idx = max(pos - 1, 0)
            else:
                pos = int(r[2]) + 1
                ref_idx = int(r[2])
            info = "%s;SVTYPE=INV;SVLEN=%s;END=%d;RE=%s;STRAND=%s%s" % ("IMPRECISE" if r[6] == "0/0" else "PRECISE", r[3], pos + size, r[4], r[7],
                                                                       (";RNAMES=" + r[11]) if opts["report_readid"] else "")
            if action:
                info += _af(r[4], r[5])
            out.append(("INV", _RECORD.format(CHR=r[0], POS=str(pos), ID="cuteSV.INV.<SVID>", REF=ref_seq[ref_idx].translate(_IUPAC), ALT="<INV>",
                                              QUAL=r[10], PASS=_filter(r[10]) if r[10] != "." else "PASS", INFO=info, GT=r[6], DR=r[5],
                                              RE=r[4], PL=r[8], GQ=r[9])))
        else:  # BND rows: [chr1, ALT, pos1, chr2, pos2, RE, DR, GT, PL, GQ, QUAL, names]
            info = "%s;SVTYPE=BND;RE=%s%s" % ("IMPRECISE" if r[7] == "0/0" else "PRECISE", r[5], (";RNAMES=" + r[11]) if opts["report_readid"] else "")
            if action:
                info += _af(r[5], r[6])
            if r[1][0] == "N":  # types A/B: the coordinate is already 1-based
                pos = int(r[2])
                try:
                    base = ref_seq[max(pos - 1, 0)]
                except Exception:
                    base = "N"
                alt = base + r[1][1:]
            else:               # types C/D
                pos = int(r[2]) + 1
                try:
                    base = ref_seq[int(r[2])]
                except Exception:
                    base = "N"
                alt = r[1][:-1] + base
            out.append(("BND", _RECORD.format(CHR=r[0], POS=str(pos), ID="cuteSV.BND.<SVID>", REF=base.translate(_IUPAC), ALT=alt, QUAL=r[10],
                                              PASS=_filter(r[10]) if r[10] != "." else "PASS", INFO=info, GT=r[7], DR=r[6], RE=r[5],
                                              PL=r[8], GQ=r[9])))
    return out


def assign_ids(per_chrom_records):
    """Serial <SVID> substitution in string-sorted contig order (cuteSV:1208-1237)."""
    counters = {"INS": 0, "DEL": 0, "BND": 0, "DUP": 0, "INV": 0}
    lines = []
    for chrom in sorted(per_chrom_records):
        for kind, line in per_chrom_records[chrom]:
            lines.append(line.replace("<SVID>", str(counters[kind])))
            counters[kind] += 1
    return lines


def read_fasta(path):
    """Minimal FASTA reader: {contig name (first word of the header): sequence}."""
    seqs, name, chunks = {}, None, []
    with open(path) as f:
        for line in f:
            if line.startswith(">"):
                if name is not None:
                    seqs[name] = "".join(chunks)
                name, chunks = line[1:].split()[0], []
            else:
                chunks.append(line.strip())
    if name is not None:
        seqs[name] = "".join(chunks)
    return seqs


class _LazySeq(object):
    """One contig of an indexed FASTA: `seq[i]` / `seq[a:b]` read just those bases (str semantics for the
    non-negative indices generate_output uses)."""

    def __init__(self, fh, length, offset, line_bases, line_width):
        self.fh, self.length, self.offset, self.lb, self.lw = fh, length, offset, line_bases, line_width

    def __len__(self):
        return self.length

    def _read(self, a, b):
        if b <= a:
            return ""
        start = self.offset + (a // self.lb) * self.lw + a % self.lb
        end = self.offset + ((b - 1) // self.lb) * self.lw + (b - 1) % self.lb + 1
        self.fh.seek(start)
        return self.fh.read(end - start).replace(b"\n", b"").replace(b"\r", b"").decode()

    def __getitem__(self, key):
        if isinstance(key, slice):
            a, b, step = key.indices(self.length)
            if step != 1:
                return self._read(0, self.length)[key]
            return self._read(a, b)
        i = key + self.length if key < 0 else key
        if not 0 <= i < self.length:
            raise IndexError("string index out of range")
        return self._read(i, i + 1)


class IndexedFasta(object):
    """dict-like {contig: sequence} over a FASTA without loading it: uses <path>.fai when present, else builds the same
    index in one pass.  Falls back to read_fasta() for files faidx could not index (ragged line lengths)."""

    def __init__(self, path):
        self.entries = {}
        self.full = None
        fai = path + ".fai"
        with open(path, "rb") as f:
            magic = f.read(2)
        if magic == b"\x1f\x8b":
            # gzip / bgzip reference (pysam.FastaFile reads .fa.gz through its .gzi): the .fai offsets address the
            # UNcompressed stream, so decompress once instead of seeking into compressed bytes
            import gzip
            import io
            self.full = {}
            name, chunks = None, []
            with gzip.open(path, "rb") as g:
                for line in io.BufferedReader(g):
                    if line.startswith(b">"):
                        if name is not None:
                            self.full[name] = b"".join(chunks).decode()
                        name, chunks = line[1:].split()[0].decode(), []
                    elif name is not None:
                        chunks.append(line.strip())
            if name is not None:
                self.full[name] = b"".join(chunks).decode()
            self.fh = None
            return
        if os.path.exists(fai):
            with open(fai) as f:
                for line in f:
                    c = line.rstrip("\n").split("\t")
                    if len(c) >= 5:
                        self.entries[c[0]] = (int(c[1]), int(c[2]), int(c[3]), int(c[4]))
        else:
            ok = self._build(path)
            if not ok:
                self.full = read_fasta(path)
        self.fh = open(path, "rb") if self.full is None else None

    def length_of(self, name):
        """Sequence length of a contig according to the index (None when the contig is absent)."""
        if self.full is not None:
            return len(self.full[name]) if name in self.full else None
        return self.entries[name][0] if name in self.entries else None

    def _build(self, path):
        name, length, offset, lb, lw, short_seen, pos = None, 0, 0, 0, 0, False, 0
        with open(path, "rb") as f:
            for line in f:
                n = len(line)
                if line.startswith(b">"):
                    if name is not None:
                        self.entries[name] = (length, offset, lb or 1, lw or 1)
                    name, length, offset, lb, lw, short_seen = line[1:].split()[0].decode(), 0, pos + n, 0, 0, False
                elif name is not None:
                    bases = len(line.rstrip(b"\r\n"))
                    if short_seen and bases:
                        return False              # a full-length line after a shorter one: not indexable
                    if lb == 0:
                        lb, lw = bases, n
                    elif bases != lb or n != lw:
                        if bases > lb:
                            return False
                        short_seen = True
                    length += bases
                pos += n
        if name is not None:
            self.entries[name] = (length, offset, lb or 1, lw or 1)
        return True

    def __contains__(self, chrom):
        return chrom in (self.full if self.full is not None else self.entries)

    def __getitem__(self, chrom):
        if self.full is not None:
            return self.full[chrom]
        return _LazySeq(self.fh, *self.entries[chrom])

    def close(self):
        if self.fh:
            self.fh.close()
            self.fh = None


def write_vcf(path, results_by_chrom, reference, contig_info, sample, argv, opts, date=None):
    """results_by_chrom: {chrom: rows in DEL, INS, INV, DUP, TRA order (cuteSV:1191-1199)}."""
    per = {}
    for chrom, rows in results_by_chrom.items():
        if chrom not in reference:
            raise Exception("No corresponding contig in reference with %s." % chrom)
        per[chrom] = format_records(rows, reference[chrom], opts)
    with open(path, "w") as f:
        f.write("\n".join(header_lines(contig_info, sample, argv, date)) + "\n")
        for line in assign_ids(per):
            f.write(line)
