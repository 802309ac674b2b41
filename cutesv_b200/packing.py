"""Host side of the boundary: decoded alignment records -> pinned-friendly columnar int32 buffers.

Packs the fields the reference reads at cuteSV:606-733 from duck-typed read objects (pysam records for CRAM / SAM
input, test doubles; plain BAM goes through the native decoder in bamio.py instead) (.flag .mapq .query_length
.query_name .reference_start .reference_end .cigartuples .get_tags()).
"""
import re

import numpy as np

_CIG = re.compile(r"(\d+)([MIDNSHP=X])")
_REFSPAN_OPS = frozenset("MD=X")


def acquire_clip_pos(cigar_string):
    """(first S length, last S length, ref span) of an SA-tag CIGAR string -- the three numbers the
    reference derives at cuteSV:466-481 (clips from a leading / trailing 'S' only; span = M+D+=+X)."""
    items = _CIG.findall(cigar_string)
    first = int(items[0][0]) if items and items[0][1] == "S" else 0
    last = int(items[-1][0]) if items and items[-1][1] == "S" else 0
    span = sum(int(n) for n, op in items if op in _REFSPAN_OPS)
    return first, last, span


def name_ranks(names):
    """Rank of every name in Python string order (the reference's tuple sorts break ties on the
    read-name string, cuteSV:764-801).  Returns (rank array, sorted unique names)."""
    uniq = sorted(set(names))
    index = {n: i for i, n in enumerate(uniq)}
    return np.fromiter((index[n] for n in names), dtype=np.int32, count=len(names)), uniq


def pack_alignments(reads, chrom_id, read_id):
    """reads: sequence of read objects; chrom_id: dict name -> contig id; read_id: dict name -> id;
    each read object needs `.reference_name`.  Returns dict(read_cols..., cigar, sa_cols...)."""
    n = len(reads)
    cols = {k: np.zeros(n, dtype=np.int32) for k in ("chrom", "ref_start", "ref_end", "flag", "mapq", "query_len", "read_id")}
    cigar_off = np.zeros(n + 1, dtype=np.int64)
    sa_off = np.zeros(n + 1, dtype=np.int64)
    cig = []
    sa = {k: [] for k in ("chrom", "pos0", "strand", "mapq", "first_clip", "last_clip", "ref_span")}
    for i, r in enumerate(reads):
        cols["chrom"][i] = chrom_id[r.reference_name]
        cols["ref_start"][i] = r.reference_start
        cols["ref_end"][i] = r.reference_end
        cols["flag"][i] = r.flag
        cols["mapq"][i] = r.mapq
        cols["query_len"][i] = r.query_length
        cols["read_id"][i] = read_id[r.query_name]
        for op, ln in r.cigartuples:
            cig.append((ln << 4) | op)
        cigar_off[i + 1] = len(cig)
        for tag in r.get_tags():
            if tag[0] == "SA":
                for ent in tag[1].split(";")[:-1]:  # cuteSV:678
                    f = ent.split(",")
                    first, last, span = acquire_clip_pos(f[3])
                    sa["chrom"].append(chrom_id[f[0]])
                    sa["pos0"].append(int(f[1]) - 1)  # SA pos is 1-based, cuteSV:497
                    sa["strand"].append(0 if f[2] == "+" else 1)
                    sa["mapq"].append(int(f[4]))
                    sa["first_clip"].append(first)
                    sa["last_clip"].append(last)
                    sa["ref_span"].append(span)
        sa_off[i + 1] = len(sa["chrom"])
    out = dict(cols)
    out["cigar_off"] = cigar_off
    out["sa_off"] = sa_off
    out["cigar"] = np.array(cig, dtype=np.uint32)
    out["sa"] = {k: np.array(v, dtype=np.int32) for k, v in sa.items()}
    return out


_COMP = str.maketrans("ACGTNacgtn", "TGCANtgcan")


def revcomp(s):
    return s.translate(_COMP)[::-1]


def merged_ins_from_cigar(cigar, ref_start, query, pos, min_siglength, merge_ins_threshold):
    """Sequence of the merged insertion signature that starts at reference position `pos`: parse_read's CIGAR walk
    (cuteSV:616-645) + generate_combine_sigs' INS chain (cuteSV:530-545) on one record.  Only for signatures that merge more
    insertions than the device buffers (piece flag 2)."""
    ref = int(ref_start)
    first = int(cigar[0]) if len(cigar) else 0
    q = -(first >> 4) if (first & 15) == 5 else 0          # shift_ins_read starts at -hardclip_left
    groups = []                                              # [first pos, last pos, [slices]]
    for cg in cigar:
        op, ln = int(cg) & 15, int(cg) >> 4
        if op != 2:
            q += ln                                          # every op except D advances the query cursor (cuteSV:631-632)
        if ln >= min_siglength and op in (1, 2):
            if op == 2:
                ref += ln
            else:
                piece = query[q - ln:q]   # Python slice semantics, as the reference (cuteSV:639)
                if groups and ref - groups[-1][1] <= merge_ins_threshold:
                    groups[-1][1] = ref
                    groups[-1][2].append(piece)
                else:
                    groups.append([ref, ref, [piece]])
        elif op in (0, 2, 3, 7, 8):
            ref += ln
    for g in groups:
        if g[0] == pos:
            return "".join(g[2])
    raise ValueError("merged insertion at %d not found in the record's CIGAR" % pos)


def ins_sequence(pieces, off, cnt, query_of, cigar_of=None, merge=None):
    """Rebuild an INS signature's sequence from its piece list; query_of(rec) -> query string.
    A piece with flag 2 (more merged insertions than the device buffers) is rebuilt from the record's CIGAR:
    cigar_of(rec) -> (uint32 CIGAR array, reference_start), merge = (min_siglength, merge_ins_threshold)."""
    out = []
    for p in range(off, off + cnt):
        rec, a, b, rc = (int(x) for x in pieces[p])
        q = query_of(rec)
        if rc == 2:
            cig, ref_start = cigar_of(rec)
            out.append(merged_ins_from_cigar(cig, ref_start, q, a, merge[0], merge[1]))
            continue
        if rc:
            q = revcomp(q)
        out.append(q[a:b])
    return "".join(out)
