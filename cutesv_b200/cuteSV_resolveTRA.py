"""Drop-in for the reference's cuteSV_resolveTRA (resolveTRA.py:30-104,257-258).

Clustering AND genotyping run on the GPU.  The reference's call_gt (resolveTRA.py:260-309) re-opens the
BAM per candidate and iterates bam.fetch() with an early exit; here pysam is used ONLY to decode the
records around the candidate breakpoints into a packed all-alignments table (BAM order), which the
device genotyper scans (csv_upload_alignments -> k_tra_genotype).  No host genotype computation."""
import numpy as np

from . import _abi, rows, runtime, workdir


def _fetch_alignments(bam_path, windows, chrom_id, name_id_of):
    """Decode every record overlapping the (merged) windows, per contig in BAM order."""
    import pysam
    bam = pysam.AlignmentFile(bam_path)
    cols = {k: [] for k in ("chrom", "start", "end", "read_id", "is_primary")}
    try:
        for chrom in sorted(windows, key=lambda c: chrom_id[c]):
            iv = sorted(windows[chrom])
            merged = []
            for s, e in iv:
                if merged and s <= merged[-1][1]:
                    merged[-1][1] = max(merged[-1][1], e)
                else:
                    merged.append([s, e])
            seen = set()
            recs = []
            for s, e in merged:
                for r in bam.fetch(chrom, s, e):
                    key = (r.reference_start, r.reference_end, r.query_name, r.flag)
                    if key in seen:  # a record spanning two disjoint windows is returned by both fetches
                        continue
                    seen.add(key)
                    recs.append(r)
            recs.sort(key=lambda r: r.reference_start)  # stable: BAM order inside one start
            for r in recs:
                cols["chrom"].append(chrom_id[chrom]); cols["start"].append(r.reference_start); cols["end"].append(r.reference_end)
                cols["read_id"].append(name_id_of(r.query_name)); cols["is_primary"].append(1 if r.flag in (0, 16) else 0)
        lens = {c: bam.get_reference_length(c) for c in windows}
    finally:
        bam.close()
    return {k: np.asarray(v, dtype=np.uint8 if k == "is_primary" else np.int32) for k, v in cols.items()}, lens


def resolution_TRA(path, chr_1, read_count, overlap_size, max_cluster_bias, bam_path, action, gt_round, sigs_index):
    if chr_1 not in sigs_index["TRA"]:
        return (chr_1, [])
    seqs = workdir.load_slice(path, "TRA", chr_1, sigs_index)
    name_id, names = workdir.name_index((seqs, 4))
    chroms = sorted(set([chr_1] + [t[2] for t in seqs]))
    chrom_id = {c: i for i, c in enumerate(chroms)}
    cols = workdir.tuples_to_columns("TRA", seqs, chrom_id, name_id)
    hi = max([1] + [int(cols[k].max()) for k in ("a", "b") if len(cols[k])])
    eng = runtime.get_engine()
    p = _abi.default_params(min_support=read_count, ratio_tra=overlap_size, bias_tra=max_cluster_bias, genotype=0, gt_round=gt_round)
    eng.set_params(p)
    eng.set_contigs(np.full(len(chroms), hi + max_cluster_bias + 2, dtype=np.int64))
    res = eng.cluster({"TRA": cols}, None, type_mask=1 << _abi.CSV_TRA)
    if action and len(res[0]):
        # windows of call_gt: [pos - bias, pos + bias] on both contigs (resolveTRA.py:264-265, 292-293)
        windows = {}
        for c in res[0]:
            windows.setdefault(chroms[int(c["chrom"])], []).append((max(int(c["pos"]) - max_cluster_bias, 0), int(c["pos"]) + max_cluster_bias))
            windows.setdefault(chroms[int(c["aux"]) >> 2], []).append((max(int(c["pos2"]) - max_cluster_bias, 0), int(c["pos2"]) + max_cluster_bias))
        extra = {}

        def nid(name):  # names outside the signature set only need distinct ids
            i = name_id.get(name)
            if i is None:
                i = extra.get(name)
                if i is None:
                    i = len(names) + len(extra)
                    extra[name] = i
            return i
        aln, lens = _fetch_alignments(bam_path, windows, chrom_id, nid)
        p.genotype = 1
        eng.set_params(p)
        eng.set_contigs(np.array([lens.get(c, hi + max_cluster_bias + 2) for c in chroms], dtype=np.int64))
        eng.upload_alignments(aln)
        try:
            res = eng.cluster({"TRA": cols}, None, type_mask=1 << _abi.CSV_TRA)
        finally:
            eng.upload_alignments(None)
    out = rows.records_to_rows(res[0], res[1], res[2], chroms, lambda i: names[i], None, bool(action))
    return (chr_1, out.get(("TRA", chr_1), []))


def run_tra(args):
    return resolution_TRA(*args)
