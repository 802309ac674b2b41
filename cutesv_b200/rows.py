"""Candidate records (csv_cand / csv_geno) -> the row lists the reference's resolution_* return.

Row layouts follow resolveINDEL.py:197-219,408-432,464-478; resolveDUP.py:114-131,170-180;
resolveINV.py:136-156,240-251; resolveTRA.py:171-182 of the reference.  These rows are what
generate_output (cuteSV_genotype.py:242-467) indexes by position.
"""
from ._abi import CSV_DEL, CSV_DUP, CSV_F_NO_READS, CSV_INS, CSV_INV, CSV_TRA, TYPE_NAMES

GT_STR = {0: "0/0", 1: "0/1", 2: "1/1", -1: "./."}
TRA_TYPES = "ABCD"


def qual_str(q):
    """str(np.float64) / str(float) as the reference prints QUAL (cuteSV_genotype.py:54)."""
    return str(float(q))


def tra_alt(bnd_type, chr2, pos2):
    """BND ALT string of resolveTRA.py:140-153,217-225 (types A/C add 1 to the mate position)."""
    bnd_pos = "%s:%s" % (chr2, pos2 + (1 if bnd_type in ("A", "C") else 0))
    if bnd_type == "A":
        return "N[%s[" % bnd_pos
    if bnd_type == "B":
        return "N]%s]" % bnd_pos
    if bnd_type == "C":
        return "[%s[N" % bnd_pos
    return "]%s]N" % bnd_pos


def _geno_fields(g, action):
    if not action or g["status"] != 0:
        return ".", "./.", ".,.,.", ".", "."
    return (str(int(g["dr"])), GT_STR[int(g["gt"])],
            "%d,%d,%d" % (int(g["pl"][0]), int(g["pl"][1]), int(g["pl"][2])),
            str(int(g["gq"])), qual_str(g["qual"]))


def record_to_row(c, g, names_buf, chrom_names, read_name, ins_seq, action, name_sets=False):
    """One candidate -> one reference row.

    read_name: callable id -> str.  ins_seq: callable input_index -> str (INS ALT source).
    name_sets: DUP/TRA RNAMES come from Python set iteration in the reference (hash-order
    dependent, resolveDUP.py:82,96 / resolveTRA.py:182); ids are emitted sorted here.
    """
    t = int(c["svtype"])
    chrom = chrom_names[int(c["chrom"])]
    ids = names_buf[int(c["names_off"]): int(c["names_off"]) + int(c["names_cnt"])]
    names = ",".join(read_name(int(i)) for i in ids)
    dr, gt, gl, gq, qual = _geno_fields(g, action)
    if t in (CSV_DEL, CSV_INS):
        row = [chrom, TYPE_NAMES[t], str(int(c["pos"])), str(int(c["len"])), str(int(c["support"])),
               "-%d,%d" % (int(c["cipos"]), int(c["cipos"])), "-%d,%d" % (int(c["cilen"]), int(c["cilen"])),
               dr, gt, gl, gq, qual, names]
        if t == CSV_INS:
            row.append(ins_seq(int(c["aux"]))[0:int(c["len"])])
        return row
    if t == CSV_DUP:
        return [chrom, "DUP", str(int(c["pos"])), str(int(c["len"])), str(int(c["support"])),
                dr, gt, gl, gq, qual, names]
    if t == CSV_INV:
        strand = "++" if int(c["aux"]) == 0 else "--"
        return [chrom, "INV", str(int(c["pos"])), str(int(c["len"])), str(int(c["support"])),
                dr, gt, strand, gl, gq, qual, names]
    if t == CSV_TRA:
        chr2 = chrom_names[int(c["aux"]) >> 2]
        alt = tra_alt(TRA_TYPES[int(c["aux"]) & 3], chr2, int(c["pos2"]))
        return [chrom, alt, str(int(c["pos"])), chr2, str(int(c["pos2"])), str(int(c["support"])),
                dr, gt, gl, gq, qual, names]
    raise ValueError("bad svtype %d" % t)


def records_to_rows(cands, genos, names_buf, chrom_names, read_name, ins_seq, action):
    """All candidates -> {(svtype_name, chrom_name): [rows]} in reference emission order.

    Candidates flagged CSV_F_NO_READS are dropped (call_gt returns [] when the contig has no
    reads-table rows, resolveINDEL.py:443-444).
    Same rows as record_to_row() per candidate; the record fields are taken out of the structured arrays column by
    column first (one .tolist() per field instead of ~25 numpy scalar accesses per candidate).
    """
    out = {}
    n = len(cands)
    if n == 0:
        return out
    f = {k: cands[k].tolist() for k in ("svtype", "chrom", "pos", "len", "support", "cipos", "cilen", "aux", "pos2", "names_off", "names_cnt", "flags")}
    g_status, g_dr, g_gt, g_gq, g_qual = (genos[k].tolist() for k in ("status", "dr", "gt", "gq", "qual"))
    g_pl = genos["pl"].tolist()
    ids_all = names_buf.tolist() if hasattr(names_buf, "tolist") else list(names_buf)
    svt, chrom_id, flags = f["svtype"], f["chrom"], f["flags"]
    for i in range(n):
        t = svt[i]
        chrom = chrom_names[chrom_id[i]]
        rows = out.setdefault((TYPE_NAMES[t], chrom), [])
        if action and (flags[i] & CSV_F_NO_READS):
            continue
        o = f["names_off"][i]
        names = ",".join([read_name(k) for k in ids_all[o:o + f["names_cnt"][i]]])
        if not action or g_status[i] != 0:
            dr, gt, gl, gq, qual = ".", "./.", ".,.,.", ".", "."
        else:
            pl = g_pl[i]
            dr, gt, gl, gq, qual = str(g_dr[i]), GT_STR[g_gt[i]], "%d,%d,%d" % (pl[0], pl[1], pl[2]), str(g_gq[i]), str(float(g_qual[i]))
        pos, ln, sup, aux = str(f["pos"][i]), str(f["len"][i]), str(f["support"][i]), f["aux"][i]
        if t == CSV_DEL or t == CSV_INS:
            ci, cl = f["cipos"][i], f["cilen"][i]
            row = [chrom, TYPE_NAMES[t], pos, ln, sup, "-%d,%d" % (ci, ci), "-%d,%d" % (cl, cl), dr, gt, gl, gq, qual, names]
            if t == CSV_INS:
                row.append(ins_seq(aux)[0:f["len"][i]])
        elif t == CSV_DUP:
            row = [chrom, "DUP", pos, ln, sup, dr, gt, gl, gq, qual, names]
        elif t == CSV_INV:
            row = [chrom, "INV", pos, ln, sup, dr, gt, "++" if aux == 0 else "--", gl, gq, qual, names]
        elif t == CSV_TRA:
            chr2 = chrom_names[aux >> 2]
            row = [chrom, tra_alt(TRA_TYPES[aux & 3], chr2, f["pos2"][i]), pos, chr2, str(f["pos2"][i]), sup, dr, gt, gl, gq, qual, names]
        else:
            raise ValueError("bad svtype %d" % t)
        rows.append(row)
    return out


def records_to_rows_slow(cands, genos, names_buf, chrom_names, read_name, ins_seq, action):
    """records_to_rows through record_to_row(), one candidate at a time (kept as the cross-check of the column-wise version)."""
    out = {}
    for i in range(len(cands)):
        c = cands[i]
        if action and (int(c["flags"]) & CSV_F_NO_READS):
            out.setdefault((TYPE_NAMES[int(c["svtype"])], chrom_names[int(c["chrom"])]), [])
            continue
        row = record_to_row(c, genos[i], names_buf, chrom_names, read_name, ins_seq, action)
        out.setdefault((TYPE_NAMES[int(c["svtype"])], chrom_names[int(c["chrom"])]), []).append(row)
    return out
