"""Record-level comparison (csv_cand / csv_geno / names) used by the parity tests."""
import numpy as np

CAND_FIELDS = ("svtype", "chrom", "pos", "len", "support", "cipos", "cilen", "search_pos", "pos2", "aux",
               "names_cnt", "flags")
GENO_FIELDS = ("dr", "dv", "gt", "gq", "qual")


def diff_records(ref, got, max_report=5):
    """ref/got = (cands, genos, names).  `cluster`, `reserved` and names_off are layout details
    and not compared; the supporting-read lists are compared slice by slice."""
    rc, rg, rn = ref
    gc, gg, gn = got
    msgs = []
    if len(rc) != len(gc):
        msgs.append("candidate count: ref %d got %d" % (len(rc), len(gc)))
    n = min(len(rc), len(gc))
    for i in range(n):
        bad = [f for f in CAND_FIELDS if rc[i][f] != gc[i][f]]
        if rg[i]["status"] != gg[i]["status"]:
            bad.append("geno.status")
        elif rg[i]["status"] == 0:
            bad += ["geno." + f for f in GENO_FIELDS if rg[i][f] != gg[i][f]]
            if tuple(rg[i]["pl"]) != tuple(gg[i]["pl"]):
                bad.append("geno.pl")
        a = rn[rc[i]["names_off"]: rc[i]["names_off"] + rc[i]["names_cnt"]]
        b = gn[gc[i]["names_off"]: gc[i]["names_off"] + gc[i]["names_cnt"]]
        if len(a) != len(b) or not np.array_equal(a, b):
            bad.append("names")
        if bad:
            msgs.append("cand %d differs in %s\n  ref %s %s\n  got %s %s" % (i, bad, rc[i], rg[i], gc[i], gg[i]))
            if len(msgs) >= max_report:
                break
    return msgs
