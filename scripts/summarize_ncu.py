"""Turns ncu output into the text summaries committed under profiles/.

  python scripts/summarize_ncu.py launches <launch-list.csv> <title>   (ncu --csv --log-file ... with --metrics)
  python scripts/summarize_ncu.py full <report.ncu-rep> <title>        (ncu --set full ...; needs ncu on PATH)
"""
import collections
import csv
import subprocess
import sys

RAW_KEYS = [
    ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB"), ("dram__bytes_write.sum", "MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "%"),
    ("smsp__inst_executed.sum", "inst"), ("launch__registers_per_thread", ""), ("launch__grid_size", ""),
    ("launch__waves_per_multiprocessor", ""), ("sm__icc_request_hit_rate.pct", "%"), ("l1tex__t_sector_hit_rate.pct", "%"),
    ("lts__t_sector_hit_rate.pct", "%"),
]
STALLS = ["long_scoreboard", "short_scoreboard", "no_instruction", "wait", "barrier", "mio_throttle", "lg_throttle", "branch_resolving",
          "math_pipe_throttle", "not_selected", "dispatch_stall", "membar"]


def launches(path, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ik, im, iv, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv or r[0] == "":
            continue
        name = r[ik].split("(")[0]
        d = per.setdefault((r[0], name), {})
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        if r[im] == "gpu__time_duration.sum":
            v = v / 1000.0 if unit in ("nsecond", "ns") else v
        if r[im].startswith("dram__bytes"):
            v = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit, 1.0)
        d[r[im]] = v
    agg = collections.OrderedDict()
    for (_, name), d in per.items():
        a = agg.setdefault(name, collections.Counter())
        a["n"] += 1
        for k, v in d.items():
            a[k] += v
    total = sum(a["gpu__time_duration.sum"] for a in agg.values())
    print("# " + title)
    print("# per-launch times are cold-cache and serialised under ncu: compare SHARES with bench.py's stages_ms_per_step")
    for name, a in sorted(agg.items(), key=lambda x: -x[1]["gpu__time_duration.sum"]):
        n = a["n"]
        print("%-44s launches=%3d avg=%8.1f us share=%5.1f%% warp_inst=%.1fM dram_read/write=%.0f/%.0f MB per launch" % (
            name[:44], n, a["gpu__time_duration.sum"] / n, 100 * a["gpu__time_duration.sum"] / total,
            a.get("smsp__inst_executed.sum", 0) / n / 1e6, a.get("dram__bytes_read.sum", 0) / n, a.get("dram__bytes_write.sum", 0) / n))
    print("# total %.1f us over %d launches" % (total, sum(a["n"] for a in agg.values())))


def full(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    print("# " + title)
    for r in rows[2:]:
        d = dict(zip(h, r))
        u = dict(zip(h, units))
        print("== %s  [launch id %s]" % (d.get("Kernel Name", "?")[:90], d.get("ID", "?")))
        for k, _ in RAW_KEYS:
            if k in d:
                print("   %-70s %s [%s]" % (k, d[k], u.get(k, "")))
        for st in STALLS:
            k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % st
            if k in d:
                print("   stall:%-63s %s" % (st, d[k]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
