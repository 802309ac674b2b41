#!/usr/bin/env python
"""bench.py -- SV signatures clustered per second (BASELINE.json metric) on N B200s.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on all host cores

A "step" = one pass of the hot path (sort -> chain-linkage clustering -> consensus -> genotype)
over one batch of synthetic signature arrays: BASELINE.json configs[1] "synthetic 30x ONT
whole-genome signature arrays, resolution_INS + resolution_DEL on 1xB200" (16 777 216 signatures,
7.75 M reads-table rows, --genotype).

  value  device-resident throughput (inputs already in HBM), CUDA events on the launching stream
  e2e    the reference-facing call Engine.cluster(): pinned HOST columns in, host rows out
         (H2D + kernels + D2H inside the timed region)
N > 1: one process per GPU (torchrun); every rank owns one genome-equivalent shard of contigs
(weak scaling), no data-path collective, ONE NCCL all-gather of the candidate records at the end.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cutesv_b200 import _abi, synth  # noqa: E402

METRIC = "sv_signatures_clustered_per_sec"
UNIT = "signatures/s"
WORKLOADS = {
    2: "config2: synthetic 30x ONT WGS signature arrays, resolution_INS + resolution_DEL, --genotype",
    3: "config3: synthetic 50x PacBio HiFi, all five SV types (INS/DEL/INV/DUP/TRA) + cal_GL genotyping (TRA from the all-alignments table)",
    4: "config2: synthetic 30x ONT WGS signature arrays, resolution_INS + resolution_DEL, --genotype",
    5: "config5: synthetic 100x ONT ultra-long (deep pile-ups), all five SV types + genotyping",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="workload scale (1.0 = BASELINE config)")
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong", action="store_true",
                    help="BASELINE config 4: ONE genome sharded by contig (LPT) over the N GPUs instead of one genome-equivalent per GPU")
    return ap.parse_args()


def workload(config_id, scale, rank):
    cfg = synth.make_config(config_id, scale, seed=synth.SEED0 + config_id + 1000 * rank)
    return cfg


def algorithmic_bytes(cfg, n_cand):
    """3 x record per signature (48 B DEL-like, 60 B INS/TRA), 16 B per reads row, 64 B per candidate."""
    per = {"DEL": 48, "DUP": 48, "INV": 48, "INS": 60, "TRA": 60}
    b = sum(per[k] * len(v["chrom"]) for k, v in cfg["sigs"].items())
    return b + 16 * len(cfg["reads"]["chrom"]) + 64 * n_cand


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_baseline(cfg, params, threads, repeats=1):
    """The oracle port (oracle/cutesv_oracle.c) on the host cores: sort + cluster + genotype."""
    from oracle import oracle_lib
    oracle_lib.lib()
    best = None
    nc = 0
    for _ in range(repeats):
        t0 = time.perf_counter()
        c, g, n = oracle_lib.cluster(params, cfg["lens"], cfg["sigs"], cfg["reads"], n_threads=threads)
        dt = time.perf_counter() - t0
        nc = len(c)
        best = dt if best is None else min(best, dt)
    return cfg["n_sigs"] / best, best, nc


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (the Python reference cannot
    travel to the GPU box; its tuple/pickle path is ~2 orders of magnitude slower)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload(args.config, args.scale, 0)
    params = _abi.default_params(**cfg["params"])
    threads = os.cpu_count() or 1
    budget_s = float(os.environ.get("CUTESV_B200_REF_BUDGET_S", "180"))
    # one full-workload pass doubles as warm-up and as the probe that sizes the per-step sample: the whole run
    # (steps x sample) is kept within ~budget_s seconds by shrinking the sample (same generator, smaller scale)
    _, dt0, _ = cpu_baseline(cfg, params, threads)
    sample_scale = args.scale
    sample = "full workload per step"
    if dt0 * args.steps > budget_s:
        frac = max(budget_s / (dt0 * args.steps), 0.01)
        sample_scale = args.scale * frac
        cfg = workload(args.config, sample_scale, 0)
        params = _abi.default_params(**cfg["params"])
        sample = "%.3f of the workload per step (same generator at scale %.4f; the full pass took %.2f s)" % (frac, sample_scale, dt0)
    times = []
    for _ in range(args.steps):
        _, dt, nc = cpu_baseline(cfg, params, threads)
        times.append(dt)
    total = sum(times)
    value = cfg["n_sigs"] * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config],
                   "scale": args.scale, "sample_scale": sample_scale, "n_signatures_per_gpu": cfg["n_sigs"],
                   "n_reads_per_gpu": int(len(cfg["reads"]["chrom"])),
                   "note": "CPU arm: rank 0 only, all host cores; per-step sample sized so that the run ends within minutes"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample + " (oracle/cutesv_oracle.c, OpenMP over (type, contig))"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def pinned_copy(torch, cols):
    out = {}
    for k, v in cols.items():
        if v is None:
            out[k] = None
            continue
        t = torch.empty(v.shape, dtype=torch.from_numpy(v[:1].copy()).dtype, pin_memory=True)
        a = t.numpy()
        a[...] = v
        out[k] = a
        out["_t_" + k] = t  # keep the pinned tensor alive
    return out


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from cutesv_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.Stream(device=dev)  # a real (non-default) stream shared by torch events and the library
    torch.cuda.set_stream(stream)

    strong = args.strong and world > 1
    cfg = workload(args.config, args.scale, 0 if strong else rank)
    if strong:  # every rank builds the same seeded genome and keeps its LPT share of the contigs
        from cutesv_b200 import shard
        owner = shard.lpt_assign(shard.contig_weights(cfg["sigs"], len(cfg["lens"])), world)
        my_sigs, my_reads, my_index = shard.shard_inputs(cfg["sigs"], cfg["reads"], owner, rank)
        cfg = dict(cfg, sigs=my_sigs, reads=my_reads, n_sigs=int(sum(len(v["chrom"]) for v in my_sigs.values())))
    params = _abi.default_params(**cfg["params"])
    eng = Engine(local, stream=stream.cuda_stream, params=params, contig_lens=cfg["lens"])
    sigs_p = {k: pinned_copy(torch, v) for k, v in cfg["sigs"].items()}
    reads_p = pinned_copy(torch, cfg["reads"])
    sigs_h = {k: {kk: vv for kk, vv in v.items() if not kk.startswith("_t_")} for k, v in sigs_p.items()}
    reads_h = {kk: vv for kk, vv in reads_p.items() if not kk.startswith("_t_")}
    dev_in = sum(v.nbytes for s in sigs_h.values() for v in s.values() if v is not None) + sum(v.nbytes for v in reads_h.values())
    # e2e inputs: the same rows grouped by contig (as the reference holds them: one list per chromosome, cuteSV:817-857)
    # with row offsets instead of the 4-byte contig column (csv_cluster_host_grouped), in pinned host memory
    n_contigs = len(cfg["lens"])
    sigs_gp = {k: pinned_copy(torch, _abi.group_by_contig(v, n_contigs)) for k, v in cfg["sigs"].items()}
    reads_gp = pinned_copy(torch, _abi.group_by_contig(cfg["reads"], n_contigs))
    sigs_g = {k: {kk: vv for kk, vv in v.items() if not kk.startswith("_t_")} for k, v in sigs_gp.items()}
    reads_g = {kk: vv for kk, vv in reads_gp.items() if not kk.startswith("_t_")}
    h2d = sum(v.nbytes for s in sigs_g.values() for v in s.values() if v is not None) + sum(v.nbytes for v in reads_g.values())
    type_mask = sum(1 << _abi.TYPE_IDS[k] for k in cfg["sigs"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if "TRA" in cfg["sigs"] and cfg["params"].get("genotype"):  # TRA genotyper input: every alignment record in BAM order
        order = np.lexsort((np.arange(len(cfg["reads"]["chrom"])), cfg["reads"]["start"], cfg["reads"]["chrom"]))
        eng.upload_alignments({k: v[order] for k, v in cfg["reads"].items()})
    # ---------------- device-resident: value ----------------
    eng.upload(sigs_h, reads_h)
    for _ in range(max(args.warmup, 3)):
        eng.cluster_device(type_mask)
    n_cand, n_names = eng.counts()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        eng.cluster_device(type_mask)
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    # ---------------- per-kernel durations: the same K steps with the SV-type lanes serialised ----------------
    # (in the timed region above the kernel chains of the SV types overlap on separate streams, so a stage's
    #  CUDA-event interval there includes time the GPU spent on another lane; the roofline of a KERNEL is taken
    #  from this second pass, where every kernel runs alone on the ctx stream between its own events)
    eng.set_lanes(False)
    eng.set_profiling(True)
    eng.cluster_device(type_mask)
    eng.fetch()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for _ in range(args.steps):
        eng.cluster_device(type_mask)
    s1.record(stream)
    barrier()
    serial_ms = s0.elapsed_time(s1) / args.steps
    cands, genos, names = eng.fetch()  # also collects the per-stage events
    stages = {k: v / args.steps for k, v in eng.stage_ms().items()}  # events accumulate over the K steps
    probe = eng.sort_probe()
    ctrs = eng.counters()
    eng.set_profiling(False)
    eng.set_lanes(True)

    # ---------------- end to end through the public call: e2e ----------------
    cap_c = max(2 * n_cand + 1024, 1024)
    pin = [torch.empty(cap_c * 64, dtype=torch.uint8, pin_memory=True), torch.empty(cap_c * 40, dtype=torch.uint8, pin_memory=True),
           torch.empty((2 * n_names + 1024) * 4, dtype=torch.uint8, pin_memory=True)]
    out = (pin[0].numpy().view(_abi.CAND_DTYPE), pin[1].numpy().view(_abi.GENO_DTYPE), pin[2].numpy().view(np.int32))
    for _ in range(max(args.warmup, 3)):
        eng.cluster(sigs_g, reads_g, type_mask, out=out, grouped=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c2, g2, n2 = eng.cluster(sigs_g, reads_g, type_mask, out=out, grouped=True)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    d2h = c2.nbytes + g2.nbytes + n2.nbytes

    # ---------------- max over ranks, one all-gather of the candidate records ----------------
    t = torch.tensor([dev_ms, e2e_s * 1000.0], device=dev, dtype=torch.float64)
    n_sig_total = torch.tensor([float(cfg["n_sigs"])], device=dev, dtype=torch.float64)
    gathered_cands = len(cands)
    allgather_ms = 0.0
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_sig_total, op=dist.ReduceOp.SUM)
        # single all-gather of fixed-width candidate records (padded to the largest shard)
        cnt = torch.tensor([len(cands)], device=dev, dtype=torch.int64)
        dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
        width = int(cnt.item())
        mine = torch.zeros(width * 64, dtype=torch.uint8, device=dev)
        if len(cands):
            mine[: len(cands) * 64] = torch.from_numpy(cands.view(np.uint8).reshape(-1).copy()).to(dev)
        allv = torch.empty(world * width * 64, dtype=torch.uint8, device=dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a0.record()
        dist.all_gather_into_tensor(allv, mine)
        a1.record()
        torch.cuda.synchronize()
        allgather_ms = a0.elapsed_time(a1)
        gathered_cands = world * width
    dev_ms_max, e2e_ms_max = float(t[0].item()), float(t[1].item())
    total_sigs = float(n_sig_total.item())

    if rank == 0:
        value = total_sigs * args.steps / (dev_ms_max / 1000.0)
        e2e_value = total_sigs * args.steps / (e2e_ms_max / 1000.0)
        peak, peak_src = measured_peak()
        # per-stage rooflines: algorithmic bytes of one step / CUDA-event time of the stage
        n_sig = cfg["n_sigs"]
        members = sum(ctrs["members"].values())
        per_step = {
            "keys": ("k_indel_keys (+ density filter)", 20.0 * n_sig, 2),
            "sort": ("k_rs_onesweep (radix scatter pass)", probe["bytes"] / args.steps, probe["launches"] // max(args.steps, 1)),
            "cluster": ("k_cluster_warp (per-cluster consensus)", 28.0 * members + 64.0 * n_cand, 2),
            "genotype": ("k_reads_pass + k_pairs_test (reads table stream)", 17.0 * len(cfg["reads"]["chrom"]) + 16.0 * ctrs["pairs"], 1),
        }
        kernels = {}
        for st, (kname, nbytes, launches_per_step) in per_step.items():
            ms = stages.get(st, 0.0)
            gbs = nbytes / 1e9 / (ms / 1e3) if ms > 0 else 0.0
            kernels[st] = {"kernel": kname, "ms_per_step": ms, "algorithmic_bytes_per_step": nbytes, "launches_per_step": launches_per_step,
                           "achieved": gbs, "frac": gbs / peak}
        # dominant KERNEL (the stages lump several kernels: keys = k_indel_keys + k_bucket_flags + k_prefilter, ...):
        # the one with the largest total time in the committed ncu launch list of this same command
        # (profiles/r01_launches_final.txt); without the file, the longest stage
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        dom_basis = "longest stage by CUDA events"
        try:
            tot = {}
            named = {"k_cluster_warp": "cluster", "k_rs_onesweep": "sort", "k_indel_keys": "keys", "k_reads_pass": "genotype"}
            for ln in open(os.path.join(ROOT, "profiles", "r01_launches_final.txt")):
                if "launches=" not in ln:
                    continue
                nm = ln.split("launches=")[0]
                n_l = int(ln.split("launches=")[1].split()[0])
                avg = float(ln.split("avg=")[1].split()[0])
                for key, st in named.items():
                    if key in nm:
                        tot[st] = tot.get(st, 0.0) + n_l * avg
            if tot and args.config == 2:
                dom = max(tot, key=tot.get)
                dom_basis = "largest kernel total in profiles/r01_launches_final.txt (ncu launch list of this command)"
        except Exception:
            pass
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(dom)
            except Exception:
                traffic = None
        alg = algorithmic_bytes(cfg, n_cand)
        dev_stage_sum = sum(v for k, v in stages.items() if k not in ("h2d", "d2h", "extract"))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config],
                       "scale": args.scale, "n_signatures_per_gpu": cfg["n_sigs"], "n_reads_per_gpu": int(len(cfg["reads"]["chrom"])),
                       "n_candidates": int(n_cand), "parallelism": ("contig-shard x%d (ONE genome, contigs LPT-packed over the GPUs)" if strong else
                                       "contig-shard x%d (one genome-equivalent of contigs per GPU)") % world,
                       "l2": "inputs (%.0f MB/step) larger than the 126 MB L2, no explicit flush" % (dev_in / 1e6),
                       "e2e_inputs": "host columns grouped by contig + row offsets (csv_cluster_host_grouped), pinned",
                       "allgather_ms": allgather_ms, "gathered_candidates": int(gathered_cands),
                       "density_filter_survivors": ctrs["domain"], "kept_clusters": ctrs["kept"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms_max / args.steps},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "hbm", "kernel": kernels[dom]["kernel"], "achieved": kernels[dom]["achieved"], "peak": peak,
                         "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kernels[dom]["algorithmic_bytes_per_step"] / max(kernels[dom]["launches_per_step"], 1),
                         "share_of_step": kernels[dom]["ms_per_step"] / max(dev_stage_sum, 1e-9),
                         "timing": "CUDA events around the stage on the ctx stream, K steps with the SV-type lanes serialised "
                                   "(%.4f ms/step; the timed region overlaps the lanes: %.4f ms/step)" % (serial_ms, dev_ms_max / args.steps),
                         "dominant_by": dom_basis,
                         "note": "latency / instruction bound, not DRAM bound (profiles/); roofline_kernels lists every stage"},
            "roofline_kernels": kernels,
            "roofline_pipeline": {"algorithmic_bytes_per_step": alg, "achieved": alg / 1e9 / (dev_ms_max / args.steps / 1e3), "unit": "GB/s",
                                  "frac": alg / 1e9 / (dev_ms_max / args.steps / 1e3) / peak},
            "stages_ms_per_step": stages,
            "ms_per_step_lanes_serialised": serial_ms,
        }
        if not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            v, dt, nc = cpu_baseline(cfg, params, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "full workload, 1 run of oracle/cutesv_oracle.c (%.2f s)" % dt}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
