"""Throughput of kernel (a) (CIGAR walk) on a vectorised ONT-shaped packet: CIGAR ops/s and GB/s
against the measured HBM peak.  python scripts/bench_extract.py [n_reads]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
import numpy as np

from cutesv_b200 import _abi, synth
from cutesv_b200.engine import Engine

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 200000
mean_indels = int(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("-") else 850
pk, names, lens = synth.synth_cigar_packet(n_reads, mean_indels=mean_indels, sa_frac=0.2 if mean_indels > 2000 else 0.0)
p = _abi.default_params()
e = Engine(0, params=p, contig_lens=lens)
e.set_profiling(True)
for _ in range(3):
    r = e.extract(pk)
ms = []
wall = []
for _ in range(10):
    t0 = time.perf_counter()
    r = e.extract(pk)
    wall.append(time.perf_counter() - t0)
    ms.append(e.stage_ms()["extract"])
n_ops = len(pk["cigar"])
alg = 4.0 * n_ops + 44.0 * n_reads + 28.0 * len(pk["sa"]["chrom"])
peak = 6482.7
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
dev = float(np.median(ms))
out = dict(n_reads=n_reads, n_cigar_ops=n_ops, counts=r["counts"], device_ms=dev, ops_per_s=n_ops / (dev / 1e3),
           achieved_GBs=alg / 1e9 / (dev / 1e3), frac_of_measured_hbm=alg / 1e9 / (dev / 1e3) / peak, e2e_ms=float(np.median(wall)) * 1e3,
           e2e_ops_per_s=n_ops / float(np.median(wall)))
# parity of this packet's signatures against the emulator on a sample of the reads
if "--check" in sys.argv:
    import emul_lib
    sub = 2000
    sl = {k: (v[:sub] if k not in ("cigar", "sa", "cigar_off", "sa_off") else v) for k, v in pk.items()}
    sl["cigar_off"] = pk["cigar_off"][: sub + 1]
    sl["sa_off"] = pk["sa_off"][: sub + 1]
    sl["cigar"] = pk["cigar"][: pk["cigar_off"][sub]]
    e.extract(sl)
    got = e.fetch_extracted()
    ref = emul_lib.extract(p, sl)
    for t in ("DEL", "INS"):
        a = sorted(zip(got["sigs"][t]["a"].tolist(), got["sigs"][t]["b"].tolist(), got["sigs"][t]["read_id"].tolist()))
        b = sorted(zip(ref["sigs"][t]["a"].tolist(), ref["sigs"][t]["b"].tolist(), ref["sigs"][t]["read_id"].tolist()))
        assert a == b, t
    out["parity_sample_reads"] = sub
print(json.dumps(out))
