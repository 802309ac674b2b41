"""Per-kernel device times of one config (lanes serialised, CUDA events around every launch).
python scripts/kernel_times.py [config] [scale] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from cutesv_b200 import _abi, synth
from cutesv_b200.engine import Engine

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cfg = synth.make_config(cid, scale)
p = _abi.default_params(**cfg["params"])
e = Engine(0, params=p, contig_lens=cfg["lens"])
mask = sum(1 << _abi.TYPE_IDS[k] for k in cfg["sigs"])
if "TRA" in cfg["sigs"]:
    r = cfg["reads"]
    order = np.lexsort((np.arange(len(r["chrom"])), r["start"], r["chrom"]))
    e.upload_alignments({k: v[order] for k, v in r.items()})
e.upload(cfg["sigs"], cfg["reads"])
for _ in range(3):
    e.cluster_device(mask)
e.fetch()
e.set_lanes(False)
e.set_profiling(True)
for _ in range(steps):
    e.cluster_device(mask)
e.fetch()
kt = e.kernel_times()
tot = sum(v[1] for v in kt.values())
print("config %d scale %g: %d signatures, per step (lanes serialised):" % (cid, scale, cfg["n_sigs"]))
for nm, (n, ms) in sorted(kt.items(), key=lambda x: -x[1][1]):
    print("  %-58s launches/step %5.1f  avg %8.2f us  total %8.1f us/step  %5.1f %%" % (nm, n / steps, 1e3 * ms / n, 1e3 * ms / steps, 100 * ms / tot))
print("  sum %.1f us/step; counters %s" % (1e3 * tot / steps, json.dumps(e.counters())))
