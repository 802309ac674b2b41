"""Device-resident steps of one config (for ncu / sanitizer captures).  python scripts/run_steps.py [config] [scale] [steps] [lanes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from cutesv_b200 import _abi, synth
from cutesv_b200.engine import Engine

cid = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 1
cfg = synth.make_config(cid, scale)
p = _abi.default_params(**cfg["params"])
e = Engine(0, params=p, contig_lens=cfg["lens"])
e.set_lanes(bool(lanes))
mask = sum(1 << _abi.TYPE_IDS[k] for k in cfg["sigs"])
if "TRA" in cfg["sigs"]:
    r = cfg["reads"]
    order = np.lexsort((np.arange(len(r["chrom"])), r["start"], r["chrom"]))
    e.upload_alignments({k: v[order] for k, v in r.items()})
e.upload(cfg["sigs"], cfg["reads"])
for _ in range(steps):
    e.cluster_device(mask)
c, g, n = e.fetch()
print("config %d scale %g: %d candidates after %d steps, %d launches, %d graph replays" % (cid, scale, len(c), steps, e.launch_count(), e.graph_replays()))
