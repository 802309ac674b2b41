"""How long does one giant chain cluster take (global-scratch CTA team)?  python scripts/pileup_timing.py N"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from cutesv_b200 import _abi, synth
from cutesv_b200.engine import Engine
from oracle import oracle_lib, compare_records
n = int(sys.argv[1])
rng = np.random.default_rng(1)
names, lens = synth.contigs(0.01)
n_reads = max(n // 3, 10)
sig = dict(chrom=np.zeros(n, np.int32), a=(50000 + rng.integers(0, 2000, n)).astype(np.int32), b=(200 + rng.integers(0, 400, n)).astype(np.int32),
           read_id=rng.integers(0, n_reads, n).astype(np.int32), c=None)
reads = dict(chrom=np.zeros(n_reads, np.int32), start=np.full(n_reads, 1000, np.int32) + rng.integers(0, 40000, n_reads).astype(np.int32),
             end=np.full(n_reads, 90000, np.int32), read_id=np.arange(n_reads, dtype=np.int32), is_primary=np.ones(n_reads, np.uint8))
p = _abi.default_params(min_support=10, genotype=1, bias_del=200)
e = Engine(0, params=p, contig_lens=lens)
e.cluster({"DEL": sig}, reads)
t = time.perf_counter(); got = e.cluster({"DEL": sig}, reads); dt = time.perf_counter() - t
print("n", n, "gpu %.3f s" % dt, "cands", len(got[0]), e.counters()["giant"], flush=True)
t = time.perf_counter(); ref = oracle_lib.cluster(p, lens, {"DEL": sig}, reads); print("oracle %.3f s" % (time.perf_counter() - t))
d = compare_records.diff_records(ref, got)
print("parity", "OK" if not d else d[:2])
