"""Authoring container only: the UNMODIFIED Python reference's rebuild + clustering phase (process_process_sigs_type +
run_del / run_ins, cuteSV:750-857, 1113-1199, --genotype) on a bounded sample of the bench workload (config 2 at a small scale),
timed beside the C port (oracle/cutesv_oracle.c) on the same arrays.  python scripts/time_python_reference.py [scale]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cutesv_b200 import _abi, synth  # noqa: E402
from oracle import oracle_lib, ref_harness  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
cfg = synth.make_config(2, scale)
p = _abi.default_params(**cfg["params"])
t0 = time.perf_counter()
rows = ref_harness.run_reference(cfg["sigs"], cfg["reads"], cfg["names"], synth.read_name, p, types_=("DEL", "INS"))
t_ref = time.perf_counter() - t0
t0 = time.perf_counter()
c, g, n = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"], n_threads=1)
t_port1 = time.perf_counter() - t0
t0 = time.perf_counter()
c, g, n = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"], n_threads=os.cpu_count())
t_portN = time.perf_counter() - t0
print(json.dumps(dict(config=2, scale=scale, n_signatures=cfg["n_sigs"], n_reads=len(cfg["reads"]["chrom"]),
                      python_reference_s=t_ref, python_reference_sig_per_s=cfg["n_sigs"] / t_ref, python_reference_cores=1,
                      python_reference_rows=sum(len(v) for v in rows.values()),
                      c_port_1core_s=t_port1, c_port_1core_sig_per_s=cfg["n_sigs"] / t_port1,
                      c_port_all_cores_s=t_portN, c_port_all_cores_sig_per_s=cfg["n_sigs"] / t_portN, cores=os.cpu_count(),
                      c_port_rows=len(c))))
