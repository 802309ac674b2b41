import sys
sys.path.insert(0, '.')
import numpy as np
from cutesv_b200 import synth, _abi
from cutesv_b200.engine import Engine
from oracle import oracle_lib, compare_records
cfg = synth.make_config(5, 0.25)
p = _abi.default_params(**cfg['params'])
e = Engine(0, params=p, contig_lens=cfg['lens'])
runs = [e.cluster(cfg['sigs'], cfg['reads']) for _ in range(3)]
print('counters', e.counters())
ref = oracle_lib.cluster(p, cfg['lens'], cfg['sigs'], cfg['reads'], n_threads=0)
for i, r in enumerate(runs):
    d = compare_records.diff_records(ref, r, max_report=3)
    print('run', i, 'n', len(r[0]), 'vs oracle:', 'OK' if not d else 'DIFF')
    for m in d: print(m[:700])
for t in range(5):
    m = runs[0][0]['svtype'] == t
    print(t, int(m.sum()))
