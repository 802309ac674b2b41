import sys, time
sys.path.insert(0, '.')
import numpy as np
from cutesv_b200 import synth, _abi
from cutesv_b200.engine import Engine
from oracle import oracle_lib, compare_records
e = Engine(0)
for cid, sc in ((2, 0.004), (3, 0.01), (2, 0.1), (2, 1.0)):
    cfg = synth.make_config(cid, sc)
    p = _abi.default_params(**cfg['params'])
    e.set_params(p); e.set_contigs(cfg['lens'])
    e.set_profiling(True)
    t = time.time(); got = e.cluster(cfg['sigs'], cfg['reads']); t1 = time.time() - t
    t = time.time(); got = e.cluster(cfg['sigs'], cfg['reads']); t2 = time.time() - t
    print('config', cid, sc, 'n_sigs', cfg['n_sigs'], 'cands', len(got[0]), 'e2e %.4f s (2nd %.4f)' % (t1, t2), flush=True)
    print('  stages', {k: round(v, 3) for k, v in e.stage_ms().items()}, e.sort_probe(), flush=True)
    t = time.time(); ref = oracle_lib.cluster(p, cfg['lens'], cfg['sigs'], cfg['reads'], n_threads=16); t3 = time.time() - t
    d = compare_records.diff_records(ref, got)
    print('  oracle %.3f s' % t3, 'DIFFS' if d else 'PARITY OK', flush=True)
    for m in d[:3]: print(m)
