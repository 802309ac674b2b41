# A/B of programmatic dependent launches + parity
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for PDL in 1 0; do
for S in 0.125 1.0; do
  CUTESV_B200_PDL=$PDL timeout 300 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline --scale $S > gpurun_out/bench_pdl${PDL}_$S.json 2> gpurun_out/bench_pdl${PDL}_$S.err || tail -5 gpurun_out/bench_pdl${PDL}_$S.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pdl${PDL}_$S.json").read().strip().splitlines()[-1])
print("pdl=$PDL scale $S: ms/step %.4f launches %d replays %s" % (d["ms_per_step"], d["gpu_launches"], d["config"]["graph_replays_in_timed_region"]))
PY
done
done
