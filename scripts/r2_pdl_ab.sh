set -u
mkdir -p gpurun_out
for CFG in "5 0.5" "3 1.0" "2 1.0"; do
  set -- $CFG
  for PDL in 0 1; do
    CUTESV_B200_PDL=$PDL timeout 300 python bench.py --config $1 --scale $2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdlab_$1_$PDL.json 2> gpurun_out/bench_pdlab.err || tail -3 gpurun_out/bench_pdlab.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/bench_pdlab_$1_$PDL.json").read().strip().splitlines()[-1])
print("config $1 scale $2 pdl=$PDL: ms/step %.4f (serialised %.4f)" % (d["ms_per_step"], d["ms_per_step_lanes_serialised"]))
PY
  done
done
