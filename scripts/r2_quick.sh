# GPU box: round-2 quick confirmation: parity suites (new front end + graphs), then a short bench of configs 2 and 3
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/tests_quick.txt 2>&1
tail -5 gpurun_out/tests_quick.txt
for C in 2 3; do
  timeout 300 python bench.py --no-cpu-baseline --config $C --steps 20 > gpurun_out/bench_q_c${C}.json 2> gpurun_out/err_c${C}.txt || tail -5 gpurun_out/err_c${C}.txt
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_q_c${C}.json"))
    print("config ${C}: ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "stages", d["stages_ms_per_step"])
except Exception as e:
    print("bench ${C} failed", e)
PY
done
