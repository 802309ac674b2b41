set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/tests_check5.txt 2>&1
tail -4 gpurun_out/tests_check5.txt
python scripts/kernel_times.py 2 1.0 10 2>&1 | tee gpurun_out/kt_c2.txt | head -12
python scripts/kernel_times.py 3 1.0 10 2>&1 | tee gpurun_out/kt_c3.txt | head -8
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err || tail -20 gpurun_out/bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print("N=1: ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"], "serial", d["ms_per_step_lanes_serialised"])
PY
