set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_cli_native_bam.py tests/test_gpu_dropin.py -m gpu -x -q > gpurun_out/tests_check3.txt 2>&1
tail -8 gpurun_out/tests_check3.txt
python scripts/kernel_times.py 2 1.0 10 2>&1 | tee gpurun_out/kt_c2.txt | head -14
timeout 600 python bench.py --config 5 --scale 0.25 --steps 10 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err || tail -5 gpurun_out/bench_c5.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_c5.json").read().strip().splitlines()[-1])
    print("config 5 x0.25: ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"])
    print("extract leg", json.dumps(d.get("e2e_extract")))
except Exception as e:
    print("bench failed", e)
PY
