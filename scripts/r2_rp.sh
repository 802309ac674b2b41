set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_rp.json 2> gpurun_out/bench_rp.err || tail -5 gpurun_out/bench_rp.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_rp.json").read().strip().splitlines()[-1])
k=d["roofline_kernels"]
print("ms/step %.4f" % d["ms_per_step"], {n: round(v["avg_us"],1) for n,v in k.items()})
PY
