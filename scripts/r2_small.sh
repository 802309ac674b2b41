# per-rank work of an 8-GPU strong-scaling step on ONE GPU: config 2 at scale 1/8 (graph replay vs per-kernel events)
set -u
mkdir -p gpurun_out
for S in 0.125 0.25; do
  timeout 300 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline --scale $S > gpurun_out/bench_small_$S.json 2> gpurun_out/bench_small_$S.err || tail -5 gpurun_out/bench_small_$S.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_small_$S.json").read().strip().splitlines()[-1])
print("scale $S: ms/step %.4f launches %d" % (d["ms_per_step"], d["gpu_launches"]))
print("  ", {k: round(v["avg_us"],1) for k,v in d["roofline_kernels"].items()})
print("  serialised sum %.1f us" % sum(v["ms_per_step"]*1e3 for v in d["roofline_kernels"].values()))
PY
done
