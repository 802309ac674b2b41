set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
N=2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_strong_${N}.json 2> gpurun_out/bench_strong_${N}.err || tail -5 gpurun_out/bench_strong_${N}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_strong_2.json").read().strip().splitlines()[-1])
print("N=2 strong: ms/step %.4f value %.3e e2e ms %.3f allgather alone ms %.4f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["config"]["allgather_ms_alone"]))
print({k: round(v["avg_us"],1) for k,v in d["roofline_kernels"].items()})
PY
