set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for MODE in p2p nccl; do
  CUTESV_B200_GATHER=$MODE timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${MODE}_${N}.json 2> gpurun_out/bench_${MODE}_${N}.err || tail -5 gpurun_out/bench_${MODE}_${N}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${MODE}_${N}.json").read().strip().splitlines()[-1])
print("${MODE} N=${N}: ms/step %.4f value %.3e e2e ms %.3f allgather alone ms %.4f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["config"]["allgather_ms_alone"]))
PY
done
