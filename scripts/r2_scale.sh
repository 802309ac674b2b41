# 8-GPU box: strong-scaling benches N = 8, 4, 2, 1 (ONE genome sharded over N GPUs, all-gather inside the step)
set -u
mkdir -p gpurun_out
for N in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_strong_${N}.json 2> gpurun_out/bench_strong_${N}.err || tail -5 gpurun_out/bench_strong_${N}.err
done
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_strong_1.json 2> gpurun_out/bench_strong_1.err || tail -5 gpurun_out/bench_strong_1.err
python - <<PY
import json
for N in (1, 2, 4, 8):
    try:
        d=json.loads(open("gpurun_out/bench_strong_%d.json" % N).read().strip().splitlines()[-1])
        print("N=%d %s: ms/step %.4f value %.3e e2e ms %.3f allgather alone ms %.4f launches %d stages %s" % (N, d["scaling"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["config"]["allgather_ms_alone"], d["gpu_launches"], {k: round(v,3) for k,v in d["stages_ms_per_step"].items()}))
    except Exception as e:
        print("N=%d failed: %s" % (N, e))
PY
