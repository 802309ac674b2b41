# GPU box with N >= 2 GPUs: merged N-rank parity + the strong-scaling bench (ONE genome sharded, all-gather inside the step)
set -u
N=${1:-2}
mkdir -p gpurun_out
true
true
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_strong_${N}.json 2> gpurun_out/bench_strong_${N}.err || tail -20 gpurun_out/bench_strong_${N}.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_strong_${N}.json").read().strip().splitlines()[-1])
    print("N=${N} strong: ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "allgather alone ms", d["config"]["allgather_ms_alone"], "replays", d["config"]["graph_replays_in_timed_region"], "launches", d["gpu_launches"])
except Exception as e:
    print("bench failed", e)
PY
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err || tail -20 gpurun_out/bench_n1.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
    print("N=1: ms/step", d["ms_per_step"], "value", d["value"], "e2e ms", d["e2e"]["ms_per_step"], "replays", d["config"]["graph_replays_in_timed_region"], "launches", d["gpu_launches"], "noise", d["ms_per_step_noise"])
    print("dominant", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["avg_us_per_launch"])
except Exception as e:
    print("bench failed", e)
PY
