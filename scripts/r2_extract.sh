set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python scripts/bench_extract.py 15000 7000 2>&1 | tail -2
