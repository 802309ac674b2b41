# 2-GPU box: 1-GPU parity suites, N-rank parity, per-kernel table, N=1 and N=2 benches
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/tests_check2.txt 2>&1
tail -6 gpurun_out/tests_check2.txt
python scripts/kernel_times.py 2 1.0 10 2>&1 | tee gpurun_out/kt_c2.txt | head -24
bash scripts/r2_multi.sh 2 2>&1 | grep -v "^$" | tail -8
