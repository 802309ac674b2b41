set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_cli_native_bam.py -m gpu -x -q > gpurun_out/tests_check4.txt 2>&1
tail -4 gpurun_out/tests_check4.txt
timeout 200 python scripts/bench_extract.py 15000 7000 2>&1 | tail -1 | cut -c1-330
timeout 200 python scripts/bench_extract.py 200000 850 2>&1 | tail -1 | cut -c1-330
python scripts/kernel_times.py 2 1.0 10 2>&1 | tee gpurun_out/kt_c2.txt | head -14
python scripts/kernel_times.py 3 1.0 10 2>&1 | tee gpurun_out/kt_c3.txt | head -16
