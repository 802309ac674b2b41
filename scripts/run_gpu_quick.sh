# quick GPU confirmation: parity suites + benches of configs 2, 3
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_dropin.py tests/test_gpu_cli.py tests/test_cli_native_bam.py -m gpu -x -q > gpurun_out/tests.txt 2>&1
tail -3 gpurun_out/tests.txt
for C in 2 3; do
  timeout 250 python bench.py --no-cpu-baseline --config $C > gpurun_out/bench_c${C}.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
done
