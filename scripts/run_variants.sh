# A/B of prebuilt library variants (cutesv_b200/variants/*.so): parity subset + benches
for V in "$@"; do
  cp cutesv_b200/variants/$V.so cutesv_b200/libcutesv_b200.so
  timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/tests_$V.txt 2>&1; tail -1 gpurun_out/tests_$V.txt
  for C in 2 3; do
    timeout 250 python bench.py --no-cpu-baseline --config $C > gpurun_out/bench_${V}_c${C}.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
  done
done
