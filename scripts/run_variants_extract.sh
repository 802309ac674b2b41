# A/B of prebuilt library variants on the extraction kernel: parity + scripts/bench_extract.py
for V in "$@"; do
  cp cutesv_b200/variants/$V.so cutesv_b200/libcutesv_b200.so
  timeout 600 python -m pytest tests/test_gpu_extract.py -m gpu -x -q > gpurun_out/tests_$V.txt 2>&1; tail -1 gpurun_out/tests_$V.txt
  timeout 300 python scripts/bench_extract.py 200000 --check > gpurun_out/extract_$V.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
  cat gpurun_out/extract_$V.json
done
