set -u
for V in none; do
  cp scripts/variants/lib_$V.so cutesv_b200/libcutesv_b200.so
  echo "== $V"; timeout 200 python scripts/bench_extract.py 15000 7000 2>&1 | tail -1 | cut -c1-330
  timeout 200 python scripts/bench_extract.py 200000 850 2>&1 | tail -1 | cut -c1-330
done
timeout 400 python -m pytest tests/test_gpu_extract.py -m gpu -x -q 2>&1 | tail -2
