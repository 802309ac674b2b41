set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_extract -s 3 -c 1 -f -o gpurun_out/full_extract python scripts/bench_extract.py 15000 7000 > gpurun_out/ncu_extract.log 2>&1
tail -2 gpurun_out/ncu_extract.log
