# A/B of the per-type lanes (CUTESV_B200_LANES=0 serialises the SV types on one stream)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_dropin.py tests/test_gpu_cli.py tests/test_cli_native_bam.py -m gpu -x -q > gpurun_out/tests.txt 2>&1
tail -3 gpurun_out/tests.txt
for L in 1 0; do
  for C in 2 3; do
    CUTESV_B200_LANES=$L timeout 250 python bench.py --no-cpu-baseline --config $C > gpurun_out/bench_lanes${L}_c${C}.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
  done
done
CUTESV_B200_LANES=1 timeout 250 python bench.py --no-cpu-baseline --config 5 --scale 0.5 > gpurun_out/bench_lanes1_c5.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 150 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --config 3 > /dev/null 2>&1
