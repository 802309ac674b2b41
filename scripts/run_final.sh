# GPU box: full confirmation run (tests, smoke, default bench with the CPU baseline, config 3 / 5 benches)
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests_full.txt 2>&1; tail -3 gpurun_out/tests_full.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -1 gpurun_out/smoke.txt
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
timeout 300 python bench.py --no-cpu-baseline --config 3 > gpurun_out/bench_final_c3.json 2>> gpurun_out/err.txt
timeout 300 python bench.py --no-cpu-baseline --config 5 --scale 0.5 > gpurun_out/bench_final_c5.json 2>> gpurun_out/err.txt
