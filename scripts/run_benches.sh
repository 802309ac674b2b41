# GPU box: the three bench lines kept under profiles/
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/err.txt || tail -5 gpurun_out/err.txt
timeout 300 python bench.py --no-cpu-baseline --config 3 > gpurun_out/bench_final_c3.json 2>> gpurun_out/err.txt
timeout 300 python bench.py --no-cpu-baseline --config 5 --scale 0.5 > gpurun_out/bench_final_c5.json 2>> gpurun_out/err.txt
