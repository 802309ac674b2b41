# where the fixed per-kernel cost of a 1/8-scale step (one rank of an 8-GPU strong-scaling run) goes
set -u
mkdir -p gpurun_out
export CUTESV_B200_GRAPHS=0
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.max --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_small.csv python scripts/run_steps.py 2 0.125 4 0 > gpurun_out/ncu_small.log 2>&1
tail -1 gpurun_out/ncu_small.log
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k_bucket_fixup|k_select_heads|k_cluster_warp|k_scan_excl' -s 40 -c 12 -f -o gpurun_out/r02_full_small python scripts/run_steps.py 2 0.125 4 0 > gpurun_out/ncu_small2.log 2>&1
tail -1 gpurun_out/ncu_small2.log
