set -u
mkdir -p gpurun_out
export CUTESV_B200_GRAPHS=0
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_cluster_warp -s 6 -c 2 -f -o gpurun_out/full_cluster python scripts/run_steps.py 2 1.0 5 0 > gpurun_out/ncu_cluster.log 2>&1
tail -2 gpurun_out/ncu_cluster.log
