set -u
mkdir -p gpurun_out
export CUTESV_B200_GRAPHS=0
timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:k_reads_pass|k_cluster_small|k_bucket_fixup|k_select_heads' -s 21 -c 7 -f -o gpurun_out/full_c2b python scripts/run_steps.py 2 1.0 5 0 > gpurun_out/ncu_full2.log 2>&1
tail -1 gpurun_out/ncu_full2.log
