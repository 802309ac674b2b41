# GPU box: the captures behind profiles/ (round 1 final).  Reports land in gpurun_out/.
set -u
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
# one whole step (the 4th: three warm-up steps of ~36 launches precede it) with time, instructions and DRAM bytes per launch
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -s 96 -c 64 --csv --log-file gpurun_out/launches_c2.csv $B > /dev/null 2>&1
# full sections of the kernels that carry the step (one step's worth: 18 matching launches after 3 warm steps)
timeout 900 ncu --set full --import-source on --clock-control none \
    -k 'regex:k_cluster_warp|k_rs_onesweep|k_indel_keys|k_reads_pass|k_prefilter|k_pairs_test|k_select_heads' -s 54 -c 18 -f -o gpurun_out/full_c2 $B > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
