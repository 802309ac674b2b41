# GPU box: ncu captures of one device-resident step of config 2 (lanes serialised, graphs off so that every kernel is a plain launch)
set -u
mkdir -p gpurun_out
export CUTESV_B200_GRAPHS=0
# launch list of the 4th step (3 warm steps precede it): time, instructions, DRAM bytes per launch
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none \
    --csv --log-file gpurun_out/launches_c2.csv python scripts/run_steps.py 2 1.0 4 0 > gpurun_out/launches_c2.log 2>&1
tail -1 gpurun_out/launches_c2.log
# full sections (with source) of one launch of every kernel that carries the step
timeout 1200 ncu --set full --import-source on --clock-control none \
    -k 'regex:k_indel_hist|k_bucket_prefix|k_indel_scatter|k_bucket_fixup$|k_member_records|k_select_heads|k_cluster_warp|k_reads_pass|k_pairs_test|k_scan_excl' \
    -s 30 -c 12 -f -o gpurun_out/full_c2 python scripts/run_steps.py 2 1.0 4 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
