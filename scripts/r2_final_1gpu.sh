# 1-GPU box: the whole -m gpu suite + smoke + the benches / captures behind profiles/r02_*
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_all_gpu.txt 2>&1
tail -4 gpurun_out/tests_all_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || tail -5 gpurun_out/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err || tail -5 gpurun_out/r02_bench_reference.err
timeout 400 python bench.py --config 3 --no-cpu-baseline > gpurun_out/r02_bench_config3.json 2> gpurun_out/err3.txt || tail -5 gpurun_out/err3.txt
timeout 400 python bench.py --config 5 --scale 0.5 --no-cpu-baseline > gpurun_out/r02_bench_config5_scale0p5.json 2> gpurun_out/err5.txt || tail -5 gpurun_out/err5.txt
python - <<PY
import json
for f in ("r02_bench_n1", "r02_bench_config3", "r02_bench_config5_scale0p5", "r02_bench_reference"):
    try:
        d=json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "ms/step %.4f value %.3e e2e %.3e" % (d["ms_per_step"], d["value"], d["e2e"]["value"]), d.get("roofline", {}).get("kernel"), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(f, "failed", e)
PY
export CUTESV_B200_GRAPHS=0
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/r02_launches_c2.csv python scripts/run_steps.py 2 1.0 4 0 > gpurun_out/launches_c2.log 2>&1
timeout 1200 ncu --set full --import-source on --clock-control none \
    -k 'regex:k_indel_hist|k_bucket_prefix|k_indel_scatter|k_bucket_fixup|k_select_heads|k_cluster_small|k_cluster_warp|k_reads_pass|k_pairs_test' \
    -s 27 -c 13 -f -o gpurun_out/r02_full_c2 python scripts/run_steps.py 2 1.0 4 0 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_extract -s 3 -c 1 -f -o gpurun_out/r02_full_extract python scripts/bench_extract.py 15000 7000 > gpurun_out/ncu_extract.log 2>&1
tail -1 gpurun_out/ncu_extract.log
unset CUTESV_B200_GRAPHS
timeout 900 python scripts/bench_cli.py 1000000 --genotype --profile > gpurun_out/r02_cli_1m.json 2> gpurun_out/r02_cli_1m.err || tail -5 gpurun_out/r02_cli_1m.err
tail -1 gpurun_out/r02_cli_1m.json
