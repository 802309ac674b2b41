set -u
mkdir -p gpurun_out
timeout 900 python scripts/bench_cli.py 1000000 --genotype --profile > gpurun_out/r02_cli_1m.json 2> gpurun_out/r02_cli_1m.err || tail -5 gpurun_out/r02_cli_1m.err
tail -1 gpurun_out/r02_cli_1m.json
grep -c WARNING gpurun_out/r02_cli_1m.err
