set -u
python scripts/kernel_times.py 2 1.0 10 2>&1 | grep "k_cluster_warp\|small_path" | cut -c1-600
CUTESV_B200_SMALL_PATH=0 python scripts/kernel_times.py 2 1.0 10 2>&1 | grep "k_cluster_warp" | cut -c1-200
