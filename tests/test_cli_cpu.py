"""CPU: CLI flag surface identical to the reference's parseArgs; task windows like cuteSV:1018-1044."""
import pytest

from cutesv_b200 import cli
from oracle import ref_harness


def test_defaults_and_flags_match_reference():
    if not ref_harness.available():
        pytest.skip("reference not present (GPU box)")
    ref_harness.modules()
    from cuteSV.cuteSV_Description import parseArgs
    for argv in (["a.bam", "r.fa", "o.vcf", "wd"],
                 ["a.bam", "r.fa", "o.vcf", "wd", "--genotype", "-s", "3", "-l", "50", "-L", "-1", "-t", "4", "-b", "500", "-p", "-1", "-q", "10",
                  "-r", "100", "-md", "500", "-mi", "500", "-sl", "20", "--max_cluster_bias_INS", "1000", "--diff_ratio_merging_INS", "0.9",
                  "--max_cluster_bias_DEL", "1000", "--diff_ratio_merging_DEL", "0.5", "--max_cluster_bias_INV", "7", "--max_cluster_bias_DUP", "8",
                  "--max_cluster_bias_TRA", "9", "--diff_ratio_filtering_TRA", "0.5", "--remain_reads_ratio", "0.7", "--report_readid",
                  "--ignore_sequence", "--retain_work_dir", "--write_old_sigs", "-S", "HG002", "--gt_round", "100", "-include_bed", "x.bed"]):
        a = vars(parseArgs(argv))
        b = vars(cli.build_parser().parse_args(argv))
        assert a == b


def test_task_windows_float_bounds():
    stats = [("c1", 1000, 0, 1000), ("c2", 1, 0, 1)]
    lens = {"c1": 25000000, "c2": 5000}
    tasks, info = cli.task_windows(stats, lambda n: lens[n], 16, 10000000)
    assert info == [["c1", 25000000], ["c2", 5000]]
    assert tasks[-1] == ["c2", 0, 5000]
    c1 = [t for t in tasks if t[0] == "c1"]
    unit = 1001 / 16 / 10
    batch = 25000000 / (int(1000 / unit) + 1)   # coverage-balanced float window size, cuteSV:1034
    assert c1[0][1] == 0 and c1[0][2] == batch
    for x, y in zip(c1, c1[1:]):
        assert y[1] == x[2]                        # contiguous, float bounds preserved
    assert c1[-1][2] in (25000000, c1[-1][1] + batch)


def test_params_from_args():
    a = cli.build_parser().parse_args(["a", "b", "c", "d", "-s", "3", "--genotype"])
    p = cli.params_from_args(a)
    assert p.min_support == 3 and p.min_support_allele == 3 and p.genotype == 1 and p.bias_del == 200
