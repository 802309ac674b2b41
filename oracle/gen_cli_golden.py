"""Runs the REAL reference CLI core (main_ctrl, cuteSV:992-1248) end to end on the synthetic BAM data
set through the test-only fake pysam, and commits the VCF body as tests/golden/cli_dataset1.json.
Authoring container only:  python -m oracle.gen_cli_golden"""
import json
import os
import pickle
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "fake_pysam"))

FLAGS = ["--genotype", "-s", "5", "--threads", "4", "--max_cluster_bias_INS", "100", "--diff_ratio_merging_INS", "0.3",
         "--max_cluster_bias_DEL", "100", "--diff_ratio_merging_DEL", "0.3"]


def materialise(d, seed=1, double_ins=0.0):
    from cutesv_b200 import synth
    ds, fasta = synth.synth_bam_dataset(seed, double_ins=double_ins)
    bam = os.path.join(d, "x.bam")
    with open(bam, "wb") as f:
        pickle.dump(ds, f)
    fa = os.path.join(d, "ref.fa")
    with open(fa, "w") as f:
        f.write("".join(">%s\n%s\n" % (k, v) for k, v in fasta.items()))
    wd = os.path.join(d, "wd")
    os.mkdir(wd)
    return bam, fa, os.path.join(d, "out.vcf"), wd


BED_ROWS = [("chrA", 20000, 50000), ("chrA", 70000, 71000), ("chrA", 100500, 104000), ("chrB", 10000, 60000), ("chrB", 88000, 90500)]


def write_bed(d):
    """-include_bed input of the third golden (regions straddle task-window borders)."""
    bed = os.path.join(d, "inc.bed")
    with open(bed, "w") as f:
        f.write("".join("%s\t%d\t%d\n" % r for r in BED_ROWS))
    return bed


# BASELINE.json configs[0]: BED-driven chr22 INS+DEL, ~5k reads, reference CPU path --threads 4, ONT preset
FLAGS_CONFIG1 = ["--genotype", "-s", "5", "--threads", "4", "--max_cluster_bias_INS", "100", "--diff_ratio_merging_INS", "0.3",
                 "--max_cluster_bias_DEL", "100", "--diff_ratio_merging_DEL", "0.3"]


def materialise_config1(d):
    from cutesv_b200 import synth
    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "sim_chr22_loci.json")))
    ds = synth.synth_config1_dataset(meta["loci"], meta["contig"], meta["contig_len"])
    bam = os.path.join(d, "c1.bam")
    with open(bam, "wb") as f:
        pickle.dump(ds, f)
    fa = os.path.join(d, "c1.fa")
    with open(fa, "w") as f:
        f.write(">%s\n%s\n" % (meta["contig"], synth.pseudo_fasta_line(meta["contig"], meta["contig_len"] + 16)))
    wd = os.path.join(d, "wd1")
    os.mkdir(wd)
    return bam, fa, os.path.join(d, "c1.vcf"), wd


# further flag sets on data set 1 (VCF formatting branches: RNAMES, no genotype / no sequences, size limits, presets)
EXTRA_FLAG_SETS = {
    "cli_dataset1_hifi_readid": ["--genotype", "--report_readid", "-s", "3", "-l", "50", "-L", "-1", "--max_cluster_bias_INS", "1000",
                                 "--diff_ratio_merging_INS", "0.9", "--max_cluster_bias_DEL", "1000", "--diff_ratio_merging_DEL", "0.5",
                                 "-q", "10", "-r", "1000", "--remain_reads_ratio", "0.8", "--threads", "3", "-S", "SAMPLE7"],
    "cli_dataset1_nogt_noseq": ["-s", "4", "--ignore_sequence", "-L", "400", "--max_cluster_bias_INS", "100", "--diff_ratio_merging_INS", "0.3",
                                "--max_cluster_bias_DEL", "100", "--diff_ratio_merging_DEL", "0.3", "--threads", "2", "-b", "30000",
                                "-md", "500", "-mi", "500", "-sl", "20", "-p", "3"],
}


# INS ties: half of the INS-carrying reads report the insertion as two equal-length I ops at the same position; -mi -1 keeps them
# apart, so the signatures tie on (chr, int(pos), len, name) and the reference orders them by their sequence strings (cuteSV:774)
TIES = dict(seed=2, double_ins=0.5, flags=FLAGS + ["-mi", "-1", "--report_readid"])


def main_ties():
    import pysam  # noqa: F401
    from oracle import ref_harness
    m = ref_harness.modules()
    from cuteSV.cuteSV_Description import parseArgs
    d = tempfile.mkdtemp()
    bam, fa, out, wd = materialise(d, TIES["seed"], TIES["double_ins"])
    argv = [bam, fa, out, wd] + TIES["flags"]
    m["main"].main_ctrl(parseArgs(argv), argv)
    lines = [l for l in open(out) if not l.startswith("##")]
    with open(os.path.join(ROOT, "tests", "golden", "cli_dataset2_ins_ties.json"), "w") as f:
        json.dump(dict(flags=TIES["flags"], seed=TIES["seed"], double_ins=TIES["double_ins"], lines=lines), f)
    print("dataset2 (INS ties):", len(lines) - 1, "records")


def main():
    import pysam  # noqa: F401  (the fake one, first on sys.path)
    from oracle import ref_harness
    m = ref_harness.modules()
    from cuteSV.cuteSV_Description import parseArgs
    d = tempfile.mkdtemp()
    bam, fa, out, wd = materialise(d)
    argv = [bam, fa, out, wd] + FLAGS
    m["main"].main_ctrl(parseArgs(argv), argv)
    lines = [l for l in open(out) if not l.startswith("##")]
    with open(os.path.join(ROOT, "tests", "golden", "cli_dataset1.json"), "w") as f:
        json.dump(dict(flags=FLAGS, lines=lines), f)
    print(len(lines) - 1, "records")
    bam, fa, out, wd = materialise_config1(d)
    argv = [bam, fa, out, wd] + FLAGS_CONFIG1
    m["main"].main_ctrl(parseArgs(argv), argv)
    lines = [l for l in open(out) if not l.startswith("##")]
    with open(os.path.join(ROOT, "tests", "golden", "cli_config1.json"), "w") as f:
        json.dump(dict(flags=FLAGS_CONFIG1, lines=lines), f)
    print("config1:", len(lines) - 1, "records")
    d2 = tempfile.mkdtemp()
    bam, fa, out, wd = materialise(d2)
    argv = [bam, fa, out, wd] + FLAGS + ["-include_bed", write_bed(d2)]
    m["main"].main_ctrl(parseArgs(argv), argv)
    lines = [l for l in open(out) if not l.startswith("##")]
    with open(os.path.join(ROOT, "tests", "golden", "cli_dataset1_bed.json"), "w") as f:
        json.dump(dict(flags=FLAGS, lines=lines), f)
    print("dataset1 + include_bed:", len(lines) - 1, "records")
    for name, flags in EXTRA_FLAG_SETS.items():
        dx = tempfile.mkdtemp()
        bam, fa, out, wd = materialise(dx)
        argv = [bam, fa, out, wd] + flags
        m["main"].main_ctrl(parseArgs(argv), argv)
        lines = [l for l in open(out) if not l.startswith("##")]
        with open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w") as f:
            json.dump(dict(flags=flags, lines=lines), f)
        print(name, len(lines) - 1, "records")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ties":
        main_ties()
    else:
        main()
