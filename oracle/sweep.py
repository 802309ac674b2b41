"""Offline parity sweeps over fresh seeds (authoring container: the live reference at /root/reference is imported).
    python -m oracle.sweep cluster LO HI     kernel-template emulator (tests/emul) vs the C oracle, adversarial inputs x {preset, random flags}
    python -m oracle.sweep reference LO HI   the C oracle vs the live reference's rows on the same inputs
    python -m oracle.sweep extract LO HI [long]   emulator of kernel (a) vs the live reference's parse_read
    python -m oracle.sweep cli LO HI         the reference's own main_ctrl (fake pysam) vs cutesv_b200.cli on a real BAM of the same records
                                             (native decoder + emulator engine), VCF bodies; 4 flag sets, INS ties on every third seed
Prints every mismatch and a summary line.  Two mismatch classes of `reference` are documented deviations:
INS rows tying on (contig, int(pos), len, read) in arbitrary input order, and the reference's own KeyError in overlap_cover
on a zero-width window (zero-length DUP, only with -l 0)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))

import numpy as np  # noqa: E402

from cutesv_b200 import _abi, packing, synth  # noqa: E402
from oracle import compare, compare_extract, compare_records, oracle_lib, ref_harness  # noqa: E402


def _cases(seed):
    for k in range(2):
        cfg = synth.adversarial(seed, max_sigs=160)
        if k:
            cfg["params"] = dict(cfg["params"], **synth.random_params(seed))
        yield k, cfg, _abi.default_params(**cfg["params"])


def cluster(seed):
    import emul_lib
    for k, cfg, p in _cases(seed):
        ref = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"])
        got = emul_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"])
        yield k, compare_records.diff_records(ref, got)


def reference(seed):
    import golden_util
    for k, cfg, p in _cases(seed):
        rows = ref_harness.run_reference(cfg["sigs"], cfg["reads"], cfg["names"], synth.read_name, p)
        exp = {kk: v for kk, v in rows.items() if v}
        res = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"])
        yield k, compare.diff_rows(exp, golden_util.to_rows(dict(sigs=cfg["sigs"], names=cfg["names"], params=p), res))


def extract(seed, kind="short"):
    import emul_lib
    rng = np.random.default_rng(seed)
    p = _abi.default_params(min_size=int(rng.choice([30, 50, 10])), max_size=int(rng.choice([-1, 100000, 2000])),
                            min_mapq=int(rng.choice([20, 0, 30])), max_split_parts=int(rng.choice([7, -1, 2, 3])),
                            min_read_len=int(rng.choice([500, 100])), min_siglength=int(rng.choice([10, 30])),
                            merge_del_threshold=int(rng.choice([0, 500])), merge_ins_threshold=int(rng.choice([100, 500, 0])))
    reads, names, _ = synth.synth_alignments_long(seed, 6) if kind == "long" else synth.synth_alignments(seed, 120)
    rnames = sorted(set(r.query_name for r in reads))
    pk = packing.pack_alignments(reads, {nm: i for i, nm in enumerate(names)}, {nm: i for i, nm in enumerate(rnames)})
    ex = emul_lib.extract(p, pk)
    cigar_of = lambda rec: (pk["cigar"][pk["cigar_off"][rec]:pk["cigar_off"][rec + 1]], int(pk["ref_start"][rec]))  # noqa: E731
    gc, gr = compare_extract.tuples_from_columns(ex, names, rnames, lambda rec: reads[rec].query_sequence, cigar_of,
                                                 (p.min_siglength, p.merge_ins_threshold))
    ref_c, ref_r = ref_harness.run_parse_reads(reads, p)
    yield 0, compare_extract.diff_extract(ref_c, ref_r, gc, gr)


def cli(seed):
    import pickle
    import re
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, "tests", "fake_pysam"))
    import pysam  # noqa: F401  (the fake one)
    import bam_writer
    from emul_engine import EmulEngine
    from cutesv_b200 import bamio
    from cutesv_b200 import cli as our_cli
    from oracle import gen_cli_golden
    m = ref_harness.modules()
    from cuteSV.cuteSV_Description import parseArgs
    bamio.build()
    flag_sets = [gen_cli_golden.FLAGS, gen_cli_golden.FLAGS + ["-mi", "-1", "--report_readid"],
                 gen_cli_golden.EXTRA_FLAG_SETS["cli_dataset1_hifi_readid"], gen_cli_golden.EXTRA_FLAG_SETS["cli_dataset1_nogt_noseq"]]
    flags = flag_sets[seed % len(flag_sets)]
    d = tempfile.mkdtemp()
    bam, fa, out, wd = gen_cli_golden.materialise(d, seed, 0.5 if seed % 3 == 0 else 0.0)
    argv = [bam, fa, out, wd] + flags
    m["main"].main_ctrl(parseArgs(argv), argv)
    ds = pickle.load(open(bam, "rb"))
    order = {nm: i for i, (nm, _) in enumerate(ds["contigs"])}
    real = os.path.join(d, "real.bam")
    bam_writer.write_bam(real, ds["contigs"], sorted(ds["reads"], key=lambda r: (order[r.reference_name], r.reference_start)), extra_unmapped=2)
    d2 = tempfile.mkdtemp()
    os.mkdir(os.path.join(d2, "wd"))
    argv2 = [real, fa, os.path.join(d2, "o.vcf"), os.path.join(d2, "wd")] + flags
    our_cli.main_ctrl(our_cli.build_parser().parse_args(argv2), argv2, engine=EmulEngine())

    def norm(line):   # RNAMES of DUP / TRA records come from Python set iteration in the reference (hash order)
        if "SVTYPE=BND" in line or "SVTYPE=DUP" in line:
            return re.sub(r"RNAMES=([^;\t]*)", lambda mm: "RNAMES=" + ",".join(sorted(mm.group(1).split(","))), line)
        return line
    ref = [norm(x) for x in open(out) if not x.startswith("##")]
    got = [norm(x) for x in open(argv2[2]) if not x.startswith("##")]
    yield 0, ([] if ref == got else ["%d vs %d lines; first difference: %r" % (len(ref), len(got), [(a, b) for a, b in zip(ref, got) if a != b][:1])])


def main():
    mode, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    extra = sys.argv[4:]
    fn = {"cluster": cluster, "reference": reference, "extract": extract, "cli": cli}[mode]
    bad = n = 0
    t0 = time.time()
    for seed in range(lo, hi):
        try:
            results = list(fn(seed, *extra))
        except Exception as e:   # the reference itself raises on some degenerate inputs
            results = [(-1, ["EXC %r" % (e,)])]
        for k, d in results:
            n += 1
            if d:
                bad += 1
                print("MISMATCH seed %d variant %d: %s" % (seed, k, str(d)[:400]))
                sys.stdout.flush()
    print("%s sweep [%d, %d): %d cases, %d mismatches, %.0f s" % (mode, lo, hi, n, bad, time.time() - t0))


if __name__ == "__main__":
    main()
