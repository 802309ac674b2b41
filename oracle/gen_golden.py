"""Generates tests/golden/* by running the UNMODIFIED reference (imported from /root/reference/src
with pysam/cigar/Bio stubbed, see ref_harness.py) on seeded inputs.  Run in the authoring container:

    python -m oracle.gen_golden

The fixtures pin the oracle (tests/test_oracle_golden.py) and, on the GPU box where the
reference does not exist, the CUDA path (tests/test_gpu_golden.py).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cutesv_b200 import _abi, synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = [("adv%03d" % s, ("adversarial", s)) for s in (0, 1, 2, 3, 5, 8, 13, 21, 34, 55, 89, 144)] + [
    ("cfg2_s0p002", ("config", 2, 0.002)), ("cfg3_s0p004", ("config", 3, 0.004)), ("cfg5_s0p001", ("config", 5, 0.001))] + [
    ("sweep%02d" % s, ("sweep", s)) for s in range(12)]   # random flag settings (synth.random_params) x adversarial inputs


def make_case(spec):
    if spec[0] == "adversarial":
        return synth.adversarial(spec[1], max_sigs=160)
    if spec[0] == "sweep":
        cfg = synth.adversarial(3000 + spec[1], max_sigs=160)
        cfg["params"] = dict(cfg["params"], **synth.random_params(spec[1]))
        return cfg
    return synth.make_config(spec[1], spec[2])


def params_dict(p):
    return {f[0]: getattr(p, f[0]) for f in p._fields_ if f[0] != "reserved"}


def golden_util_load(name):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util
    return golden_util.load_case(name)


def main():
    os.makedirs(OUT, exist_ok=True)
    m = ref_harness.modules()
    index = []
    for name, spec in CASES:
        cfg = make_case(spec)
        p = _abi.default_params(**cfg["params"])
        rows = ref_harness.run_reference(cfg["sigs"], cfg["reads"], cfg["names"], synth.read_name, p)
        arrays = {"lens": cfg["lens"]}
        for t, cols in cfg["sigs"].items():
            for k, v in cols.items():
                if v is not None:
                    arrays["sig_%s_%s" % (t, k)] = v
        for k, v in cfg["reads"].items():
            arrays["reads_" + k] = v
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
        exp = {"%s|%s" % k: v for k, v in rows.items() if v}
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump({"spec": list(spec), "names": cfg["names"], "params": params_dict(p), "rows": exp}, f)
        index.append(name)
        print(name, sum(len(v) for v in exp.values()), "rows")
    # cal_GL over its whole reachable domain + rescale cases (cuteSV_genotype.py:25-56)
    g = m["genotype"]
    tab = []
    for c0 in range(0, 131):
        for c1 in range(0, 131):
            if c0 + c1 == 0:
                continue
            r = g.cal_GL(c0, c1)
            tab.append([c0, c1, r[0], r[1], int(r[2]), str(r[3])])
    for c0, c1 in ((500, 7), (7, 500), (1000, 1000), (99, 2), (2, 99), (12345, 1)):
        r = g.cal_GL(c0, c1)
        tab.append([c0, c1, r[0], r[1], int(r[2]), str(r[3])])
    with open(os.path.join(OUT, "cal_gl.json"), "w") as f:
        json.dump(tab, f)
    # cal_CIPOS(np.std(list), n) known answers (resolveINDEL.py:191-194; cuteSV_genotype.py:58-60)
    rng = np.random.default_rng(11)
    kat = []
    for n in (1, 2, 3, 7, 8, 9, 31, 64, 127, 128, 129, 200, 257, 1000, 2921, 3541, 5000):
        v = [int(x) for x in (rng.integers(0, 2000, n) + int(rng.integers(0, 200000000)))]
        kat.append({"v": v, "cipos": g.cal_CIPOS(np.std(v), len(v)), "std_hex": float(np.std(v)).hex()})
    with open(os.path.join(OUT, "cipos_kat.json"), "w") as f:
        json.dump(kat, f)
    # extraction: tuples of the reference's parse_read on seeded synthetic alignment packets
    ex_cases = [("extract_s0", 0, 300, dict()),
                ("extract_s1", 1, 300, dict(min_mapq=0, max_split_parts=-1, merge_del_threshold=500, merge_ins_threshold=500, min_read_len=100)),
                ("extract_s2", 2, 400, dict(min_size=50, max_size=-1, min_siglength=30, max_split_parts=3))]
    for name, seed, n, kw in ex_cases:
        reads, _, _ = synth.synth_alignments(seed, n)
        c, r = ref_harness.run_parse_reads(reads, _abi.default_params(**kw))
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(dict(seed=seed, n_reads=n, params=kw, candidate={k: [list(t) for t in v] for k, v in c.items()},
                           rows=[list(t) for t in r]), f)
    # TRA genotyping (call_gt re-opens the BAM): the reference on a fake-pysam BAM built from the reads table
    import tempfile
    for name in ("adv034", "adv144", "cfg3_s0p004", "adv013"):
        case = golden_util_load(name)
        if not case["params"].genotype:
            continue
        aln = ref_harness.sorted_alignments(case["reads"])
        with tempfile.TemporaryDirectory() as d:
            bam = os.path.join(d, "aln.bam")
            ref_harness.write_fake_bam(bam, aln, case["names"], case["lens"], synth.read_name)
            rows = ref_harness.run_reference(case["sigs"], case["reads"], case["names"], synth.read_name, case["params"], types_=("TRA",),
                                             tra_bam=bam)
        with open(os.path.join(OUT, "tragt_%s.json" % name), "w") as f:
            json.dump({"%s|%s" % k: v for k, v in rows.items() if v}, f)
        print("tragt", name, sum(len(v) for v in rows.values()))
    # the parity sink: VCF records the reference's generate_output + SVID loop produce from its own rows
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util
    import vcf_util
    for name in ("adv034", "adv144", "cfg3_s0p004", "cfg2_s0p002"):
        case = golden_util.load_case(name)
        byc = vcf_util.rows_by_chrom(case["rows"])
        refseq = vcf_util.synthetic_reference(case["names"], [min(int(x), 6000000) for x in case["lens"]])
        lines = ref_harness.reference_vcf_lines(byc, refseq, bool(case["params"].genotype))
        with open(os.path.join(OUT, "vcf_%s.json" % name), "w") as f:
            json.dump(lines, f)
        print("vcf", name, len(lines))
    with open(os.path.join(OUT, "index.json"), "w") as f:
        json.dump(index, f)
    print("cal_gl", len(tab), "cipos", len(kat))


if __name__ == "__main__":
    main()
