"""More extraction goldens from the REAL reference's parse_read (cuteSV:606-681): BASELINE config-5-shaped records
(>= 10^4 CIGAR ops, clips, chained insertions, 2-6 SA segments in every strand pattern) and further flag settings on
the short-read packets.  Only the reference's OUTPUT tuples are committed; the inputs are regenerated from the seed.
Authoring container only:  python -m oracle.gen_extract_golden"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cutesv_b200 import _abi, synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, generator, seed, n_reads, flags)
CASES = [
    ("extract_l0", "long", 100, 12, dict()),
    ("extract_l1", "long", 101, 12, dict(max_size=-1, max_split_parts=-1, min_mapq=0)),
    ("extract_l2", "long", 102, 12, dict(min_size=50, min_siglength=12, merge_ins_threshold=40, merge_del_threshold=200, max_split_parts=4)),
    ("extract_l3", "long", 103, 10, dict(max_size=2000, min_read_len=100, min_mapq=30, merge_ins_threshold=500)),
    ("extract_l4", "long", 104, 10, dict(max_size=-1, min_size=10, min_siglength=10, max_split_parts=7)),
    ("extract_s3", "short", 3, 300, dict(max_size=-1, max_split_parts=-1, min_mapq=0, min_read_len=100)),
    ("extract_s4", "short", 4, 300, dict(min_size=10, min_siglength=30, merge_del_threshold=500)),
    ("extract_s5", "short", 5, 300, dict(max_size=2000, max_split_parts=2, min_mapq=30)),
    ("extract_s6", "short", 6, 400, dict(merge_ins_threshold=0, merge_del_threshold=0, min_siglength=10, max_split_parts=5)),
]


def reads_of(kind, seed, n):
    return synth.synth_alignments_long(seed, n) if kind == "long" else synth.synth_alignments(seed, n)


def main():
    for name, kind, seed, n, kw in CASES:
        reads, _, _ = reads_of(kind, seed, n)
        c, r = ref_harness.run_parse_reads(reads, _abi.default_params(**kw))
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(dict(kind=kind, seed=seed, n_reads=n, params=kw, candidate={k: [list(t) for t in v] for k, v in c.items()},
                           rows=[list(t) for t in r]), f)
        print(name, {k: len(v) for k, v in c.items()}, "rows", len(r), "max ops", max(len(x.cigartuples) for x in reads))


if __name__ == "__main__":
    main()
