"""Wall time of the whole CLI (BAM -> VCF) on a synthetic coordinate-sorted BAM, with the stage split main_ctrl reports.
The BAM is written in parallel: every worker synthesises the records of its own contigs (synth.synth_alignments), encodes
and BGZF-compresses them; BGZF blocks concatenate, so the parent only adds the header and the EOF block.
python scripts/bench_cli.py [n_reads] [--workers W] [--keep path.bam] [--genotype]"""
import json
import multiprocessing as mp
import os
import struct
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PER_PART_CONTIGS = 3


def _part(job):
    k, n_reads = job
    import bam_writer
    from cutesv_b200 import synth
    reads, names, lens = synth.synth_alignments(1000 + k, n_reads=n_reads, n_contigs=PER_PART_CONTIGS, with_seq=True)
    ren = {nm: "%s_p%03d" % (nm, k) for nm in names}
    for r in reads:
        r.reference_name = ren[r.reference_name]
        r.query_name = "p%03d_%s" % (k, r.query_name)
        tags = []
        for tag, val in r.tags:
            if tag == "SA":
                val = ";".join((ren[e.split(",", 1)[0]] + "," + e.split(",", 1)[1]) for e in val.split(";") if e) + ";"
            tags.append((tag, val))
        r.tags = tags
    order = {ren[nm]: i for i, nm in enumerate(names)}
    reads.sort(key=lambda r: (order[r.reference_name], r.reference_start))
    base = k * PER_PART_CONTIGS
    stream = bytearray()
    mapped = [0] * PER_PART_CONTIGS
    for r in reads:
        stream += bam_writer._record(r, base + order[r.reference_name], False)
        mapped[order[r.reference_name]] += 1
    blocks = b"".join(bam_writer._bgzf_block(bytes(stream[o:o + 60000])) for o in range(0, len(stream), 60000))
    fasta = {ren[nm]: synth.pseudo_fasta_line(ren[nm], int(ln)) for nm, ln in zip(names, lens)}
    return k, [(ren[nm], int(ln)) for nm, ln in zip(names, lens)], mapped, blocks, len(reads), fasta


def write_bam_parallel(path, n_reads, workers):
    import bam_writer
    per = 20000
    parts = max(1, (n_reads + per - 1) // per)
    jobs = [(k, min(per, n_reads - k * per)) for k in range(parts)]
    with mp.Pool(min(workers, parts)) as pool:
        res = sorted(pool.map(_part, jobs, chunksize=1))
    contigs = [c for _, cs, _, _, _, _ in res for c in cs]
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % c for c in contigs)
    head = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(contigs)))
    for n, ln in contigs:
        head += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", ln)
    with open(path, "wb") as f:
        for o in range(0, len(head), 60000):
            f.write(bam_writer._bgzf_block(bytes(head[o:o + 60000])))
        for _, _, _, blocks, _, _ in res:
            f.write(blocks)
        f.write(bam_writer._bgzf_block(b""))
    with open(path + ".bai", "wb") as f:
        f.write(b"BAI\1" + struct.pack("<i", len(contigs)))
        for _, _, mapped, _, _, _ in res:
            for m in mapped:
                f.write(struct.pack("<i", 1) + struct.pack("<Ii", 37450, 2) + struct.pack("<QQQQ", 0, 0, m, 0) + struct.pack("<i", 0))
        f.write(struct.pack("<Q", 0))
    fa = os.path.splitext(path)[0] + ".fa"
    with open(fa, "w") as f:
        for _, _, _, _, _, fasta in res:
            for nm, line in fasta.items():
                f.write(">%s\n%s\n" % (nm, line))
    return sum(r[4] for r in res), fa


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if args else 100000
    workers = int(sys.argv[sys.argv.index("--workers") + 1]) if "--workers" in sys.argv else min(os.cpu_count() or 8, 64)
    keep = sys.argv[sys.argv.index("--keep") + 1] if "--keep" in sys.argv else None
    d = tempfile.mkdtemp()
    bam = keep or os.path.join(d, "bench.bam")
    t0 = time.perf_counter()
    n_written, fa = write_bam_parallel(bam, n, workers)
    t_write = time.perf_counter() - t0
    wd = os.path.join(d, "wd")
    os.mkdir(wd)
    out = os.path.join(d, "out.vcf")
    from cutesv_b200 import bamio, cli
    bamio.build()
    argv = [bam, fa, out, wd, "-s", "5", "--threads", "16", "--max_cluster_bias_INS", "100", "--diff_ratio_merging_INS", "0.3",
            "--max_cluster_bias_DEL", "100", "--diff_ratio_merging_DEL", "0.3"]
    if "--genotype" in sys.argv:
        argv.append("--genotype")
    a = cli.build_parser().parse_args(argv)
    engine = None
    if "--emulator" in sys.argv:   # authoring container (no GPU): the test-only pipeline emulator, to exercise this script
        sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
        from emul_engine import EmulEngine
        engine = EmulEngine()
    dev = {}
    if "--profile" in sys.argv and engine is None:   # device-side split of csv_extract_append (CUDA events) beside its wall time
        from cutesv_b200.engine import Engine
        engine = Engine(0)
        engine.set_profiling(True)
        inner = engine.extract

        def timed_extract(packed, append=False):
            w0 = time.perf_counter()
            r = inner(packed, append=append)
            dev["wall_s"] = dev.get("wall_s", 0.0) + time.perf_counter() - w0
            for k, v in engine.stage_ms().items():
                if v:
                    dev[k + "_ms"] = dev.get(k + "_ms", 0.0) + v
            dev["calls"] = dev.get("calls", 0) + 1
            dev["cigar_bytes"] = dev.get("cigar_bytes", 0) + int(packed["cigar"].nbytes)
            return r
        engine.extract = timed_extract
    t0 = time.perf_counter()
    cli.main_ctrl(a, argv, engine=engine)
    wall = time.perf_counter() - t0
    n_rec = sum(1 for line in open(out) if not line.startswith("#"))
    print(json.dumps(dict(n_records_in_bam=n_written, bam_mb=os.path.getsize(bam) / 1e6, bam_write_s=t_write, cli_wall_s=wall,
                          records_per_s=n_written / wall, vcf_records=n_rec, genotype="--genotype" in sys.argv,
                          stages_s={k: round(v, 4) for k, v in cli.main_ctrl.last_stages.items()},
                          extract_calls={k: (round(v, 3) if isinstance(v, float) else v) for k, v in dev.items()})))


if __name__ == "__main__":
    main()
