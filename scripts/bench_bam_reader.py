"""Host data-loader timing: native BAM decoder (bamio) on a synthetic BAM; optional comparison with the
per-read Python packer on the same records.
python scripts/bench_bam_reader.py [n_reads] [--python-packer] [--keep path.bam]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bam_writer  # noqa: E402
from cutesv_b200 import bamio, packing, synth  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if args else 20000
    keep = sys.argv[sys.argv.index("--keep") + 1] if "--keep" in sys.argv else None
    bamio.build()
    path = keep or os.path.join(tempfile.mkdtemp(), "t.bam")
    reads = None
    names = lens = None
    if not (keep and os.path.exists(keep)):
        reads, names, lens = synth.synth_alignments(11, n_reads=n, with_seq=True)
        order = {nm: i for i, nm in enumerate(names)}
        reads.sort(key=lambda r: (order[r.reference_name], r.reference_start))
        bam_writer.write_bam(path, list(zip(names, (int(x) for x in lens))), reads)
    size = os.path.getsize(path)
    for threads in (1, 2, 4, 8, 16):
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            rd = bamio.BamReader(path, threads=threads)
            rd.set_chrom_ids({nm: i for i, nm in enumerate(sorted(rd.references))})
            k = ops = 0
            while True:
                pk = rd.next_packet(50000, copy=False)
                if pk is None:
                    break
                k += len(pk["chrom"])
                ops += len(pk["cigar"])
            rd.name_ranks()
            dt = time.perf_counter() - t0
            rd.close()
            best = dt if best is None else min(best, dt)
        print("native threads=%2d: %d records, %d CIGAR ops, %.1f MB bam in %.3f s (best of 3) -> %.0f records/s, %.1f MB/s compressed"
              % (threads, k, ops, size / 1e6, best, k / best, size / 1e6 / best))
    if "--python-packer" in sys.argv and reads is not None:
        class _Ids(dict):
            def __missing__(self, key):
                self[key] = len(self)
                return self[key]
        t0 = time.perf_counter()
        packing.pack_alignments(reads, {nm: i for i, nm in enumerate(sorted(names))}, _Ids())
        dt = time.perf_counter() - t0
        print("python packer (records already decoded by the caller): %.3f s -> %.0f records/s" % (dt, len(reads) / dt))


if __name__ == "__main__":
    main()
