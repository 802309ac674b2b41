"""Host data-loader timing: native BAM decoder (bamio) vs the per-read Python packer on the same records.
python scripts/bench_bam_reader.py [n_reads]"""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    bamio.build()
    reads, names, lens = synth.synth_alignments(11, n_reads=n, with_seq=True)
    order = {nm: i for i, nm in enumerate(names)}
    reads.sort(key=lambda r: (order[r.reference_name], r.reference_start))
    d = tempfile.mkdtemp()
    path = os.path.join(d, "t.bam")
    bam_writer.write_bam(path, list(zip(names, (int(x) for x in lens))), reads)
    size = os.path.getsize(path)
    chrom_id = {nm: i for i, nm in enumerate(sorted(names))}
    for threads in (1, 4, 16):
        t0 = time.perf_counter()
        rd = bamio.BamReader(path, threads=threads)
        rd.set_chrom_ids(chrom_id)
        k = ops = 0
        while True:
            pk = rd.next_packet(50000)
            if pk is None:
                break
            k += len(pk["chrom"])
            ops += len(pk["cigar"])
        rank = rd.name_ranks()
        dt = time.perf_counter() - t0
        rd.close()
        print("native threads=%2d: %d records, %d CIGAR ops, %.1f MB bam in %.3f s -> %.0f records/s, %.1f MB/s compressed"
              % (threads, k, ops, size / 1e6, dt, k / dt, size / 1e6 / dt))
    t0 = time.perf_counter()
    ids = {}
    class _Ids(dict):
        def __missing__(self, key):
            self[key] = len(self)
            return self[key]
    packing.pack_alignments(reads, chrom_id, _Ids())
    dt = time.perf_counter() - t0
    print("python packer (records already decoded by the caller): %.3f s -> %.0f records/s" % (dt, len(reads) / dt))


if __name__ == "__main__":
    main()
