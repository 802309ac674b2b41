"""TEST-ONLY stand-in for pysam backed by a pickle of synthetic alignment records
(cutesv_b200.synth.synth_bam_dataset).  Lets the CLI shell (and, in the authoring container, the
reference's own main_ctrl) run end to end without htslib."""
import pickle

CMATCH, CINS, CDEL, CREF_SKIP, CSOFT_CLIP, CHARD_CLIP, CPAD, CEQUAL, CDIFF, CBACK = range(10)


class AlignmentFile(object):
    def __init__(self, path, mode=None, reference_filename=None):
        with open(path, "rb") as f:
            d = pickle.load(f)
        self.contigs = d["contigs"]
        self.reads = d["reads"]
        self.by_chrom = {}
        for r in self.reads:
            self.by_chrom.setdefault(r.reference_name, []).append(r)
        for v in self.by_chrom.values():
            v.sort(key=lambda r: r.reference_start)

    def get_index_statistics(self):
        return [(n, len(self.by_chrom.get(n, [])), 0, len(self.by_chrom.get(n, []))) for n, _ in self.contigs]

    def get_reference_length(self, name):
        return dict(self.contigs)[name]

    def fetch(self, chrom, start=None, end=None):
        for r in self.by_chrom.get(chrom, []):
            if (end is None or r.reference_start < end) and (start is None or r.reference_end > start):
                yield r

    def close(self):
        pass


class FastaFile(object):
    def __init__(self, path):
        self.seqs, name, chunks = {}, None, []
        with open(path) as f:
            for line in f:
                if line.startswith(">"):
                    if name is not None:
                        self.seqs[name] = "".join(chunks)
                    name, chunks = line[1:].split()[0], []
                else:
                    chunks.append(line.strip())
        if name is not None:
            self.seqs[name] = "".join(chunks)

    def fetch(self, chrom):
        return self.seqs[chrom]

    def close(self):
        pass
