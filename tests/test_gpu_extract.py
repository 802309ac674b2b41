"""-m gpu: kernel (a) (CIGAR walk + split-read engine) through the C-ABI against the emulator of the
same logic (which tests/test_extract_cpu.py pins against the REAL reference's parse_read) and
against golden tuples from the reference; then extract -> cluster end to end on the device."""
import collections
import json
import os

import numpy as np
import pytest

import emul_lib
import golden_util
from cutesv_b200 import _abi, packing, synth
from oracle import compare_extract, compare_records, oracle_lib

pytestmark = pytest.mark.gpu


def _packet(seed, n=300, kind="short"):
    reads, names, lens = synth.synth_alignments_long(seed, n) if kind == "long" else synth.synth_alignments(seed, n)
    rnames = sorted(set(r.query_name for r in reads))
    rid = {nm: i for i, nm in enumerate(rnames)}
    cid = {nm: i for i, nm in enumerate(names)}
    return reads, names, lens, rnames, packing.pack_alignments(reads, cid, rid)


def _canon(ex):
    out = {}
    for t, cols in ex["sigs"].items():
        rows = list(zip(*[cols[k].tolist() for k in ("chrom", "a", "b", "read_id", "c")]))
        if t in ("DEL", "DUP"):
            rows = [r[:4] for r in rows]
        out[t] = collections.Counter(rows)
    r = ex["rows"]
    out["rows"] = collections.Counter(zip(r["chrom"].tolist(), r["start"].tolist(), r["end"].tolist(), r["read_id"].tolist(), r["is_primary"].tolist()))
    return out


@pytest.mark.parametrize("seed", range(12))
def test_extract_matches_emulator(engine, seed):
    reads, names, lens, rnames, pk = _packet(seed)
    rng = np.random.default_rng(seed)
    p = _abi.default_params(min_size=int(rng.choice([30, 50, 10])), max_size=int(rng.choice([-1, 100000, 2000])),
                            min_mapq=int(rng.choice([20, 0, 30])), max_split_parts=int(rng.choice([7, -1, 2, 3])),
                            min_read_len=int(rng.choice([500, 100])), min_siglength=int(rng.choice([10, 30])),
                            merge_del_threshold=int(rng.choice([0, 500])), merge_ins_threshold=int(rng.choice([100, 500, 0])))
    engine.set_params(p)
    engine.set_contigs(lens)
    engine.extract(pk)
    got = engine.fetch_extracted()
    ref = emul_lib.extract(p, pk)
    assert _canon(got) == _canon(ref)
    # INS sequences rebuilt from the piece table
    gc, gr = compare_extract.tuples_from_columns(got, names, rnames, lambda rec: reads[rec].query_sequence)
    rc, rr = compare_extract.tuples_from_columns(ref, names, rnames, lambda rec: reads[rec].query_sequence)
    assert not compare_extract.diff_extract(rc, rr, gc, gr)


@pytest.mark.parametrize("name", ["extract_s0", "extract_s1", "extract_s2", "extract_s3", "extract_s4", "extract_s5", "extract_s6",
                                  "extract_l0", "extract_l1", "extract_l2", "extract_l3", "extract_l4"])
def test_extract_matches_reference_golden(engine, name):
    """Tuples of the REAL reference's parse_read.  extract_l*: BASELINE config-5-shaped records (>= 10^4 CIGAR ops, clips on both
    ends, 2-6 SA segments in every strand pattern, MaxSize -1, chains of more merged insertions than the kernel buffers)."""
    meta = json.load(open(os.path.join(golden_util.GOLDEN, name + ".json")))
    reads, names, lens, rnames, pk = _packet(meta["seed"], meta["n_reads"], meta.get("kind", "short"))
    p = _abi.default_params(**meta["params"])
    engine.set_params(p)
    engine.set_contigs(lens)
    engine.extract(pk)
    got = engine.fetch_extracted()
    cigar_of = lambda rec: (pk["cigar"][pk["cigar_off"][rec]:pk["cigar_off"][rec + 1]], int(pk["ref_start"][rec]))
    gc, gr = compare_extract.tuples_from_columns(got, names, rnames, lambda rec: reads[rec].query_sequence, cigar_of,
                                                 (p.min_siglength, p.merge_ins_threshold))
    ref_c = {k: [tuple(t) for t in v] for k, v in meta["candidate"].items()}
    ref_r = [tuple(t) for t in meta["rows"]]
    assert not compare_extract.diff_extract(ref_c, ref_r, gc, gr)


def test_extract_then_cluster_on_device(engine):
    """Signatures never leave the GPU between extraction and clustering."""
    reads, names, lens, rnames, pk = _packet(5, 2500)
    p = _abi.default_params(min_support=2, genotype=1, min_mapq=0, min_read_len=100)
    engine.set_params(p)
    engine.set_contigs(lens)
    engine.extract(pk)
    ex = engine.fetch_extracted()
    engine.cluster_device(0x1F)
    got = engine.fetch()
    sigs = {t: dict(v) for t, v in ex["sigs"].items()}
    sigs["DEL"]["c"] = None
    sigs["DUP"]["c"] = None
    ref = oracle_lib.cluster(p, lens, sigs, ex["rows"])
    d = compare_records.diff_records(ref, got)
    assert not d, "\n".join(d[:3])
    assert len(got[0]) > 0


def _slice_packet(pk, lo, hi):
    """Records [lo, hi) of a packed alignment packet as a packet of their own."""
    out = {k: pk[k][lo:hi] for k in ("chrom", "ref_start", "ref_end", "flag", "mapq", "query_len", "read_id")}
    c0, c1 = int(pk["cigar_off"][lo]), int(pk["cigar_off"][hi])
    s0, s1 = int(pk["sa_off"][lo]), int(pk["sa_off"][hi])
    out["cigar_off"] = (pk["cigar_off"][lo:hi + 1] - c0).astype(np.int64)
    out["sa_off"] = (pk["sa_off"][lo:hi + 1] - s0).astype(np.int64)
    out["cigar"] = pk["cigar"][c0:c1]
    out["sa"] = {k: v[s0:s1] for k, v in pk["sa"].items()}
    return out


@pytest.mark.parametrize("seed,cuts", [(3, (0, 97, 98, 250, 400)), (8, (0, 1, 399, 400)), (11, (0, 200, 200, 400))])
def test_append_packets_equal_single_packet(engine, seed, cuts):
    """csv_extract_append over several packets (ragged, empty ones included) leaves the same device-resident signatures,
    INS sequences and reads rows as ONE csv_extract over all records, and the clustering on them is identical."""
    reads, names, lens, rnames, pk = _packet(seed, 400)
    p = _abi.default_params(min_support=2, genotype=1, min_mapq=0, min_read_len=100)
    engine.set_params(p)
    engine.set_contigs(lens)
    engine.extract(pk)
    one = engine.fetch_extracted()
    engine.cluster_device(0x1F)
    res_one = engine.fetch()
    engine.extract_reset()
    seqs = []
    base = 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        r = engine.extract(_slice_packet(pk, lo, hi), append=True)
        n_new = r["counts"]["INS"] - r["first"]["INS"]
        po, pc, pieces = engine.fetch_ins_pieces(r["first"]["INS"], n_new, r["first_pieces"], r["n_pieces"] - r["first_pieces"])
        assert (pieces[:, 0] >= base).all() and (pieces[:, 0] < base + (hi - lo)).all() if len(pieces) else True   # global record indices
        for i in range(n_new):
            seqs.append(packing.ins_sequence(pieces, int(po[i]), int(pc[i]), lambda rec: reads[rec].query_sequence))
        base += hi - lo
    many = engine.fetch_extracted()
    assert _canon(many) == _canon(one)
    oc, orr = compare_extract.tuples_from_columns(one, names, rnames, lambda rec: reads[rec].query_sequence)
    mc, mr = compare_extract.tuples_from_columns(many, names, rnames, lambda rec: reads[rec].query_sequence)
    assert not compare_extract.diff_extract(oc, orr, mc, mr)
    # the per-packet strings are the strings of the final INS rows, in device order
    s = many["sigs"]["INS"]
    assert len(seqs) == len(s["chrom"]) and [len(x) for x in seqs] == s["c"].tolist()
    engine.cluster_device(0x1F)
    res_many = engine.fetch()
    def canon(res):
        c, g, nm = res
        return sorted((int(x["svtype"]), int(x["chrom"]), int(x["pos"]), int(x["len"]), int(x["support"]), int(y["dr"]), int(y["gt"]),
                       tuple(sorted(nm[x["names_off"]:x["names_off"] + x["names_cnt"]].tolist()))) for x, y in zip(c, g))
    assert canon(res_many) == canon(res_one) and len(res_one[0]) > 0


def test_remap_read_ids_and_swap_rows(engine):
    reads, names, lens, rnames, pk = _packet(2, 300)
    engine.set_params(_abi.default_params(min_mapq=0, min_read_len=100))
    engine.set_contigs(lens)
    engine.extract(pk)
    before = engine.fetch_extracted()
    n_ids = int(max(int(v["read_id"].max()) for v in before["sigs"].values() if len(v["read_id"]))) + 1
    n_ids = max(n_ids, int(before["rows"]["read_id"].max()) + 1)
    rank = np.random.default_rng(1).permutation(n_ids).astype(np.int32)
    engine.remap_read_ids(rank)
    after = engine.fetch_extracted()
    for t in before["sigs"]:
        assert np.array_equal(after["sigs"][t]["read_id"], rank[before["sigs"][t]["read_id"]])
    assert np.array_equal(after["rows"]["read_id"], rank[before["rows"]["read_id"]])
    n = len(after["sigs"]["INS"]["chrom"])
    assert n >= 3
    engine.swap_ins_rows([(0, 2), (2, 1)])
    sw = engine.fetch_extracted()
    perm = [2, 0, 1] + list(range(3, n))   # row 0 <- old 2, row 2 <- old 0 then swapped with row 1
    for k in ("chrom", "a", "b", "read_id", "c"):
        assert np.array_equal(sw["sigs"]["INS"][k], after["sigs"]["INS"][k][perm]), k
    assert np.array_equal(sw["piece_off"], after["piece_off"][perm]) and np.array_equal(sw["piece_cnt"], after["piece_cnt"][perm])


def _many_segment_packet():
    """Three records; the middle one carries 70 SA segments (more than the 64 the split-read engine holds)."""
    reads, names, lens = synth.synth_alignments(21, 3)
    for r in reads:
        r.flag, r.mapq = 0, 60
    r = reads[1]
    L = r.query_length
    ents = []
    for j in range(70):
        a = (j * L) // 71
        b = ((j + 1) * L) // 71
        ents.append("%s,%d,+,%dS%dM%dS,60,1" % (r.reference_name, r.reference_end + 50 * j + 1, a, max(b - a, 1), max(L - b, 0)))
    r.tags = [("NM", 1), ("SA", ";".join(ents) + ";")]
    rnames = sorted(set(x.query_name for x in reads))
    pk = packing.pack_alignments(reads, {nm: i for i, nm in enumerate(names)}, {nm: i for i, nm in enumerate(rnames)})
    return reads, names, lens, rnames, pk


def test_more_than_64_segments_is_counted_not_fatal(engine):
    """--max_split_parts -1 with a record of 70 segments: round 1 failed the whole call (CSV_E_INPUT); now the record's split-read
    analysis is skipped and counted, everything else (its CIGAR signatures, the other records) is extracted as usual."""
    reads, names, lens, rnames, pk = _many_segment_packet()
    p = _abi.default_params(max_split_parts=-1, min_mapq=0, min_read_len=100)
    engine.set_params(p)
    engine.set_contigs(lens)
    engine.extract(pk)
    got = engine.fetch_extracted()
    assert engine.extract_skipped() == 1
    ref = emul_lib.extract(p, pk)
    assert _canon(got) == _canon(ref)
    # the same packet without the long record gives the same split signatures: only that record's were skipped
    p7 = _abi.default_params(max_split_parts=7, min_mapq=0, min_read_len=100)   # the reference's default drops the record's split analysis too
    engine.set_params(p7)
    engine.extract(pk)
    assert engine.extract_skipped() == 0
    assert _canon(engine.fetch_extracted())["DEL"] == _canon(got)["DEL"]
