"""packing.ins_block_from_packed (vectorised INS sequences from BAM's 4-bit bases) == packing.ins_sequence (per-signature
Python slices, the reference's semantics: cuteSV:639 and the split-read slices of cuteSV:231-455), and InsStore's indexing."""
import numpy as np

from cutesv_b200 import packing

_NIB = "=ACMGRSVTWYHKDBN"


def _pack(queries):
    seq4, lo, hi = [], [], []
    for q in queries:
        lo.append(len(seq4))
        codes = [_NIB.index(c) for c in q]
        if len(codes) & 1:
            codes.append(0)
        seq4.extend((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))
        hi.append(len(seq4))
    return np.array(seq4, dtype=np.uint8), np.array(lo, dtype=np.int64), np.array(hi, dtype=np.int64)


def _case(seed, n_rec=40, n_sig=300):
    rng = np.random.default_rng(seed)
    queries = ["".join(rng.choice(list("ACGTN"), int(rng.integers(0, 90)))) for _ in range(n_rec)]
    queries[3] = ""            # '*' in the BAM: nothing stored
    qlen = np.array([len(q) for q in queries], dtype=np.int32)
    qlen[3] = 57               # l_seq may still be set when the bases are absent -> treated as missing
    pieces, po, pc = [], [], []
    for _ in range(n_sig):
        k = int(rng.choice([1, 1, 1, 2, 3, 0]))
        po.append(len(pieces)); pc.append(k)
        for _ in range(k):
            rec = int(rng.integers(0, n_rec))
            L = int(qlen[rec])
            a = int(rng.integers(-5, L + 10))
            b = int(rng.integers(-5, L + 10))
            rc = int(rng.choice([0, 0, 0, 0, 1]))
            pieces.append((rec, a, b, rc))
    return queries, qlen, np.array(pieces, dtype=np.int32).reshape(-1, 4), np.array(po, dtype=np.int32), np.array(pc, dtype=np.int32)


def test_ins_block_matches_python_slices():
    for seed in range(6):
        queries, qlen, pieces, po, pc = _case(seed)
        seq4, lo, hi = _pack(queries)
        want = [packing.ins_sequence(pieces, int(po[i]), int(pc[i]), lambda rec: queries[rec]) for i in range(len(po))]
        bases, off, rest = packing.ins_block_from_packed(pieces, po, pc, seq4, lo, hi, qlen)
        store = packing.InsStore()
        store.add_strings(["x", "yy"])          # an earlier block
        first = len(store)
        store.add_block(bases, off)
        assert len(store) == first + len(po)
        slow = set(rest.tolist())
        assert slow, "the case must exercise the fallback (reverse complement / negative indices)"
        assert len(slow) < len(po)
        for i in slow:
            store[first + i] = want[i]
        for i in range(len(po)):
            assert store[first + i] == want[i], (seed, i, pieces[po[i]:po[i] + pc[i]].tolist())
        assert list(store) == ["x", "yy"] + want
        assert store[0] == "x" and store[1] == "yy"


def test_ins_block_pieces_in_any_order():
    """The device allocates piece slots with atomics: signatures' piece ranges come in no particular order."""
    for seed in (11, 12):
        queries, qlen, pieces, po, pc = _case(seed, n_sig=120)
        seq4, lo, hi = _pack(queries)
        want = [packing.ins_sequence(pieces, int(po[i]), int(pc[i]), lambda rec: queries[rec]) for i in range(len(po))]
        rng = np.random.default_rng(seed)
        perm = rng.permutation(len(po))          # shuffle the signatures' ranges inside the piece array (+ an unused slot)
        new_pieces, new_po = [(0, 0, 0, 0)], np.zeros(len(po), dtype=np.int32)
        for i in perm:
            new_po[i] = len(new_pieces)
            new_pieces.extend(pieces[po[i]:po[i] + pc[i]].tolist())
        new_pieces = np.array(new_pieces, dtype=np.int32).reshape(-1, 4)
        bases, off, rest = packing.ins_block_from_packed(new_pieces, new_po, pc, seq4, lo, hi, qlen)
        store = packing.InsStore()
        store.add_block(bases, off)
        slow = set(rest.tolist())
        assert len(slow) < len(po)
        for i in range(len(po)):
            if i not in slow:
                assert store[i] == want[i], (seed, i)
