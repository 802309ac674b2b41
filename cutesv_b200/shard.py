"""Contig sharding across GPUs.

Every resolution_* call of the reference is keyed by (svtype, contig) and reads only that contig's
signatures and reads-table rows (cuteSV:1116-1189), so contigs are independent units: they are
bin-packed over the ranks (LPT on the per-contig signature count), each rank runs the whole
pipeline on its contigs with no data-path collective, and ONE all-gather of the fixed-width
candidate records assembles the result in the reference's order.
"""
import numpy as np

from . import _abi


def lpt_assign(weights, n_ranks):
    """Longest-processing-time bin packing: contig -> rank.  Deterministic."""
    weights = np.asarray(weights, dtype=np.int64)
    order = sorted(range(len(weights)), key=lambda c: (-int(weights[c]), c))
    load = [0] * n_ranks
    owner = np.zeros(len(weights), dtype=np.int32)
    for c in order:
        r = min(range(n_ranks), key=lambda k: (load[k], k))
        owner[c] = r
        load[r] += int(weights[c])
    return owner


def contig_weights(sigs, n_contigs):
    w = np.zeros(n_contigs, dtype=np.int64)
    for cols in sigs.values():
        if cols is not None and len(cols["chrom"]):
            w += np.bincount(cols["chrom"], minlength=n_contigs)
    return w


def shard_inputs(sigs, reads, owner, rank):
    """The rank's share of the columnar inputs (contig ids stay global).  Returns (sigs, reads,
    index) where index[type] maps a local signature index back to its index in the full arrays
    (INS candidates name the signature that carries their ALT sequence, csv_cand.aux)."""
    out, index = {}, {}
    for t, cols in sigs.items():
        if cols is None:
            out[t] = None
            continue
        m = owner[cols["chrom"]] == rank
        out[t] = {k: (v[m] if v is not None else None) for k, v in cols.items()}
        index[t] = np.flatnonzero(m).astype(np.int32)
    rd = None
    if reads is not None:
        m = owner[reads["chrom"]] == rank
        rd = {k: v[m] for k, v in reads.items()}
    return out, rd, index


def globalize_aux(cands, index):
    """Rewrite csv_cand.aux of INS rows from shard-local to global signature indices."""
    m = cands["svtype"] == _abi.CSV_INS
    if m.any() and "INS" in index:
        cands = cands.copy()
        cands["aux"][m] = index["INS"][cands["aux"][m]]
    return cands


def globalize_aux_by_rank(cands, index_by_rank):
    """Gathered records (csv_fetch_gathered): csv_cand.reserved[1] is the source rank; rewrite the aux of INS rows from that
    rank's shard-local signature index to the index in the full arrays.  index_by_rank[r] = shard index table of rank r."""
    m = cands["svtype"] == _abi.CSV_INS
    if not m.any():
        return cands
    cands = cands.copy()
    src = cands["reserved"][:, 1]
    for r, idx in enumerate(index_by_rank):
        mm = m & (src == r)
        if mm.any():
            cands["aux"][mm] = idx[cands["aux"][mm]]
    return cands


def owned_mask(owner, rank):
    return (np.asarray(owner) == rank).astype(np.uint8)


def run_sharded(eng, cfg, rank, world, repeats=1, type_mask=None):
    """One genome on `world` GPUs through the product API: LPT contig shards, csv_set_shard, device-resident pipeline,
    csv_allgather.  eng: an Engine with csv_comm_init done.  Returns dict(results=[merged (cands, genos, names) per repeat],
    index=shard index tables, owner=contig -> rank)."""
    owner = lpt_assign(contig_weights(cfg["sigs"], len(cfg["lens"])), world)
    sigs, reads, index = shard_inputs(cfg["sigs"], cfg["reads"], owner, rank)
    eng.set_shard(owned_mask(owner, rank))
    if type_mask is None:
        type_mask = sum(1 << _abi.TYPE_IDS[k] for k in cfg["sigs"])
    if "TRA" in cfg["sigs"] and cfg["params"].get("genotype"):
        # the TRA genotyper scans BAM-order alignment records around pos1 (own contig) AND pos2 (any contig): whole table
        r = cfg["reads"]
        order = np.lexsort((np.arange(len(r["chrom"])), r["start"], r["chrom"]))
        eng.upload_alignments({k: v[order] for k, v in r.items()})
    eng.upload(sigs, reads)
    out = []
    for _ in range(repeats):
        eng.cluster_device(type_mask)
        eng.counts()
        eng.allgather()
        out.append(eng.fetch_gathered())
    return dict(results=out, index=index, owner=owner)


def merge_results(parts):
    """parts: list over ranks of (cands, genos, names).  Returns one result in the single-GPU
    order: svtype, contig id, emission order (per-contig order is preserved inside every part)."""
    cands = np.concatenate([p[0] for p in parts]) if parts else np.zeros(0, _abi.CAND_DTYPE)
    genos = np.concatenate([p[1] for p in parts]) if parts else np.zeros(0, _abi.GENO_DTYPE)
    off = 0
    shifted = []
    for c, g, n in parts:
        c = c.copy()
        c["names_off"] += off
        off += len(n)
        shifted.append(c)
    cands = np.concatenate(shifted) if shifted else cands
    names = np.concatenate([p[2] for p in parts]) if parts else np.zeros(0, np.int32)
    # stable sort by (svtype, contig): inside one contig all rows come from one rank, in order
    key = cands["svtype"].astype(np.int64) * (1 << 32) + cands["chrom"].astype(np.int64)
    order = np.argsort(key, kind="stable")
    return cands[order], genos[order], names


def all_gather_records(dist, cands, genos, names, device=None):
    """One padded all-gather of the fixed-width records (torch.distributed; NCCL on GPUs, gloo in
    the CPU tests).  Returns the list over ranks of (cands, genos, names)."""
    import torch
    world = dist.get_world_size()
    dev = device if device is not None else "cpu"
    counts = torch.tensor([len(cands), len(names)], dtype=torch.int64, device=dev)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)
    max_c = max(int(x[0]) for x in allc)
    max_n = max(int(x[1]) for x in allc)
    width = max_c * (64 + 40) + max_n * 4
    buf = np.zeros(max(width, 8), dtype=np.uint8)
    buf[: len(cands) * 64] = cands.view(np.uint8).reshape(-1)
    buf[max_c * 64: max_c * 64 + len(genos) * 40] = genos.view(np.uint8).reshape(-1)
    buf[max_c * 104: max_c * 104 + len(names) * 4] = names.view(np.uint8).reshape(-1)
    mine = torch.from_numpy(buf).to(dev)
    out = torch.empty(world * len(buf), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, mine)
    out = out.cpu().numpy().reshape(world, len(buf))
    parts = []
    for r in range(world):
        nc, nn = int(allc[r][0]), int(allc[r][1])
        c = out[r, : nc * 64].copy().view(_abi.CAND_DTYPE)
        g = out[r, max_c * 64: max_c * 64 + nc * 40].copy().view(_abi.GENO_DTYPE)
        n = out[r, max_c * 104: max_c * 104 + nn * 4].copy().view(np.int32)
        parts.append((c, g, n))
    return parts
