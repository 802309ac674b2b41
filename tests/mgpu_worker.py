"""Worker of tests/test_gpu_multi.py (launched with torch.distributed.run, one rank per GPU):
ONE genome contig-sharded over the ranks (LPT), every rank runs the CUDA pipeline on its shard (csv_set_shard),
csv_allgather (ONE ncclAllGather + device merge) and rank 0 compares the merged records with the oracle on the
whole genome.  Prints "MGPU OK" on success."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--scale", type=float, default=0.05)
    a = ap.parse_args()
    import torch.distributed as dist
    from cutesv_b200 import _abi, shard, synth
    from cutesv_b200.engine import Engine
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    dist.init_process_group("gloo")   # only ships the NCCL unique id and the shard index tables
    cfg = synth.make_config(a.config, a.scale)
    p = _abi.default_params(**cfg["params"])
    eng = Engine(local, params=p, contig_lens=cfg["lens"])
    uid = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0], rank, world)
    res = shard.run_sharded(eng, cfg, rank, world, repeats=3)
    index = [None] * world
    dist.all_gather_object(index, res["index"].get("INS"))
    ok = True
    if rank == 0:
        from oracle import compare_records, oracle_lib
        aln = None
        if "TRA" in cfg["sigs"] and cfg["params"].get("genotype"):   # the same BAM-order table run_sharded() uploads
            r = cfg["reads"]
            order = np.lexsort((np.arange(len(r["chrom"])), r["start"], r["chrom"]))
            aln = {k: v[order] for k, v in r.items()}
        ref = oracle_lib.cluster(p, cfg["lens"], cfg["sigs"], cfg["reads"], n_threads=8, aln=aln)
        for k, got in enumerate(res["results"]):
            got = (shard.globalize_aux_by_rank(got[0], index), got[1], got[2])
            d = compare_records.diff_records(ref, got)
            if d:
                ok = False
                print("MGPU DIFF (repeat %d):\n%s" % (k, "\n".join(d[:5])))
        print("rank0: %d merged candidates, %d reference rows, graph replays %d" % (len(res["results"][-1][0]), len(ref[0]), eng.graph_replays()))
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    eng.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and flag[0]:
        print("MGPU OK")
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
