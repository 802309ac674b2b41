"""CPU, world_size 2 over gloo: contig sharding (LPT), per-rank pipeline, ONE padded all-gather of
the candidate records, merge back into the single-device order.  The per-rank compute is the
test-only emulator (no GPU here); on the GPU box bench.py --gpus N runs the same plumbing on NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, seed_cfg, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import torch.distributed as dist
    import emul_lib
    from cutesv_b200 import _abi, shard, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.make_config(*seed_cfg)
    p = _abi.default_params(**cfg["params"])
    owner = shard.lpt_assign(shard.contig_weights(cfg["sigs"], len(cfg["lens"])), world)
    sigs, reads, index = shard.shard_inputs(cfg["sigs"], cfg["reads"], owner, rank)
    mine = emul_lib.cluster(p, cfg["lens"], sigs, reads)
    mine = (shard.globalize_aux(mine[0], index), mine[1], mine[2])
    parts = shard.all_gather_records(dist, *mine)
    merged = shard.merge_results(parts)
    if rank == 0:
        np.savez(os.path.join(out_dir, "merged.npz"), cands=merged[0], genos=merged[1], names=merged[2])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg", [(3, 0.01), (2, 0.004)])
def test_two_rank_shard_and_gather(tmp_path, cfg):
    from cutesv_b200 import _abi, synth
    from oracle import compare_records, oracle_lib
    port = _free_port()
    mp.spawn(_worker, args=(2, port, cfg, str(tmp_path)), nprocs=2, join=True)
    z = np.load(os.path.join(str(tmp_path), "merged.npz"))
    full = synth.make_config(*cfg)
    p = _abi.default_params(**full["params"])
    ref = oracle_lib.cluster(p, full["lens"], full["sigs"], full["reads"])
    d = compare_records.diff_records(ref, (z["cands"], z["genos"], z["names"]))
    assert not d, "\n".join(d[:3])
    assert len(ref[0]) > 0


def test_lpt_balances():
    from cutesv_b200 import shard
    w = [100, 90, 80, 10, 10, 10, 5, 5]
    owner = shard.lpt_assign(w, 2)
    loads = [sum(x for x, o in zip(w, owner) if o == r) for r in range(2)]
    assert max(loads) <= (sum(w) / 2.0) * 4.0 / 3.0  # LPT bound
    assert (shard.lpt_assign(w, 2) == owner).all()
