"""-m gpu, needs >= 2 devices: merged N-rank result (contig shards + csv_allgather) == oracle on the whole genome, through both
gathers (peer-to-peer mail boxes over CUDA IPC = default, ncclAllGather) and through a forced re-negotiation of the padded size."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("cid,scale,env", [(2, 0.05, {}), (3, 0.05, {}), (2, 0.02, {"CUTESV_B200_GATHER_PAD": "8"}),
                                            (2, 0.05, {"CUTESV_B200_GATHER": "nccl"}), (3, 0.05, {"CUTESV_B200_GATHER": "nccl", "CUTESV_B200_GATHER_PAD": "8"})])
def test_sharded_ranks_match_oracle(cid, scale, env):
    n = _n_devices()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "mgpu_worker.py"), "--config", str(cid), "--scale", str(scale)]
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=e, cwd=ROOT)
    assert r.returncode == 0 and "MGPU OK" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
