"""N > 1 on real GPUs: the `tcu_*_all` calls (row bands / column groups per rank, NCCL
exchange inside the library) must be bit-identical to the single-GPU calls on every rank.
Needs two B200s in the box; the driver's 1-GPU run skips it (tools/multigpu_check.py is the
body, also run by hand with `gpurun --gpus 2`)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("peer_memory", [True, False])
def test_all_rank_calls_match_single_gpu(gpu, peer_memory):
    """With peer memory (CUDA IPC: the exchanges are stores from the producing kernels into the
    other ranks' buffers) and without it (TRIMAL_CUDA_NO_PEER=1: NCCL point-to-point transfers,
    the path a box takes whose ranks cannot map each other's memory)."""
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs in the box")
    world = min(gpu.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "multigpu_check.py")]
    env = dict(os.environ)
    env.pop("TRIMAL_CUDA_NO_PEER", None)
    if not peer_memory:
        env["TRIMAL_CUDA_NO_PEER"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"all_ranks_bit_identical": true' in r.stdout
