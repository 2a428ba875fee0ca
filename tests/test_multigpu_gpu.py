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


def test_all_rank_calls_match_single_gpu(gpu):
    if gpu.device_count() < 2:
        pytest.skip("needs two GPUs in the box")
    world = min(gpu.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"all_ranks_bit_identical": true' in r.stdout
