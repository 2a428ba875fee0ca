"""N>1 host logic on CPU: two gloo ranks each produce their band of the packed
identity array (values from the oracle standing in for the GPU) and an
all-gather of the slices must reproduce the full array."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, random_msa


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, L, tmpdir):
    import sys
    sys.path.insert(0, ROOT)
    import oracle
    from pytrimal_b200.sharding import band_partition, band_slice, row_blocks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = random_msa(np.random.default_rng(3), n, L)
    full = oracle.Port().identity(m, ord("X"))
    bounds = band_partition(n, world)
    off, cnt = band_slice(n, bounds, rank)
    mine = torch.from_numpy(full[off:off + cnt].copy())
    sizes = [band_slice(n, bounds, g)[1] for g in range(world)]
    pad = max(sizes)
    buf = torch.zeros(pad)
    buf[:cnt] = mine
    gathered = [torch.zeros(pad) for _ in range(world)]
    dist.all_gather(gathered, buf)
    rebuilt = torch.cat([gathered[g][:sizes[g]] for g in range(world)]).numpy()
    ok = rebuilt.shape == full.shape and (rebuilt.view(np.uint32) == full.view(np.uint32)).all()
    # timing-style reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == world
    open(os.path.join(tmpdir, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [130, 257, 700])
def test_two_rank_bands_rebuild_full_array(tmp_path, n):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, 90, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / f"ok{r}").read() == "1"
