"""Run under torchrun with N >= 2 ranks on one B200 box: every `tcu_*_all` call (NCCL over
NVLink) must return, on every rank, exactly what the single-GPU call returns on that rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_check.py [--shape 700x900] [--big]

Prints one JSON line per rank-0 check; exits non-zero on the first mismatch."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pytrimal_b200 as pb
from pytrimal_b200.synthetic import synthetic_msa


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="700x900,1500x333,257x4100")
    ap.add_argument("--big", action="store_true", help="also C3-sized identity+similarity timing")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # rendezvous + object broadcast only; data moves over NCCL inside the library
    X = ord("X")
    smx = pb.SimilarityMatrix.aa()
    ok = True
    with pb.Communicator.from_torch(local) as comm:
        for shape in args.shape.split(","):
            n, L = (int(v) for v in shape.split("x"))
            m = synthetic_msa(n, L, 11)
            rng = np.random.default_rng(5)
            save_seq = np.arange(n, dtype=np.int32)
            save_seq[rng.random(n) < 0.1] = -1
            with pb.DeviceAlignment(m, device=local) as d:
                g1, h1, mx1 = d.gaps()
                gm1, _, _ = d.gaps(save_seq=save_seq)
                sp1 = d.spurious(0.5, indet=X)
                id1 = d.identity(X, keep_on_device=True)
                mdk1, num1, den1 = d.similarity(smx, gaps=g1, indet=X)
                idm1 = d.identity(X, save_seq=save_seq)
                rep1 = d.representatives(0.6, indet=X)
            with pb.DeviceAlignment(m, device=local) as d:
                g2, h2, mx2 = d.gaps(comm=comm)
                gm2, _, _ = d.gaps(save_seq=save_seq, comm=comm)
                sp2 = d.spurious(0.5, indet=X, comm=comm)
                id2 = d.identity(X, comm=comm)
                t_id = d.timings
                mdk2, num2, den2 = d.similarity(smx, gaps=g2, indet=X, comm=comm)
                t_sim = d.timings
                idm2 = d.identity(X, save_seq=save_seq, comm=comm)
                rep2 = d.representatives(0.6, indet=X, comm=comm)
            # sharded upload (each rank 1/N of the rows + all-gather) must give the same device
            # matrix as uploading everything: compare byte histogram, lengths and representatives
            import ctypes as C
            from pytrimal_b200 import _lib
            lib = pb.load()
            hm = torch.from_numpy(m).pin_memory()
            h = C.c_void_p()
            _lib.check(lib.tcu_msa_create_all(comm._h, C.c_void_p(hm.data_ptr()), n, L, L, C.byref(h)))
            hist = (C.c_ulonglong * 256)()
            _lib.check(lib.tcu_byte_histogram(h, hist))
            rep3 = np.zeros(n, np.int32)
            k3 = C.c_int(0)
            _lib.check(lib.tcu_representatives_all(h, comm._h, None, X, C.c_float(0.6),
                                                   rep3.ctypes.data_as(C.POINTER(C.c_int)), C.byref(k3)))
            lib.tcu_msa_destroy(h)
            checks = {
                "create_all": (np.array(hist[:], np.uint64) ==
                               np.bincount(m.reshape(-1), minlength=256).astype(np.uint64)).all()
                and rep3[:k3.value].tolist() == rep1.tolist(),
                "representatives": rep1.tolist() == rep2.tolist(),
                "gaps": (g1 == g2).all() and (h1 == h2).all() and mx1 == mx2,
                "gaps_masked": (gm1 == gm2).all(),
                "spurious": (bits(sp1) == bits(sp2)).all(),
                "identity": (bits(id1) == bits(id2)).all(),
                "identity_masked": (bits(idm1) == bits(idm2)).all(),
                "similarity": (bits(num1) == bits(num2)).all() and (bits(den1) == bits(den2)).all()
                and (bits(mdk1) == bits(mdk2)).all(),
            }
            flags = torch.tensor([int(bool(v)) for v in checks.values()])
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
            good = bool(flags.min().item())
            ok = ok and good
            if rank == 0:
                print(json.dumps({"shape": [n, L], "world": world, "all_ranks_bit_identical": good,
                                  "checks": {k: bool(v) for k, v in checks.items()},
                                  "identity_all_ms": t_id, "similarity_all_ms": t_sim}), flush=True)
        if args.big:
            n, L = 10000, 5000
            m = synthetic_msa(n, L, 3)
            with pb.DeviceAlignment(m, device=local) as d:
                g, _, _ = d.gaps(comm=comm)
                for it in range(2):
                    d.identity(X, comm=comm)
                    t_id = d.timings
                    mdk, num, den = d.similarity(smx, gaps=g, indet=X, comm=comm)
                    t_sim = d.timings
                csum = int(bits(num).astype(np.uint64).sum() + bits(den).astype(np.uint64).sum())
            sums = [None] * world
            dist.all_gather_object(sums, csum)
            same = len(set(sums)) == 1
            ok = ok and same
            if rank == 0:
                print(json.dumps({"workload": "C3 10000x5000 identity_all + similarity_all",
                                  "world": world, "ranks_agree": same, "checksum": csum,
                                  "identity_all_ms": t_id, "similarity_all_ms": t_sim}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
