#!/usr/bin/env python
"""Multi-GPU check of the sharded path (run under torchrun, one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Every rank solves its contiguous shard of ONE seeded batch; results are exchanged with the library's grouped NCCL all-gather
(sfb_allgather_results).  Checks on every rank: the gathered arrays equal a single-GPU solve of the whole batch bit for bit
(instances are independent, so sharding must not change a single bit), for the dense QP path and the EKF cycle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smooth_feedback_b200 as sfb  # noqa: E402
from smooth_feedback_b200.generators import random_ekf_numpy, random_qp_numpy  # noqa: E402
from smooth_feedback_b200.sharding import shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    h = sfb.Handle(local)
    h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    comm = sfb.Communicator.from_torch_distributed(h)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cm = sfb.to_colmajor
    B, n, m = 4096 * world, 10, 20
    P, q, A, l, u = random_qp_numpy(B, n, m, seed=5)
    prm = sfb.QPSolverParams(max_iter=4000)
    full = sfb.solve_dense_batch(t(cm(P)), t(q), t(cm(A)), t(l), t(u), prm, handle=h)
    lo, hi = shard_range(B, world, rank)
    sl = slice(lo, hi)
    mine = sfb.solve_dense_batch(t(cm(P[sl])), t(q[sl]), t(cm(A[sl])), t(l[sl]), t(u[sl]), prm, handle=h)
    g = comm.all_gather([mine.x, mine.y, mine.obj, mine.status, mine.iter])
    comm.wait(host=True)
    for a, b, name in zip(g, (full.x, full.y, full.obj, full.status, full.iter), "x y obj status iter".split()):
        assert torch.equal(a, b), f"rank {rank}: gathered {name} differs from the single-GPU solve"
    Bk = 8192 * world
    Pk, Ak, Qk, Hk, Rk, innov = random_ekf_numpy(Bk, 6, 3, seed=5)
    dfull, Pfull = sfb.ekf_step_batch(t(cm(Pk)), t(cm(Ak)), t(cm(Qk)), 0.1, t(cm(Hk)), t(cm(Rk)), t(innov), handle=h)
    lo, hi = shard_range(Bk, world, rank)
    sl = slice(lo, hi)
    dm, Pm = sfb.ekf_step_batch(t(cm(Pk[sl])), t(cm(Ak[sl])), t(cm(Qk[sl])), 0.1, t(cm(Hk[sl])), t(cm(Rk[sl])), t(innov[sl]), handle=h)
    gd, gP = comm.all_gather([dm, Pm])
    comm.wait(host=True)
    assert torch.equal(gd, dfull) and torch.equal(gP, Pfull), f"rank {rank}: gathered EKF results differ"
    dist.barrier()
    if rank == 0:
        print(f"multi_gpu_check ok: world={world}, dense QP batch {B} and EKF batch {Bk}: gathered == single-GPU, bit for bit")
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
