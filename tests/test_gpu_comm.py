"""The library's own collective (sfb_comm_*, sfb_allgather_results: NCCL resolved at run time).  On the single GPU of the
test tier the communicator has world size 1 (the all-gather degenerates to a copy, but unique id, ncclCommInitRank, the
grouped launch on the communicator stream and the stream ordering are all exercised); tests/multi_gpu_check.py is the same
check under torchrun on 2+ GPUs (run with gpurun --gpus N, results under profiles/)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_allgather_world1_orders_after_the_solve():
    import torch

    import smooth_feedback_b200 as sfb
    from smooth_feedback_b200.generators import random_qp_torch

    h = sfb.Handle(0)
    h.set_stream(torch.cuda.current_stream().cuda_stream)
    comm = sfb.Communicator(h, 1, 0, sfb.Communicator.unique_id())
    B, n, m = 2048, 10, 20
    P, q, A, l, u = random_qp_torch(B, n, m, seed=3)
    r = sfb.solve_dense_batch(P, q, A, l, u, sfb.QPSolverParams(max_iter=4000), handle=h)
    g = comm.all_gather([r.x, r.y, r.obj, r.status, r.iter])  # enqueued behind the solve, on the communicator's stream
    comm.wait(host=True)
    for a, b in zip(g, (r.x, r.y, r.obj, r.status, r.iter)):
        assert torch.equal(a, b)
    assert (r.status == 0).all()
    comm.close()
