"""N>1 path on CPU: world_size-2 gloo processes exercise shard_range / packing / the single all-gather.

The CUDA engine is replaced by a stand-in solver with the same signature (the CPU oracle -- test infrastructure),
so what is tested here is exactly the host-side multi-GPU logic of smooth_feedback_b200/sharding.py.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_solver(P_cm, q, A_cm, l, u, prm=None, warm_x=None, warm_y=None):
    from oracle import oracle as orc
    from smooth_feedback_b200.qp import QPBatchResult

    P = np.swapaxes(P_cm.numpy(), 1, 2); A = np.swapaxes(A_cm.numpy(), 1, 2)
    o = orc.qp_solve_batch(P, q.numpy(), A, l.numpy(), u.numpy(), params=orc.default_params(max_iter=4000),
                           warm_x=None if warm_x is None else warm_x.numpy(), warm_y=None if warm_y is None else warm_y.numpy())
    t = torch.from_numpy
    return QPBatchResult(x=t(o.x), y=t(o.y), obj=t(o.obj), status=t(o.status), iter=t(o.iter.astype(np.int64)),
                         active=t(o.active), flags=torch.zeros(len(o.obj), dtype=torch.int32))


def _worker(rank, world, port, B, n, m, q_out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from smooth_feedback_b200 import sharding
        from smooth_feedback_b200.generators import random_ekf_numpy, random_qp_numpy
        from smooth_feedback_b200.qp import to_colmajor

        P, q, A, l, u = random_qp_numpy(B, n, m, seed=5)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        g = sharding.solve_dense_batch_sharded(t(to_colmajor(P)), t(q), t(to_colmajor(A)), t(l), t(u), None, solver=_oracle_solver)
        # EKF: stand-ins built on the oracle too
        from oracle import oracle as orc

        Pk, Ak, Qk, Hk, Rk, innov = random_ekf_numpy(B, 6, 3, seed=5)
        cm = lambda a: np.ascontiguousarray(np.swapaxes(a, 1, 2))
        pred = lambda P_, A_, Q_, tau: t(cm(orc.ekf_predict_batch(cm(P_.numpy()), cm(A_.numpy()), cm(Q_.numpy()), tau)))
        def upd(P_, H_, R_, inn):
            d_, Pn = orc.ekf_update_batch(cm(P_.numpy()), cm(H_.numpy()), cm(R_.numpy()), inn.numpy())
            return t(d_), t(cm(Pn))
        delta, Pu = sharding.ekf_step_sharded(t(cm(Pk)), t(cm(Ak)), t(cm(Qk)), t(cm(Hk)), t(cm(Rk)), t(innov), 0.1, predict=pred, update=upd)
        # sparse shared-pattern fleet: stand-in = densified oracle
        from smooth_feedback_b200.generators import mpc_structured_batch, mpc_structured_pattern, sparse_to_dense

        pat = mpc_structured_pattern(Nx=2, Nu=1, nivals=2, Ki=3)
        Pv, qs, Av, ls, us = mpc_structured_batch(pat, B, seed=7)

        def sparse_solver(pattern, P_vals, q_, A_vals, l_, u_, prm=None, warm_x=None, warm_y=None):
            Pd, Ad = sparse_to_dense(pattern, P_vals.numpy(), A_vals.numpy())
            return _oracle_solver(t(to_colmajor(Pd)), q_, t(to_colmajor(Ad)), l_, u_, prm, warm_x, warm_y)

        gs = sharding.solve_sparse_batch_sharded(pat, t(Pv), t(qs), t(Av), t(ls), t(us), None, solver=sparse_solver)
        if rank == 0:
            q_out.put((g.x.numpy(), g.y.numpy(), g.obj.numpy(), g.status.numpy(), g.iter.numpy(), delta.numpy(), Pu.numpy(),
                       gs.x.numpy(), gs.status.numpy(), gs.iter.numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [9, 16])  # ragged (last shard shorter) and even
def test_sharded_solve_equals_single_process(oracle, B):
    from smooth_feedback_b200.generators import random_ekf_numpy, random_qp_numpy

    n, m, world = 6, 9, 2
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + B
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, n, m, q_out)) for r in range(world)]
    for p in procs:
        p.start()
    got = q_out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    P, q, A, l, u = random_qp_numpy(B, n, m, seed=5)
    o = oracle.qp_solve_batch(P, q, A, l, u, params=oracle.default_params(max_iter=4000))
    assert np.array_equal(got[0], o.x) and np.array_equal(got[1], o.y) and np.array_equal(got[2], o.obj)
    assert np.array_equal(got[3], o.status) and np.array_equal(got[4], o.iter.astype(np.int64))
    Pk, Ak, Qk, Hk, Rk, innov = random_ekf_numpy(B, 6, 3, seed=5)
    oPp = oracle.ekf_predict_batch(Pk, Ak, Qk, 0.1)
    od, oPu = oracle.ekf_update_batch(oPp, Hk, Rk, innov)
    assert np.array_equal(got[5], od) and np.array_equal(np.swapaxes(got[6], 1, 2), oPu)
    # sparse fleet: sharded == single process
    from smooth_feedback_b200.generators import mpc_structured_batch, mpc_structured_pattern, sparse_to_dense

    pat = mpc_structured_pattern(Nx=2, Nu=1, nivals=2, Ki=3)
    Pv, qs, Av, ls, us = mpc_structured_batch(pat, B, seed=7)
    Pd, Ad = sparse_to_dense(pat, Pv, Av)
    os_ = oracle.qp_solve_batch(Pd, qs, Ad, ls, us, params=oracle.default_params(max_iter=4000))
    assert np.array_equal(got[7], os_.x) and np.array_equal(got[8], os_.status) and np.array_equal(got[9], os_.iter.astype(np.int64))


def test_shard_range_partitions_exactly():
    from smooth_feedback_b200.sharding import shard_range, shard_size

    for B in (0, 1, 7, 8, 9, 65536, 1 << 20):
        for G in (1, 2, 3, 4, 8):
            blocks = [shard_range(B, G, r) for r in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert all(0 <= hi - lo <= shard_size(B, G) for lo, hi in blocks)
