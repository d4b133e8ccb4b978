"""Multi-GPU execution of the hot path: independent instances sharded over ranks, one all-gather of results.

The reference is a single-instance, single-thread library (no distributed code at all), so nothing is
ported here.  Instances of a batch never interact, hence:

  * rank r of G owns the contiguous block [r * ceil(B/G), min(B, (r+1) * ceil(B/G)))  -- inputs are placed per rank
    and never exchanged;
  * the only collective of the path is ONE all-gather of the packed per-instance results
        QP : [x (n) | y (m) | objective | status | iter]      -> (n + m + 3) scalars per instance
        EKF: [delta (d) | P_new (d*d)]                          -> (d + d*d) scalars per instance
    over NCCL (NVLink 5 / NVSwitch on the 8xB200 box); the CPU tests run the same code over gloo.

One process per GPU, launched with torchrun; `torch.distributed` is plumbing only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional


def shard_range(batch: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block of instances owned by `rank` (possibly empty for trailing ranks)."""
    per = -(-batch // world)
    lo = min(batch, rank * per)
    return lo, min(batch, lo + per)


def shard_size(batch: int, world: int) -> int:
    return -(-batch // world)


def pack_qp_results(res, n: int, m: int):
    """QPBatchResult of torch tensors -> one [B, n+m+3] tensor (status / iter are exactly representable)."""
    import torch

    B = res.x.shape[0]
    out = torch.empty((B, n + m + 3), dtype=res.x.dtype, device=res.x.device)
    out[:, :n] = res.x
    out[:, n:n + m] = res.y
    out[:, n + m] = res.obj
    out[:, n + m + 1] = res.status
    out[:, n + m + 2] = res.iter
    return out


@dataclass
class GatheredQP:
    x: object
    y: object
    obj: object
    status: object
    iter: object


def unpack_qp_results(packed, n: int, m: int) -> GatheredQP:
    import torch

    return GatheredQP(x=packed[:, :n], y=packed[:, n:n + m], obj=packed[:, n + m],
                      status=packed[:, n + m + 1].to(torch.int32), iter=packed[:, n + m + 2].to(torch.int64))


def all_gather_rows(local, batch: int, group=None):
    """All-gather row blocks that were cut with `shard_range`; returns the [batch, ...] tensor on every rank.

    Every rank contributes exactly ceil(batch/world) rows (the tail is padded) so that a single
    all_gather_into_tensor suffices.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    per = shard_size(batch, world)
    if local.shape[0] != per:
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out[:batch]


def solve_dense_batch_sharded(P_cm, q, A_cm, l, u, prm=None, group=None, solver: Optional[Callable] = None,
                              warm_x=None, warm_y=None) -> GatheredQP:
    """Solve a batch that is REPLICATED on every rank by sharding it, then all-gather the results.

    (When each rank already holds only its own shard, call `solve_dense_batch` on it and `all_gather_rows` on
    `pack_qp_results(...)` -- that is what bench.py does.)  `solver` defaults to the CUDA engine; the CPU tests
    inject a stand-in with the same signature.
    """
    import torch.distributed as dist

    if solver is None:
        from .qp import solve_dense_batch as solver
    B, n, _ = P_cm.shape
    m = A_cm.shape[2]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(B, world, rank)
    sl = slice(lo, hi)
    kw = {}
    if warm_x is not None:
        kw = dict(warm_x=warm_x[sl].contiguous(), warm_y=warm_y[sl].contiguous())
    res = solver(P_cm[sl].contiguous(), q[sl].contiguous(), A_cm[sl].contiguous(), l[sl].contiguous(),
                 u[sl].contiguous(), prm, **kw)
    packed = pack_qp_results(res, n, m)
    return unpack_qp_results(all_gather_rows(packed, B, group), n, m)


def solve_sparse_batch_sharded(pattern, P_vals, q, A_vals, l, u, prm=None, group=None, solver: Optional[Callable] = None,
                               warm_x=None, warm_y=None) -> GatheredQP:
    """Sparse shared-pattern counterpart of `solve_dense_batch_sharded` (the MPC fleet: one pattern, agents sharded over
    ranks).  `pattern` is this rank's `SparsePattern` (each rank analyses the pattern on its own device); `solver`
    defaults to `solve_sparse_batch`, the CPU tests inject a stand-in with the same signature."""
    import torch.distributed as dist

    if solver is None:
        from .qp_sparse import solve_sparse_batch as solver
    B, n = q.shape
    m = l.shape[1]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(B, world, rank)
    sl = slice(lo, hi)
    kw = {}
    if warm_x is not None:
        kw = dict(warm_x=warm_x[sl].contiguous(), warm_y=warm_y[sl].contiguous())
    res = solver(pattern, P_vals[sl].contiguous(), q[sl].contiguous(), A_vals[sl].contiguous(), l[sl].contiguous(),
                 u[sl].contiguous(), prm, **kw)
    packed = pack_qp_results(res, n, m)
    return unpack_qp_results(all_gather_rows(packed, B, group), n, m)


def ekf_step_sharded(P_cm, A_cm, Q_cm, H_cm, R_cm, innov, tau: float, group=None, predict: Optional[Callable] = None,
                     update: Optional[Callable] = None):
    """One EKF predict + update for a replicated batch of filters: shard, run, all-gather {delta, P_new}.

    Default: the fused single-pass kernel (`ekf_step_batch`); `predict` / `update` stand-ins (CPU tests) run the two
    halves separately."""
    import torch
    import torch.distributed as dist

    B, d, _ = P_cm.shape
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    lo, hi = shard_range(B, world, rank)
    sl = slice(lo, hi)
    c = lambda t: t[sl].contiguous()
    if predict is None and update is None:
        from .ekf import ekf_step_batch

        delta, Pu = ekf_step_batch(c(P_cm), c(A_cm), c(Q_cm), tau, c(H_cm), c(R_cm), c(innov))
    else:
        if predict is None:
            from .ekf import ekf_predict_batch as predict
        if update is None:
            from .ekf import ekf_update_batch as update
        Pp = predict(c(P_cm), c(A_cm), c(Q_cm), tau)
        delta, Pu = update(Pp, c(H_cm), c(R_cm), c(innov))
    packed = torch.cat([delta, Pu.reshape(hi - lo, d * d)], dim=1)
    out = all_gather_rows(packed, B, group)
    return out[:, :d], out[:, d:].reshape(B, d, d)
