"""The one collective of the path, issued from inside libsfb (NCCL resolved at run time): an all-gather of batch results.

`Communicator` wraps sfb_comm_* / sfb_allgather_results (include/sfb.h).  Each rank (one process per GPU) solves its own
contiguous shard; the result arrays are gathered as ONE grouped NCCL operation on the communicator's own stream, ordered
after the solve, so the next solve overlaps the exchange.  `torch.distributed` is used only to hand the 128-byte NCCL id from
rank 0 to the other ranks (plumbing); the data path never touches it.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import Handle


class Communicator:
    def __init__(self, handle: Handle, world: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        self._h, self.world, self.rank = handle, int(world), int(rank)
        self._c = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        handle.check(_lib.lib().sfb_comm_create(handle.raw, self.world, self.rank, buf, C.byref(self._c)))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = _lib.lib().sfb_comm_unique_id(buf)
        if rc != 0:
            raise _lib.SfbError(rc, _lib.lib().sfb_last_error_message(None).decode())
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, handle: Handle, group=None) -> "Communicator":
        import torch.distributed as dist

        world, rank = dist.get_world_size(group), dist.get_rank(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(handle, world, rank, box[0])

    def all_gather(self, send: list, recv: list | None = None) -> list:
        """send: contiguous CUDA tensors of this rank's shard (equal shapes on every rank).  Returns the gathered tensors
        [world * B_local, ...]; asynchronous: call wait() before consuming them."""
        import torch

        if recv is None:
            recv = [torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in send]
        k = len(send)
        for s, r in zip(send, recv):
            assert s.is_cuda and s.is_contiguous() and r.is_contiguous() and r.numel() == self.world * s.numel()
        sp = (C.c_void_p * k)(*[t.data_ptr() for t in send])
        rp = (C.c_void_p * k)(*[t.data_ptr() for t in recv])
        nb = (C.c_size_t * k)(*[t.numel() * t.element_size() for t in send])
        self._h.check(_lib.lib().sfb_allgather_results(self._c, k, sp, rp, nb))
        return recv

    def wait(self, host: bool = False) -> None:
        self._h.check(_lib.lib().sfb_comm_wait(self._c, int(host)))

    def close(self) -> None:
        if self._c:
            _lib.lib().sfb_comm_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
