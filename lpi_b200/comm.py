"""lpi_comm_* through ctypes: the C-ABI form of the two exchange steps of the multi-GPU path (include/lpi_b200.h), for hosts that do not
want torch.distributed in the loop.  `LpiComm` quacks enough like a process group for `retrieval.merge_recall` / `lpi_step.train_step`
callers that pass it explicitly; the default Python path keeps using torch.distributed (the same NCCL underneath).

Reference anchor: gather_features, retrieval/methods/sprompt.py:38-82."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import call, lib, ptr, stream_ptr


def unique_id() -> bytes:
    """Created on rank 0; ship the bytes to the other ranks (file, socket, torch store, ...)."""
    n = lib().lpi_comm_unique_id_bytes()
    buf = C.create_string_buffer(n)
    call("comm_unique_id", buf)
    return buf.raw


class LpiComm:
    def __init__(self, n_ranks: int, rank: int, uid: bytes):
        """Collective: every rank calls it with the same id, after torch.cuda.set_device(local GPU)."""
        self.n_ranks, self.rank = int(n_ranks), int(rank)
        self._h = C.c_void_p()
        call("comm_init", C.byref(self._h), self.n_ranks, self.rank, C.create_string_buffer(uid, len(uid)))

    @classmethod
    def from_torch_group(cls, group=None) -> "LpiComm":
        """Bootstrap over an existing torch.distributed group (any backend): rank 0's id is broadcast as an object."""
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(world, rank, box[0])

    def all_gather(self, send: torch.Tensor, recv: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[n_ranks, *send.shape] <- every rank's `send` (contiguous CUDA tensor), rank-major."""
        send = send.contiguous()
        if recv is None:
            recv = torch.empty(self.n_ranks, *send.shape, device=send.device, dtype=send.dtype)
        call("comm_allgather", self._h, ptr(send), ptr(recv), C.c_longlong(send.numel() * send.element_size()), stream_ptr())
        return recv

    def all_reduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("all_reduce_sum_ takes a contiguous fp32 tensor")
        call("comm_allreduce_sum_f32", self._h, ptr(t), ptr(t), C.c_longlong(t.numel()), stream_ptr())
        return t

    def close(self):
        if self._h:
            call("comm_destroy", self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
