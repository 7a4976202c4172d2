"""Tensor-parallel plumbing: one process per GPU, ``torch.distributed`` for rendezvous, a raw NCCL
communicator (torch-bundled libnccl.so.2, via ctypes) handed to the C++ engine for the residual-path
all-reduce.  Mirrors ``Comm::from_rank`` (/root/reference/src/openai/pipelines/pipeline.rs:805-812)
and ``AllReduce`` (/root/reference/src/openai/distributed.rs:547-654); the NCCL unique id travels
over the host control plane (reference: env/TCP, communicator.rs:218-324; here: torch.distributed).
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import torch
import torch.distributed as dist

from ._lib import BackendError, check, lib


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


_nccl = None


def _libnccl() -> C.CDLL:
    global _nccl
    if _nccl is None:
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2"))
        cands += ["libnccl.so.2"]
        err = None
        for c in cands:
            try:
                _nccl = C.CDLL(c, mode=C.RTLD_GLOBAL)
                break
            except OSError as e:       # pragma: no cover
                err = e
        if _nccl is None:
            raise BackendError(f"libnccl.so.2 not found: {err}")
        _nccl.ncclGetErrorString.restype = C.c_char_p
    return _nccl


def _ok(rc: int, what: str) -> None:
    if rc != 0:
        raise BackendError(f"{what}: {_libnccl().ncclGetErrorString(rc).decode()}")


class Comm:
    """Raw NCCL communicator for (rank, world) on the current CUDA device."""

    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world = rank, world
        self.handle = C.c_void_p(0)
        if world == 1:
            return                                       # dummy Comm (distributed.rs:12-33)
        n = _libnccl()
        uid = _UniqueId()
        if rank == 0:
            _ok(n.ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
        buf = torch.tensor(list(bytes(uid)), dtype=torch.uint8)
        if dist.get_backend(group) == "nccl":
            buf = buf.cuda()
        dist.broadcast(buf, src=0, group=group)
        C.memmove(C.byref(uid), bytes(buf.cpu().numpy().tolist()), 128)
        _ok(n.ncclCommInitRank(C.byref(self.handle), C.c_int(world), uid, C.c_int(rank)), "ncclCommInitRank")

    def all_reduce_f32_(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum all-reduce on the current stream through the backend's entry point."""
        if self.world == 1:
            return t
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise BackendError("all_reduce_f32_: contiguous f32 tensor expected")
        lib().b200_allreduce_f32(self.handle, C.c_void_p(t.data_ptr()), C.c_int64(t.numel()),
                                 C.c_int64(torch.cuda.current_stream().cuda_stream))
        check("all_reduce")
        return t

    def destroy(self) -> None:
        if self.handle:
            _libnccl().ncclCommDestroy(self.handle)
            self.handle = C.c_void_p(0)
