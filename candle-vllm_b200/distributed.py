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


class PeerInboxes:
    """NVLink peer-memory inboxes for the engine's fused all-reduce (``b200_llama_set_peer_inboxes``): allocate ours, swap the
    CUDA IPC handles over ``torch.distributed``, map the peers'.  ``B200_TP_NCCL=1`` keeps the engine on NCCL; so does any
    rank failing to allocate / map (``.active`` is False on ALL ranks then, ``.error`` says why)."""

    def __init__(self, model, rank: int, world: int, group=None):
        self.model, self.rank, self.world = model, rank, world
        self.mine, self.mapped, self.error = None, [], ""
        if world == 1 or os.environ.get("B200_TP_NCCL", "0") not in ("", "0"):
            return
        L = lib()
        on_gpu = dist.get_backend(group) == "nccl"
        nbytes = int(L.b200_llama_peer_inbox_bytes(model._h))
        handle = (C.c_ubyte * 64)()
        ok, self.error = True, ""
        # every rank walks through the same collectives whatever happens locally; success is agreed on at the end
        mine_ptr = L.b200_ipc_alloc(C.c_size_t(nbytes), handle)
        if L.b200_last_error() or not mine_ptr:
            ok, self.error, mine_ptr = False, "b200_ipc_alloc: " + L.b200_last_error_message().decode(), None
        mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8)
        mine = mine.cuda() if on_gpu else mine
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine, group=group)
        ptrs = []
        for r in range(world):
            if r == rank or not ok:
                ptrs.append(mine_ptr)
                continue
            h = (C.c_ubyte * 64)(*gathered[r].cpu().numpy().tolist())
            p = L.b200_ipc_open(h)
            if L.b200_last_error() or not p:
                ok, self.error = False, "b200_ipc_open: " + L.b200_last_error_message().decode()
                ptrs.append(None)
                continue
            self.mapped.append(p)
            ptrs.append(p)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32)
        flag = flag.cuda() if on_gpu else flag
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # also the barrier: every inbox is mapped and zeroed
        if int(flag.item()) == 0:
            for p in self.mapped:
                L.b200_ipc_close(C.c_void_p(p))
            if mine_ptr:
                L.b200_ipc_free(C.c_void_p(mine_ptr))
            self.mapped = []
            self.error = self.error or "a peer rank could not map the inboxes"
            return                                                     # inactive: the engine stays on NCCL
        self.mine = mine_ptr
        arr = (C.c_void_p * world)(*ptrs)
        L.b200_llama_set_peer_inboxes(model._h, arr, C.c_int32(world))
        check("b200_llama_set_peer_inboxes")

    @property
    def active(self) -> bool:
        return self.mine is not None

    def timed_out(self) -> bool:
        """True when some row gave up waiting for a peer (a rank died); results are NaN from then on."""
        return self.active and int(lib().b200_llama_peer_timeouts(self.model._h)) != 0

    def close(self, group=None) -> None:
        if self.mine is None:
            return
        L = lib()
        torch.cuda.synchronize()
        L.b200_llama_set_peer_inboxes(self.model._h, None, C.c_int32(0))
        if dist.is_initialized():
            dist.barrier(group=group)                   # nobody is still pushing into an inbox that is about to go away
        for p in self.mapped:
            L.b200_ipc_close(C.c_void_p(p))
        L.b200_ipc_free(C.c_void_p(self.mine))
        self.mine, self.mapped = None, []
