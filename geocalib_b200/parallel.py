"""Multi-GPU plumbing: one process per GPU, images sharded contiguously over ranks.

Per-image intrinsics (BASELINE configs 2-4): images are independent problems -> every rank solves
its own shard, there is NO data-path collective.  `shard_range` is all that is needed.

Shared intrinsics (config 5): one focal / distortion for the union of all ranks' images.  The
engine reduces each rank's images to <= 25 Schur-complement sums and all-reduces them once per LM
step over NCCL (NVLink/NVSwitch) on the compute stream; `make_comm` creates the communicator the
C ABI needs (`gclm_comm_create`), using torch.distributed only to broadcast the 128-byte NCCL id.
"""

from __future__ import annotations

import ctypes
import os
from typing import Tuple

import numpy as np
import torch

from . import _lib


def shard_range(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `batch` images owned by `rank` (sizes differ by <= 1)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(batch, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def broadcast_bytes(payload: bytes, nbytes: int, src: int = 0, device=None) -> bytes:
    """Broadcast a small byte string from `src` with whatever backend the default group has."""
    import torch.distributed as dist

    t = torch.zeros(nbytes, dtype=torch.uint8)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    if dist.get_backend() == "nccl":
        t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


class Comm:
    """Owns a gclm_comm (an NCCL communicator private to the engine)."""

    def __init__(self, handle: ctypes.c_void_p, world: int, rank: int):
        self.handle, self.world, self.rank = handle, world, rank
        self.p2p = False

    def __del__(self):
        try:
            if self.handle:
                _lib.load().gclm_comm_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def make_comm(device: torch.device, p2p: bool = None) -> Comm:
    """Create the engine's communicator over the ranks of the default process group."""
    import torch.distributed as dist

    lib = _lib.load()
    world, rank = dist.get_world_size(), dist.get_rank()
    buf = (ctypes.c_char * _lib.COMM_ID_BYTES)()
    if rank == 0:
        _lib.check(lib.gclm_comm_unique_id(buf), "gclm_comm_unique_id")
    uid = broadcast_bytes(bytes(buf.raw), _lib.COMM_ID_BYTES, 0, device)
    h = ctypes.c_void_p()
    idx = device.index if device.index is not None else torch.cuda.current_device()
    _lib.check(lib.gclm_comm_create(uid, world, rank, idx, ctypes.byref(h)), "gclm_comm_create")
    comm = Comm(h, world, rank)
    if p2p is None:
        p2p = os.environ.get("GCLM_P2P", "1") != "0"
    if p2p and world <= 8:
        attach_peer_memory(comm, device)
    return comm


def attach_peer_memory(comm: Comm, device: torch.device) -> bool:
    """Exchange CUDA-IPC handles of the per-rank exchange buffers so that the per-step reduction
    kernel can do its all-reduce with direct NVLink stores (include/gclm.h: gclm_comm_p2p_*).

    Collective over the default group.  If ANY rank cannot export / map the buffers (no peer access,
    IPC disabled in the container, ...) every rank keeps the NCCL path -- the decision is agreed with
    a MIN all-reduce so that no rank waits on a peer that never writes."""
    import torch.distributed as dist

    lib = _lib.load()
    on_gpu = dist.get_backend() == "nccl"
    mine = (ctypes.c_char * _lib.IPC_HANDLE_BYTES)()
    ok = lib.gclm_comm_p2p_export(comm.handle, mine) == _lib.GCLM_OK
    t = torch.frombuffer(bytearray(mine.raw), dtype=torch.uint8).clone()
    t = t.to(device) if on_gpu else t
    gathered = [torch.empty_like(t) for _ in range(comm.world)]
    dist.all_gather(gathered, t)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device if on_gpu else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 1:
        blob = b"".join(bytes(g.cpu().numpy().tobytes()) for g in gathered)
        ok = lib.gclm_comm_p2p_attach(comm.handle, blob) == _lib.GCLM_OK
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device if on_gpu else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    comm.p2p = int(flag.item()) == 1
    if not comm.p2p:
        _lib.check(lib.gclm_comm_p2p_disable(comm.handle), "gclm_comm_p2p_disable")
    dist.barrier()  # every buffer is zeroed and mapped everywhere before the first exchange
    return comm.p2p


# ---- host-side mirror of the device Schur reduction (csrc: schur_image / shared_update_kernel) ----
def schur_partials(H: np.ndarray, G: np.ndarray, lam: float) -> np.ndarray:
    """Per-image partial sums [C | E^T D~^-1 E | g_c | E^T D~^-1 g_g] for one image.

    H (n,n), G (n,) with the 2 gravity columns first (reference: lm_optimizer.py:370-383)."""
    ni = H.shape[0] - 2
    D = H[:2, :2].astype(np.float64).copy()
    D[np.diag_indices(2)] += np.maximum(np.diag(D) * lam, 1e-6)
    E, C = H[:2, 2:].astype(np.float64), H[2:, 2:].astype(np.float64)
    Di = np.linalg.inv(D)
    out = np.zeros(2 * ni * ni + 2 * ni)
    out[: ni * ni] = C.ravel()
    out[ni * ni: 2 * ni * ni] = (E.T @ Di @ E).ravel()
    out[2 * ni * ni: 2 * ni * ni + ni] = G[2:]
    out[2 * ni * ni + ni:] = E.T @ Di @ G[:2]
    return out


def schur_solve(total: np.ndarray, ni: int, lam: float) -> np.ndarray:
    """Shared-intrinsics step dc from the (all-reduced) partial sums."""
    C = total[: ni * ni].reshape(ni, ni).copy()
    C[np.diag_indices(ni)] += np.maximum(np.diag(C) * lam, 1e-6)
    S = C - total[ni * ni: 2 * ni * ni].reshape(ni, ni)
    rhs = total[2 * ni * ni: 2 * ni * ni + ni] - total[2 * ni * ni + ni:]
    return np.linalg.solve(S, rhs)


def schur_backsub(H: np.ndarray, G: np.ndarray, dc: np.ndarray, lam: float) -> np.ndarray:
    """Per-image gravity step dg = D~^-1 (g_g - E dc)."""
    D = H[:2, :2].astype(np.float64).copy()
    D[np.diag_indices(2)] += np.maximum(np.diag(D) * lam, 1e-6)
    return np.linalg.solve(D, G[:2] - H[:2, 2:] @ dc)
