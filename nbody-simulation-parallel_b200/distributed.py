"""One-process-per-GPU plumbing (torchrun): shard geometry, NCCL bootstrap, result assembly.

torch.distributed is used ONLY as plumbing: to broadcast the 128-byte NCCL unique id that
libnb200's own communicator is built from, for barriers, and to assemble per-rank rows of a
result on the host.  The data path (the per-step position all-gather) is issued by libnb200
itself on its own NCCL communicator and CUDA streams (csrc/nb200_api.cu).

Targets are sharded by contiguous index range, in whole tiles of 256 bodies:
rank g owns tiles [g*T, (g+1)*T), T = ceil(ceil(n/256)/world)  (mirrors place_shard()).
"""
from __future__ import annotations

import os

import numpy as np

TILE = 256


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) body indices owned by ``rank`` -- must equal nb200_shard_range()."""
    tiles = max(1, -(-n // TILE))
    per = -(-tiles // world)
    lo = min(n, rank * per * TILE)
    hi = min(n, (rank + 1) * per * TILE)
    return lo, max(lo, hi)


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def broadcast_bytes(blob: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a fixed-size byte string from ``src`` over the default process group
    (works on gloo and nccl: goes through a uint8 tensor on the group's device)."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).clone().to(dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def all_gather_bytes(blob: bytes) -> list[bytes]:
    """All-gather equal-size byte strings over the default process group (gloo or nccl)."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).clone().to(dev)
    parts = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    return [bytes(p.cpu().numpy().tobytes()) for p in parts]


def create_rank_context(pkg, dim: int, n: int, precision: int, device: int | None = None,
                        exchange: str = "auto"):
    """Build this rank's NBodyCuda context.

    The NCCL unique id comes from rank 0's libnb200 and is broadcast here; then every rank exports
    its CUDA IPC blob, the blobs are all-gathered, and the fused NVLink peer-store exchange is
    attached (``exchange`` = "auto" | "p2p" | "nccl").  If any rank cannot attach (no peer access),
    all ranks stay on the NCCL all-gather."""
    import ctypes

    import torch.distributed as dist

    rank, world = dist.get_rank(), dist.get_world_size()
    uid = None
    if world > 1:
        if rank == 0:
            buf = ctypes.create_string_buffer(pkg._lib.UNIQUE_ID_BYTES)
            rc = pkg._lib.load().nb200_get_unique_id(buf)
            if rc != 0:
                raise pkg.NB200Error("nb200_get_unique_id failed: " +
                                     (pkg._lib.load().nb200_last_error(None) or b"").decode())
            uid = buf.raw
        uid = broadcast_bytes(uid, pkg._lib.UNIQUE_ID_BYTES, src=0)
    if device is None:
        device = env_rank_world()[2]
    ctx = pkg.NBodyCuda(dim, n, precision, device=device, rank=rank, world=world, unique_id=uid)
    if world > 1 and exchange in ("auto", "p2p"):
        import torch

        blobs = all_gather_bytes(ctx.ipc_export())
        ok = 1
        try:
            ctx.ipc_attach(blobs)
        except pkg.NB200Error:
            if exchange == "p2p":
                raise
            ok = 0
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([ok], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t.item()) == 0:
            ctx.set_option("exchange", 0)
    return ctx


def assemble_rows(rows_full: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """Every rank passes a full-size array in which only rows [lo,hi) are meaningful; returns the
    complete array on every rank (sum of the zero-masked contributions)."""
    import torch
    import torch.distributed as dist

    masked = np.zeros_like(rows_full)
    masked[lo:hi] = rows_full[lo:hi]
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.from_numpy(masked).to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
