"""Multi-GPU: queries are independent given the reference GLVs (core.ml:113 is a pure function of
(name, seq)), which the reference exploits with fork-per-job (pplacer_run.ml:618-649,
multiprocessing.ml:279).  Here: one process per GPU, the reference GLVs replicated (each rank runs
the device precompute itself - seconds - instead of receiving 11-54 GB over NVLink), unique queries
split into contiguous shards, results gathered on rank 0.  There is NO collective on the data path;
torch.distributed (NCCL on GPUs, gloo in the CPU tests) carries only the final gather.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

from . import _lib


def shard_range(Q: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of Q unique queries: the first Q % world ranks get one extra."""
    base, extra = divmod(Q, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_results(out: np.ndarray, n_out: np.ndarray, status: np.ndarray, Q: int, rank: int, world: int,
                   device: Optional[str] = None):
    """Gather per-rank (results[Qr,K], n_out[Qr], status[Qr]) to rank 0 in query order.  Returns the
    full arrays on rank 0 and None elsewhere.  Byte tensors carry the pp_result records."""
    import torch
    import torch.distributed as dist

    K = out.shape[1]
    if world == 1:
        return out, n_out, status
    sizes = [shard_range(Q, r, world)[1] - shard_range(Q, r, world)[0] for r in range(world)]
    cap = max(sizes)
    rec = _lib.RESULT_DTYPE.itemsize
    buf = np.zeros(cap * (K * rec + 8), dtype=np.uint8)
    n = out.shape[0]
    buf[:n * K * rec] = np.frombuffer(out.tobytes(), dtype=np.uint8)
    tail = buf[cap * K * rec:]
    tail[:n * 4] = np.frombuffer(n_out.astype(np.int32).tobytes(), dtype=np.uint8)
    tail[cap * 4:cap * 4 + n * 4] = np.frombuffer(status.astype(np.int32).tobytes(), dtype=np.uint8)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    gathered = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
    dist.gather(t, gathered, dst=0)
    if rank != 0:
        return None
    full = np.zeros((Q, K), dtype=_lib.RESULT_DTYPE)
    full_n = np.zeros(Q, dtype=np.int32)
    full_s = np.zeros(Q, dtype=np.int32)
    for r in range(world):
        lo, hi = shard_range(Q, r, world)
        b = gathered[r].cpu().numpy()
        m = hi - lo
        full[lo:hi] = np.frombuffer(b[:m * K * rec].tobytes(), dtype=_lib.RESULT_DTYPE).reshape(m, K)
        tl = b[cap * K * rec:]
        full_n[lo:hi] = np.frombuffer(tl[:m * 4].tobytes(), dtype=np.int32)
        full_s[lo:hi] = np.frombuffer(tl[cap * 4:cap * 4 + m * 4].tobytes(), dtype=np.int32)
    return full, full_n, full_s


def place_sharded(query_chars: np.ndarray, place_fn: Callable, keep_at_most: int, rank: int, world: int,
                  device: Optional[str] = None):
    """Run `place_fn(chars_shard) -> (out, n_out, status)` on this rank's shard and gather on rank 0.
    `place_fn` is Placer.place bound to this rank's GPU (or any stand-in with the same contract)."""
    Q = query_chars.shape[0]
    lo, hi = shard_range(Q, rank, world)
    if hi > lo:
        out, n_out, status = place_fn(query_chars[lo:hi])
    else:
        out = np.zeros((0, keep_at_most), dtype=_lib.RESULT_DTYPE)
        n_out = np.zeros(0, dtype=np.int32)
        status = np.zeros(0, dtype=np.int32)
    return gather_results(out, n_out, status, Q, rank, world, device)
