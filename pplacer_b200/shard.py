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


_bufs = {}


def _buf(key, nbytes, device=None, pin=False):
    """cached staging buffers: the gather runs every step, its buffers are allocated once"""
    import torch
    t = _bufs.get(key)
    if t is None or t.numel() < nbytes:
        if device is not None:
            t = torch.empty(nbytes, dtype=torch.uint8, device=device)
        else:
            t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=pin)
        _bufs[key] = t
    return t[:nbytes]


def gather_results(out: np.ndarray, n_out: np.ndarray, status: np.ndarray, Q: int, rank: int, world: int,
                   device: Optional[str] = None, reuse: bool = False):
    """Gather per-rank (results[Qr,K], n_out[Qr], status[Qr]) to rank 0 in query order: ONE collective
    (torch.distributed gather: NCCL on GPUs, gloo in the CPU tests) of one byte row per rank - the pp_result records,
    then the counts, then the status words.  Returns the full arrays on rank 0 and None elsewhere.
    reuse=True hands back arrays that the NEXT call overwrites (a loop that gathers every step need not pay for
    fresh 4 MB-per-rank allocations and their page faults each time)."""
    import torch
    import torch.distributed as dist

    K = out.shape[1]
    if world == 1:
        return out, n_out, status
    sizes = [shard_range(Q, r, world)[1] - shard_range(Q, r, world)[0] for r in range(world)]
    cap = max(sizes)
    rec = _lib.RESULT_DTYPE.itemsize
    row = (cap * (K * rec + 8) + 15) // 16 * 16
    n = out.shape[0]
    on_gpu = device is not None
    send = _buf(("send", on_gpu), row, pin=on_gpu)
    sv = send.numpy()
    sv[:n * K * rec] = np.ascontiguousarray(out).view(np.uint8).reshape(-1)
    sv[cap * K * rec:cap * K * rec + n * 4] = np.ascontiguousarray(n_out, dtype=np.int32).view(np.uint8)
    sv[cap * K * rec + cap * 4:cap * K * rec + cap * 4 + n * 4] = np.ascontiguousarray(status, dtype=np.int32).view(np.uint8)
    if on_gpu:
        d_send = _buf(("d_send", str(device)), row, device=device)
        d_send.copy_(send, non_blocking=True)
    else:
        d_send = send
    recv = None
    if rank == 0:
        recv = _buf(("recv", str(device)), world * row, device=device) if on_gpu else _buf(("recv", None), world * row)
    dist.gather(d_send, list(recv.view(world, row).unbind(0)) if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    if on_gpu:
        host = _buf(("host", on_gpu), world * row, pin=True)
        host.copy_(recv)                       # one device -> pinned-host copy for all ranks (synchronises)
    else:
        host = recv
    hv = host.numpy().reshape(world, row)
    if reuse:
        key = ("full", Q, K)
        if key not in _bufs:
            _bufs[key] = (np.zeros((Q, K), dtype=_lib.RESULT_DTYPE), np.zeros(Q, dtype=np.int32), np.zeros(Q, dtype=np.int32))
        full, full_n, full_s = _bufs[key]
    else:
        full = np.zeros((Q, K), dtype=_lib.RESULT_DTYPE)
        full_n = np.zeros(Q, dtype=np.int32)
        full_s = np.zeros(Q, dtype=np.int32)
    for r in range(world):
        lo, hi = shard_range(Q, r, world)
        m = hi - lo
        full[lo:hi] = hv[r, :m * K * rec].view(_lib.RESULT_DTYPE).reshape(m, K)
        full_n[lo:hi] = hv[r, cap * K * rec:cap * K * rec + m * 4].view(np.int32)
        full_s[lo:hi] = hv[r, cap * K * rec + cap * 4:cap * K * rec + cap * 4 + m * 4].view(np.int32)
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
