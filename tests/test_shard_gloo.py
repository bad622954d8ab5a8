"""The N>1 path on CPU: two gloo ranks shard a query set, each 'places' its shard with a stand-in
place function (the CUDA library cannot run here), rank 0 gathers in query order."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from pplacer_b200 import _lib, shard


def _fake_place(chars):
    """Deterministic stand-in with pp_place's output contract: encodes the query bytes."""
    Q = chars.shape[0]
    out = np.zeros((Q, 3), dtype=_lib.RESULT_DTYPE)
    n_out = np.zeros(Q, dtype=np.int32)
    status = np.zeros(Q, dtype=np.int32)
    for q in range(Q):
        h = int(chars[q].astype(np.int64).sum())
        n_out[q] = 1 + h % 3
        status[q] = 0 if h % 7 else 103
        for k in range(n_out[q]):
            out[q, k]["edge"] = h + k
            out[q, k]["log_like"] = -float(h) / (k + 1)
            out[q, k]["ml_ratio"] = 1.0 / n_out[q]
    return out, n_out, status


def _worker(rank, world, port, chars, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = shard.place_sharded(chars, _fake_place, 3, rank, world)
        # the every-step form bench.py uses: cached staging and output arrays, called twice (the second call overwrites)
        lo, hi = shard.shard_range(chars.shape[0], rank, world)
        for _ in range(2):
            res2 = shard.gather_results(*_fake_place(chars[lo:hi]), chars.shape[0], rank, world, reuse=True)
        if rank == 0:
            assert all(np.array_equal(a, b) for a, b in zip(res, res2))
        if rank == 0:
            ret["out"] = res[0].tobytes()
            ret["n_out"] = res[1].tolist()
            ret["status"] = res[2].tolist()
        else:
            assert res is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_place_matches_single():
    rng = np.random.default_rng(0)
    chars = rng.integers(65, 85, size=(11, 40), dtype=np.uint8)   # odd Q: uneven shards
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, chars, ret), nprocs=2, join=True)
    out, n_out, status = _fake_place(chars)
    assert ret["n_out"] == n_out.tolist() and ret["status"] == status.tolist()
    assert ret["out"] == out.tobytes()
