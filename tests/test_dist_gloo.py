"""world_size-2 CPU (gloo) test of the N>1 host path: contiguous world shards + the cost allgather
+ identical selection on every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from eggshell_b200 import mpc


def test_shard_range_partitions_exactly():
    for total in (1, 7, 64, 65536, 1048576 + 3):
        for ws in (1, 2, 3, 8):
            spans = [mpc.shard_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
            for g in (0, total // 2, total - 1):
                r, l = mpc.owner_of(g, total, ws)
                assert spans[r][0] + l == g


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    lo, hi = mpc.shard_range(total, rank, ws)
    # deterministic per-world "rollout cost" that depends only on the global world index
    g = np.arange(lo, hi, dtype=np.float64)
    local = torch.from_numpy(np.cos(g * 0.37) + 1e-3 * g)
    allc = mpc.allgather_costs(local)
    idx, val = mpc.select_best(allc, k=3)
    q.put((rank, allc.numpy().copy(), idx.copy(), val.copy()))
    dist.barrier()
    dist.destroy_process_group()


def _run(total):
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, total, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.arange(total, dtype=np.float64)
    expect = np.cos(g * 0.37) + 1e-3 * g
    for rank, allc, idx, val in res:
        assert np.array_equal(allc, expect)
        assert np.array_equal(idx, np.argsort(expect, kind="stable")[:3])
    assert np.array_equal(res[0][2], res[1][2])


def test_allgather_even_shards_gloo():
    _run(64)


def test_allgather_uneven_shards_gloo():
    _run(37)
