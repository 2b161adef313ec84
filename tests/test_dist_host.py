"""Host-side logic of the multi-GPU paths on CPU: partitioning and the world_size-2 rendezvous over gloo."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partition():
    from bmm_mcmc_b200.dist import shard_rows
    for n in (1, 2, 7, 1000, 1001, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_rows(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (a, b), (c, d) in zip(edges, edges[1:]):
                assert b == c and a <= b
            assert all(lo % 2 == 0 or lo == n for lo, _ in edges)
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 3  # pairs differ by one; the odd tail row may be missing


def test_chain_split_partition():
    from bmm_mcmc_b200.dist import chain_split
    for n in (1, 5, 1024, 4096):
        for world in (1, 2, 4, 8):
            edges = [chain_split(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(b == c for (_, b), (c, _) in zip(edges, edges[1:]))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from bmm_mcmc_b200.dist import exchange_unique_id, shard_rows
    dist.init_process_group("gloo", rank=rank, world_size=world)
    uid = exchange_unique_id(lambda: bytes(range(128)), rank, world)
    # N-sharded sufficient statistics: integer counts summed over ranks equal the unsharded counts
    import torch
    rng = np.random.default_rng(0)
    X = (rng.random((1001, 7)) < 0.4).astype(np.int64)
    z = rng.integers(0, 3, 1001)
    lo, hi = shard_rows(1001, world, rank)
    V = np.stack([X[lo:hi][z[lo:hi] == k].sum(0) for k in range(3)])
    t = torch.from_numpy(V.copy())
    dist.all_reduce(t)
    full = np.stack([X[z == k].sum(0) for k in range(3)])
    q.put((rank, uid == bytes(range(128)), bool((t.numpy() == full).all())))
    dist.destroy_process_group()


def test_world2_gloo_rendezvous_and_count_allreduce():
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == [(0, True, True), (1, True, True)]
