"""Multi-GPU plumbing, one process per GPU.

* Independent chains (any sampler): split by `chain_split`; no collective.
* One uncollapsed chain over many observations (full / stick-breaking): rows block-partitioned by
  `shard_rows`; the library all-reduces the integer counts c_k, V_kd once per sweep over NCCL
  (bmm_dist_init), theta / pi / alpha are then drawn redundantly on every rank from identical Philox
  counters, so the chain is bit-identical for any number of GPUs.

`init(...)` hands the library's NCCL unique id from rank 0 to the other ranks through an existing
`torch.distributed` process group (any backend: `gloo` on CPU hosts, `nccl` on GPU boxes); torch is
only the rendezvous, the data path never touches it.
"""
import ctypes as C
import os

import numpy as np

from . import _lib


def shard_rows(n_global, world, rank):
    """[lo, hi) of the rows rank `rank` owns: contiguous blocks, sizes differ by at most one, and
    every boundary is even so a Philox draw pair (rows 2m, 2m+1) never straddles two ranks' tiles."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    pairs = (int(n_global) + 1) // 2
    base, extra = divmod(pairs, world)
    lo_p = rank * base + min(rank, extra)
    hi_p = lo_p + base + (1 if rank < extra else 0)
    return min(2 * lo_p, int(n_global)), min(2 * hi_p, int(n_global))


def chain_split(n_chains, world, rank):
    """[lo, hi) of the global chain indices rank `rank` runs (chain c keeps Philox key (seed, c))."""
    base, extra = divmod(int(n_chains), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def exchange_unique_id(get_id, rank, world, group=None):
    """Rank 0 calls `get_id()` (128 bytes); everyone returns those bytes (torch.distributed broadcast)."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = torch.tensor(list(get_id()), dtype=torch.uint8)
    dev = None
    if dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        buf = buf.to(dev)
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().tolist())


def init(rank, world, device, group=None, p2p_cap_ints=1 << 20):
    """Create the library's NCCL communicator for this process (no-op for world == 1) and the peer-memory
    inboxes of the per-sweep count exchange (`p2p_cap_ints` >= K + K*P; BMM_P2P=0: NCCL all-reduce instead).
    Returns True when the peer-memory exchange is in use."""
    L = _lib.lib()
    if world == 1:
        _lib.check(L.bmm_dist_init(0, 1, None, int(device)))
        return False

    def get_id():
        raw = (C.c_uint8 * 128)()
        _lib.check(L.bmm_dist_unique_id(raw))
        return bytes(raw)

    uid = exchange_unique_id(get_id, rank, world, group)
    arr = (C.c_uint8 * 128)(*uid)
    _lib.check(L.bmm_dist_init(int(rank), int(world), arr, int(device)))
    # The count exchange goes over peer memory when the inboxes can be mapped (BMM_P2P=0 forces the NCCL
    # all-reduce): the producer side is the tail of the sweep kernel, the consumer side the head of the
    # parameter-update kernel, so a sweep costs no extra launch and no collective call.
    if p2p_cap_ints and os.environ.get("BMM_P2P", "1") != "0":
        return attach_p2p(rank, world, p2p_cap_ints, group)
    return False


def attach_p2p(rank, world, cap_ints, group=None):
    """Map every rank's count inbox into every other rank (CUDA IPC over NVLink) so the per-sweep count
    exchange is a one-shot push + local sum instead of an NCCL all-reduce.  Falls back silently to NCCL
    if the handles cannot be opened."""
    import torch
    import torch.distributed as dist
    L = _lib.lib()
    L.bmm_dist_p2p_local.argtypes = [C.c_uint64, C.POINTER(C.c_uint8)]
    L.bmm_dist_p2p_attach.argtypes = [C.POINTER(C.c_uint8)]
    h = (C.c_uint8 * 64)()
    rc = L.bmm_dist_p2p_local(int(cap_ints), h)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    mine = torch.tensor([rc == 0] + list(h), dtype=torch.uint8, device=dev)
    allh = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine, group=group)
    allh = torch.stack(allh).cpu().numpy()
    if not allh[:, 0].all():
        L.bmm_dist_p2p_detach()
        return False
    flat = (C.c_uint8 * (64 * world))(*allh[:, 1:].reshape(-1).tolist())
    ok = L.bmm_dist_p2p_attach(flat) == 0
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if not flag.item():
        L.bmm_dist_p2p_detach()
    return bool(flag.item())


def finalize():
    _lib.lib().bmm_dist_finalize()
