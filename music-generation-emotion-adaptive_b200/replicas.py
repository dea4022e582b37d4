"""Multi-GPU layout of the path: full-model replicas, requests sharded by batch, one final token gather.

Every generation request is independent (reference ``sample_kvcache`` is per prompt,
api_cache.py:159-184) and so is every ``classify`` text, so each GPU holds a full replica and a
contiguous slice of the requests; nothing is exchanged on the decode path.  The only communication
is the gather of the finished token lists (``[B_local, T]`` int32, <= 264 KB per GPU at config 3),
done once per job with one ``all_gather`` of a padded tensor (NCCL over NVLink on GPUs, gloo on CPU
in the tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``n_items`` requests owned by ``rank`` (sizes differ by <= 1)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(items: Sequence, rank: int, world_size: int) -> List:
    lo, hi = shard_range(len(items), rank, world_size)
    return list(items[lo:hi])


_GATHER_BUFFERS: dict = {}


def gather_token_lists(local: Sequence[Sequence[int]], n_total: int, group=None,
                       device: Optional[torch.device] = None, as_arrays: bool = False, max_len: Optional[int] = None):
    """All ranks receive the token lists of all ``n_total`` requests in request order.

    ``local`` must be this rank's ``shard_range`` slice.  ONE ``all_gather`` (into views of one contiguous buffer) of an int32 tensor
    ``[max_local, 1 + max_len]`` (column 0 = length, rows padded with -1) and ONE device -> host copy of the gathered block.
    ``max_len``: the known output stride (longest prompt + max_new_tokens).  When given, no length exchange happens at all
    (the caller knows it from its own arguments); when None, one extra ``all_reduce(MAX)`` finds it.
    ``as_arrays=True`` returns one int32 numpy array per request (views of the gathered buffer) instead of Python lists:
    building half a million Python ints costs more than the exchange itself at 8 GPUs.
    """
    import numpy as np
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        if len(local) != n_total:
            raise ValueError("single process: local must hold every request")
        return [np.asarray(x, dtype=np.int32) for x in local] if as_arrays else [list(x) for x in local]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo, hi = shard_range(n_total, rank, world)
    if len(local) != hi - lo:
        raise ValueError(f"rank {rank} holds {len(local)} results, expected {hi - lo}")
    on_gpu = dist.get_backend(group) == "nccl"
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu"))
    max_local = -(-n_total // world)
    my_max = max((len(x) for x in local), default=0)
    if max_len is None:
        t = torch.tensor([my_max], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        max_len = int(t.item())
    elif my_max > max_len:
        raise ValueError(f"a result of {my_max} tokens exceeds max_len={max_len}")
    # staging buffers are reused across calls (a pinned allocation costs milliseconds -- more than the exchange itself)
    key = (max_local, max_len, world, str(dev), on_gpu)
    bufs = _GATHER_BUFFERS.get(key)
    if bufs is None:
        if len(_GATHER_BUFFERS) > 8:
            _GATHER_BUFFERS.clear()
        stage = torch.empty((max_local, 1 + max_len), dtype=torch.int32, pin_memory=on_gpu)
        allo_t = torch.empty((world, max_local, 1 + max_len), dtype=torch.int32, device=dev)
        host = torch.empty((world, max_local, 1 + max_len), dtype=torch.int32, pin_memory=on_gpu)
        bufs = _GATHER_BUFFERS[key] = (stage, allo_t, host)
    stage, allo_t, host = bufs
    nbuf = stage.numpy()
    for i, x in enumerate(local):
        n = len(x)
        nbuf[i, 0] = n
        nbuf[i, 1:1 + n] = x
        nbuf[i, 1 + n:] = -1
    nbuf[len(local):] = -1
    buf = stage.to(dev, non_blocking=True) if on_gpu else stage
    dist.all_gather(list(allo_t.unbind(0)), buf, group=group)   # views of ONE buffer: NCCL gathers in place, no copy-out
    if on_gpu:
        host.copy_(allo_t, non_blocking=True)               # one device -> host copy for all ranks' rows
        torch.cuda.current_stream(dev).synchronize()
        allo = host.numpy().copy() if as_arrays else host.numpy()   # as_arrays hands out views: detach them from the reused buffer
    else:
        allo = allo_t.numpy().copy() if as_arrays else allo_t.numpy()
    result = []
    for r in range(world):
        o = allo[r]
        rlo, rhi = shard_range(n_total, r, world)
        for i in range(rhi - rlo):
            row = o[i, 1:1 + int(o[i, 0])]
            result.append(row if as_arrays else row.tolist())
    return result
