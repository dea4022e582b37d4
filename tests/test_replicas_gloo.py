"""CPU, world_size 2, gloo: the only N>1 logic of the path -- batch sharding + the final token gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import mgea_b200 as mg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_generate(prompt):             # deterministic stand-in for the engine: ragged, request-dependent output
    return list(prompt) + [(sum(prompt) + i) % 97 for i in range(len(prompt) % 5 + 1)]


def _worker(rank, world, port, n_total, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        prompts = [[i, i + 1, (3 * i) % 11] for i in range(n_total)]
        mine = mg.shard(prompts, rank, world)
        local = [_fake_generate(p) for p in mine]
        full = mg.gather_token_lists(local, n_total)
        arrs = mg.gather_token_lists(local, n_total, as_arrays=True)      # same rows as int32 arrays (bench e2e path)
        assert [a.tolist() for a in arrs] == full and all(str(a.dtype) == "int32" for a in arrs)
        ret[rank] = full
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 8])
def test_shard_generate_gather_world2(n_total):
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, n_total, ret), nprocs=world, join=True)
        want = [_fake_generate([i, i + 1, (3 * i) % 11]) for i in range(n_total)]
        assert ret[0] == want and ret[1] == want


def test_gather_single_process_is_identity():
    assert mg.gather_token_lists([[1, 2], [3]], 2) == [[1, 2], [3]]
    with pytest.raises(ValueError):
        mg.gather_token_lists([[1]], 2)
