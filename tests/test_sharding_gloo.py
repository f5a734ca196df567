"""Multi-process logic of the N>1 path on CPU: world_size 2, gloo backend.  The data path needs no collective;
what is tested is the slicing, the max-over-ranks timing reduction and the optional final gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from numga_b200.sharding import gather_values, max_over_ranks, shard_bounds


def test_shard_bounds_cover_the_batch_and_stay_aligned():
    for n in (0, 1, 5, 1000, 1003, 1 << 20, (1 << 28) + 7):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert all(a % 4 == 0 for a, b in spans if b > a)


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from numga_b200 import B200Context
        from numga_b200.sharding import shard_multivector
        ctx = B200Context((3, 0, 1), device='cpu', compile_only=True)
        full = np.arange(n * 4, dtype=np.float32).reshape(n, 4)
        p = ctx.multivector.antivector(full)
        R = ctx.multivector.motor(np.ones(8, np.float32))
        mine = shard_multivector(p, world, rank)
        assert shard_multivector(R, world, rank) is R                  # broadcast operand: replicated by value
        start, stop = shard_bounds(n, world, rank)
        assert mine.shape == (stop - start,)
        assert np.array_equal(mine.numpy(), full[start:stop])
        out = (R >> mine)                                              # compile-only: right shape, no launch
        assert out.shape == mine.shape
        gathered = gather_values(mine.values, dist)                    # optional final gather (gloo here, NCCL on GPUs)
        assert np.array_equal(gathered.numpy(), full)
        slow = max_over_ranks(10.0 + rank, dist, 'cpu')                # timing reduction used by bench.py
        assert slow == 10.0 + world - 1
        open(os.path.join(tmp, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [1003, 4096])
def test_two_rank_sharding_gather_and_timing(tmp_path, n):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f'ok{r}') for r in range(world))
