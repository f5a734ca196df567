"""Multi-process logic of the N>1 path on CPU: world_size 2, gloo backend.  The data path needs no collective;
what is tested is the slicing, the max-over-ranks timing reduction, the optional final gather, and -- with the kernels
executed by the numpy IR interpreter (tests/cpu_emulation.py) -- that every rank's slice of the sandwich and of the
rigid-body step (sharded by chain) equals the corresponding slice of the single-process / oracle result."""
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


def _compute_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from cpu_emulation import emulate_launches, emulated_context
        from numga_b200.sharding import shard_chains, shard_multivector
        from oracle import numga_oracle as O
        rng = np.random.default_rng(0)                                   # same stream on every rank
        n = 1003
        full, Rv = rng.standard_normal((n, 4)), rng.standard_normal(8)
        want = (O.context((3, 0, 1), np.float64).multivector('even_grade', Rv) >>
                O.context((3, 0, 1), np.float64).multivector('antivector', full)).values
        with emulate_launches():
            ctx = emulated_context((3, 0, 1), torch.float64)
            mine = shard_multivector(ctx.multivector.antivector(full), world, rank)
            out = ctx.multivector.motor(Rv) >> mine                      # this rank's slice, computed
            start, stop = shard_bounds(n, world, rank)
            # (einsum pre-contracts a broadcast operand, so this case is equal to rounding, not bitwise: DESIGN.md)
            assert np.abs(out.numpy() - want[start:stop]).max() <= 1e-12 * np.abs(want).max()
            assert np.abs(gather_values(out.values, dist).numpy() - want).max() <= 1e-12 * np.abs(want).max()
            # config 5 sharded by chain: 5 chains over 2 ranks (3 + 2), one ChainStep launch of two sub-steps each
            from numga_b200.examples.physics import ChainStep, FusedStep, replicate_chains, setup_bodies
            pctx = emulated_context('x+y+z+w0', torch.float64)
            bodies, sets = setup_bodies(pctx, n_bodies=12)
            big, bsets = replicate_chains(bodies, sets, 5)
            big = big.copy(rate=big.rate + pctx.multivector.bivector(0.3 * rng.standard_normal((60, 6))))
            whole = ChainStep(big, bsets, unroll=2, bodies_per_tile=24).integrate(big, 1 / 200)
            local, lsets = shard_chains(big, bsets, world, rank)
            assert local.motor.shape == ((36,) if rank == 0 else (24,))
            got = ChainStep(local, lsets, unroll=2, bodies_per_tile=24).integrate(local, 1 / 200)
            b0 = 0 if rank == 0 else 36
            assert np.array_equal(got.motor.numpy(), whole.motor.numpy()[b0:b0 + local.motor.shape[0]])
            assert np.array_equal(got.rate.numpy(), whole.rate.numpy()[b0:b0 + local.motor.shape[0]])
            assert np.array_equal(gather_values(got.motor.values, dist).numpy(), whole.motor.numpy())
        open(os.path.join(tmp, f'ok{rank}'), 'w').write('ok')
    finally:
        dist.destroy_process_group()


def test_two_ranks_compute_their_slices_and_match_the_oracle(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_compute_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f'ok{r}') for r in range(world))


@pytest.mark.parametrize('n', [1003, 4096])
def test_two_rank_sharding_gather_and_timing(tmp_path, n):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f'ok{r}') for r in range(world))
