"""Batch-axis sharding across the GPUs of one box: one process per GPU, contiguous slices, no collective on
the data path.  torch.distributed is only used for rendezvous, barriers, the max-over-ranks
of timings and the OPTIONAL final gather of output slices.
"""
import torch


def shard_bounds(n_items, world_size, rank, align=4):
    """[start, stop) of rank's contiguous slice of the batch; slice starts are multiples of `align` items so
    every shard keeps 16-byte aligned component rows (128-bit vector kernels)."""
    per = -(-n_items // world_size)
    per = -(-per // align) * align
    start = min(n_items, rank * per)
    return start, min(n_items, start + per)


def shard_multivector(mv, world_size, rank):
    """This rank's slice of a host/device multivector batch (leading batch axis); broadcast operands
    (no batch axis) are replicated by value."""
    if len(mv.shape) == 0:
        return mv
    start, stop = shard_bounds(mv.shape[0], world_size, rank)
    return mv[start:stop]


def shard_chains(bodies, constraint_sets, world_size, rank, bodies_per_chain=None):
    """This rank's share of a block-structured rigid-body system (config 5): whole chains only, so constraints never
    cross ranks and the step needs no exchange.  Body indices of the constraint sets are rebased to the
    local slice.  Returns (bodies, constraint_sets) of the same classes."""
    from numga_b200.examples.physics import Body, Constraint, _block_structure
    n = bodies.motor.shape[0]
    if bodies_per_chain is None:
        bodies_per_chain, _ = _block_structure(n, [c.body_idx for c in constraint_sets])
    n_chains = n // bodies_per_chain
    c0, c1 = shard_bounds(n_chains, world_size, rank, align=1)
    b0, b1 = c0 * bodies_per_chain, c1 * bodies_per_chain
    local = Body(*(getattr(bodies, f)[b0:b1] for f in Body._FIELDS))
    sets = []
    for c in constraint_sets:
        per_chain = c.body_idx.shape[1] // n_chains
        k0, k1 = c0 * per_chain, c1 * per_chain
        sets.append(Constraint(c.body_idx[:, k0:k1] - b0, c.anchors[:, k0:k1], c.compliance[k0:k1]))
    return local, sets


def gather_values(values, dist, total_items=None):
    """OPTIONAL final gather of per-rank SoA value slices [n_local, k] -> [sum n_local, k] on every rank
    (NCCL all_gather over NVLink on GPUs, gloo on CPU in tests).  Not part of the hot path."""
    world = dist.get_world_size()
    n_local = torch.tensor([values.shape[0]], device=values.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    width = max(sizes)
    rows = values.movedim(-1, 0).contiguous()                   # [k, n_local] component rows
    padded = torch.zeros((rows.shape[0], width), device=values.device, dtype=values.dtype)
    padded[:, :rows.shape[1]] = rows
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    full = torch.cat([o[:, :s] for o, s in zip(out, sizes)], dim=1)
    return full.movedim(0, -1)


def max_over_ranks(value, dist, device):
    t = torch.tensor([float(value)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
