"""Multi-GPU sharding of the dispersion path: one process per GPU, independent units (models of a sweep,
MC chains, grid points -- reference point.py:101-105, model3D.py:50-57) split into contiguous blocks, no
collective inside the solver; results are gathered once per block with torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests)."""
import numpy as np


def shard_range(n_units, rank, world):
    """Contiguous block [lo, hi) of rank: the first (n_units % world) ranks get one extra unit."""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def padded_shard_size(n_units, world):
    return (int(n_units) + int(world) - 1) // int(world)


def sharded_forward(solve_fn, layers, nlay, periods, kind=2, group=None, device=None):
    """Solves this rank's block of models and all-gathers (c, u, nfound, flags) so every rank holds the
    full result.

    solve_fn(layers_block, nlay_block, periods, kind) -> dict of torch tensors c[m,K], u[m,K], nfound[m],
    flags[m] on ``device`` (DispersionSolver.forward on the GPU box; a stub in the gloo tests).
    layers: torch tensor [5, M, Lmax] (every rank passes the same global batch, or a view of it).
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return solve_fn(layers, nlay, periods, kind)
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    M = int(layers.shape[1])
    lo, hi = shard_range(M, rank, world)
    blk = padded_shard_size(M, world)
    res = solve_fn(layers[:, lo:hi].contiguous(), nlay[lo:hi].contiguous(), periods, kind)
    K = int(res["c"].shape[1])
    dev = res["c"].device if device is None else device
    out = {}
    for name, width, dtype in (("c", K, torch.float32), ("u", K, torch.float32), ("nfound", 0, torch.int32),
                               ("flags", 0, torch.int32)):
        t = res.get(name)
        if t is None:
            out[name] = None
            continue
        shape = (blk, width) if width else (blk,)
        pad = torch.zeros(shape, dtype=dtype, device=dev)
        pad[: hi - lo] = t
        full = torch.empty((world * blk,) + shape[1:], dtype=dtype, device=dev)
        dist.all_gather_into_tensor(full, pad, group=group)
        # drop the padding of each block
        keep = np.concatenate([np.arange(r * blk, r * blk + (shard_range(M, r, world)[1] - shard_range(M, r, world)[0]))
                               for r in range(world)])
        out[name] = full[torch.as_tensor(keep, device=dev)]
    return out


def gather_chain_rows(rows, group=None):
    """All-gathers per-chain result rows [n_local, 3+P] = (misfit, L, accepted, *params) -- the mcTrack row
    layout of reference point.py:57,73,76 -- from every rank (equal n_local per rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return rows
    world = dist.get_world_size(group)
    full = torch.empty((world * rows.shape[0],) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(full, rows.contiguous(), group=group)
    return full


def best_misfit(local_best, group=None):
    """all_reduce(min) of the best misfit seen so far (scalar tensor)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(local_best, op=dist.ReduceOp.MIN, group=group)
    return local_best
