"""CPU tier: the N>1 sharding / gather path with world_size 2 over gloo (the solver is stubbed: the
collective plumbing, block boundaries and padding are what is tested here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pysurfinv_b200.distributed import sharded_forward, gather_chain_rows, best_misfit, shard_range


def _stub_solve(layers, nlay, periods, kind):
    # deterministic function of the inputs so the gathered result can be checked against a global solve
    K = len(periods)
    per = torch.as_tensor(np.asarray(periods, dtype=np.float32))
    c = layers[1, :, :1] + 0.01 * per[None, :] + kind
    return {"c": c.float(), "u": (0.9 * c).float(), "nfound": torch.full((layers.shape[1],), K, dtype=torch.int32),
            "flags": nlay.to(torch.int32) * 0}


def _worker(rank, world, port, M, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    layers = torch.rand((5, M, 6), generator=g)
    nlay = torch.full((M,), 6, dtype=torch.int32)
    per = [8.0, 16.0, 32.0]
    out = sharded_forward(_stub_solve, layers, nlay, per, kind=2)
    ref = _stub_solve(layers, nlay, per, 2)
    ok = all(torch.equal(out[k], ref[k]) for k in ("c", "u", "nfound", "flags"))
    lo, hi = shard_range(M, rank, world)
    rows = torch.full((4, 5), float(rank))
    allrows = gather_chain_rows(rows)
    ok = ok and allrows.shape == (4 * world, 5) and float(allrows[4 * (world - 1), 0]) == world - 1
    b = best_misfit(torch.tensor([1.0 + rank]))
    ok = ok and float(b) == 1.0
    q.put((rank, bool(ok), lo, hi))
    dist.destroy_process_group()


@pytest.mark.parametrize("M", [7, 8])
def test_world2_gloo(M):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, M, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert all(r[1] for r in res)
    assert res[0][2] == 0 and res[0][3] == res[1][2] and res[1][3] == M
