"""world_size-2 gloo run on CPU of the host-side multi-process plumbing: every rank derives the same slab plan
and the per-rank output ranges tile the z-axis (no GPU, no NCCL)."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ndsm_b200.dist import slab_range
    from ndsm_b200.mg import Plan
    n = 129
    x = np.linspace(0, 1, n)
    p = Plan([x, x.copy(), x.copy()])
    ndist, zs = p.slab_partition(world, 16)
    k0, k1 = slab_range(n, world, rank)
    mine = torch.tensor([k0, k1, ndist] + [int(v) for v in zs.ravel()], dtype=torch.int64)
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    ok = all(torch.equal(a[2:], allv[0][2:]) for a in allv)           # identical plan on every rank
    ranges = sorted((int(a[0]), int(a[1])) for a in allv)
    ok = ok and ranges[0][0] == 0 and ranges[-1][1] == n and all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    ok = ok and (k0, k1) == (int(zs[0][rank]), int(zs[0][rank + 1]))    # output range == finest slab
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_slab_plan_is_consistent_across_ranks():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29511 + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
