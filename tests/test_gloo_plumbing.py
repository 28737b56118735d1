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


def _heap_ops(seed=3, n=400):
    """A call sequence like the one a solve makes on the symmetric heap: hierarchy arenas, inboxes and face buffers
    of assorted sizes allocated and freed in interleaved order."""
    rng = np.random.default_rng(seed)
    ops, live = [], []
    for i in range(n):
        if live and rng.random() < 0.45:
            j = live.pop(int(rng.integers(len(live))))
            ops.append(-(j + 1))
        else:
            ops.append(int(rng.choice([8, 520, 4096, 2 ** 20 + 24, 13 * 2 ** 20, 96 * 2 ** 20])))
            live.append(i)
    return np.array(ops, dtype=np.int64)


def _sym_heap_offsets(ops, seg=1 << 30):
    from ndsm_b200 import load_library
    lib = load_library()
    out = np.zeros(len(ops), dtype=np.int64)
    rc = lib.ndsm_b200_plan_sym_heap(seg, ops.ctypes.data, len(ops), out.ctypes.data)
    assert rc == 0
    return out


def test_symmetric_heap_allocator_never_overlaps_and_reuses_freed_space():
    """csrc/sym_alloc.hpp (the offset allocator behind the peer-memory transport): live blocks never overlap, freed
    space is coalesced and handed out again, offsets are 512-byte aligned."""
    ops = _heap_ops()
    off = _sym_heap_offsets(ops)
    live = {}
    for i, op in enumerate(ops):
        if op > 0:
            assert off[i] >= 0 and off[i] % 512 == 0
            size = (int(op) + 511) // 512 * 512
            for (o, s) in live.values():
                assert off[i] + size <= o or o + s <= off[i], "overlap"
            live[i] = (int(off[i]), size)
        else:
            live.pop(int(-op - 1))
    # after everything is freed the whole segment is one block again: the largest request fits at offset 0
    tail = np.array([-(i + 1) for i in live] + [1 << 29], dtype=np.int64)
    allops = np.concatenate([ops, tail])
    assert _sym_heap_offsets(allops)[-1] == 0
    # a request larger than the segment reports failure instead of an offset
    assert _sym_heap_offsets(np.array([(1 << 30) + 1], dtype=np.int64))[0] == -1


def _heap_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off = torch.from_numpy(_sym_heap_offsets(_heap_ops()))
    allv = [torch.zeros_like(off) for _ in range(world)]
    dist.all_gather(allv, off)
    q.put((rank, bool(all(torch.equal(a, allv[0]) for a in allv))))
    dist.destroy_process_group()


def test_symmetric_heap_layout_is_identical_on_every_rank():
    """Peer addresses are computed as base[peer] + (p - base[me]): that only works if every rank, replaying the same
    allocation sequence, arrives at the same offsets."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29711 + (os.getpid() % 200)
    procs = [ctx.Process(target=_heap_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
