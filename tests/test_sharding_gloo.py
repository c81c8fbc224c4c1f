"""Multi-rank host logic on CPU: world_size-2 gloo processes shard the reference views and exchange image shards with
the same all-gather the NCCL path uses (mve_b200/sharding.py)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mve_b200 import sharding


def test_block_sharding_partition():
    for n in (1, 5, 16, 17, 64, 129):
        for w in (1, 2, 3, 8):
            owned = [sharding.owned_views(n, r, w) for r in range(w)]
            flat = [v for o in owned for v in o]
            assert flat == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
            for r, o in enumerate(owned):
                for v in o:
                    assert sharding.owner_of(v, n, w) == r


def _worker(rank, world, port, n_views, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    owned = sharding.owned_views(n_views, rank, world)
    rng = np.random.default_rng(5)
    full = rng.integers(0, 255, size=(n_views, 6, 8, 3), dtype=np.uint8)      # same on every rank
    local = torch.from_numpy(full[owned])
    got = sharding.all_gather_images(local, world)
    ok = bool((got.numpy() == full).all())
    # per-rank "filled" counters combine by SUM, times by MAX (bench.py)
    t = torch.tensor([float(len(owned))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    q.put((rank, ok, float(t.item())))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_all_gather_images_world2_gloo():
    for n_views in (6, 7):       # even and ragged shards
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_views, q)) for r in range(2)]
        [p.start() for p in procs]
        res = [q.get(timeout=120) for _ in procs]
        [p.join(timeout=60) for p in procs]
        assert all(ok for _, ok, _ in res), res
        assert all(abs(tot - n_views) < 1e-9 for _, _, tot in res)


def _worker_needed(rank, world, port, n_views, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    owned = sharding.owned_views(n_views, rank, world)
    rng = np.random.default_rng(7)
    full = rng.integers(0, 255, size=(n_views, 4, 6, 3), dtype=np.uint8)      # same on every rank
    # every rank needs its own views, its two ring neighbours' border views and one far view
    needed = sorted(set(owned) | {(owned[0] - 1) % n_views, (owned[-1] + 1) % n_views, (owned[0] + n_views // 2) % n_views})
    got, nbytes = sharding.exchange_needed_images(torch.from_numpy(full[owned]), owned, needed, n_views, rank, world)
    ok = sorted(got) == needed and all((got[v].numpy() == full[v]).all() for v in needed)
    expect = sum(1 for v in needed if v not in owned) * 4 * 6 * 3
    q.put((rank, bool(ok), nbytes == expect))
    dist.destroy_process_group()


def test_exchange_needed_images_world3_gloo():
    """Only the needed views travel (point-to-point), whatever the shard sizes."""
    for n_views in (9, 10):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker_needed, args=(r, 3, port, n_views, q)) for r in range(3)]
        [p.start() for p in procs]
        res = [q.get(timeout=120) for _ in procs]
        [p.join(timeout=60) for p in procs]
        assert all(ok and nb for _, ok, nb in res), res
