"""CPU tests of the N > 1 path (world_size 2, gloo): source sharding + the all-reduce of the integer
gridded accumulators.  The per-rank trajectories come from the C oracle (no GPU here); what is under
test is the host logic in rwrt_b200.dist and the order-independence of the integer grids."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity
import rwrt_b200  # noqa: F401
from rwrt_b200 import dist as rdist
from rwrt_b200 import synthetic
from oracle import gridding as ogrid


def test_shard_bounds_partition():
    for n in (0, 1, 7, 8, 63720, 100003):
        for world in (1, 2, 3, 8):
            b = [rdist.shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [s1 - s0 for s0, s1 in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        rdist.shard_bounds(10, 2, 2)


def test_shard_indices_partition():
    for n in (0, 1, 7, 555840):
        for world in (1, 2, 8):
            for mode in ("block", "stride"):
                parts = [rdist.shard_indices(n, r, world, mode) for r in range(world)]
                allidx = np.sort(np.concatenate(parts)) if parts else np.zeros(0)
                assert np.array_equal(allidx, np.arange(n))
                sizes = [len(p) for p in parts]
                assert max(sizes) - min(sizes) <= 1
    assert np.array_equal(rdist.shard_indices(10, 1, 3, "stride"), [1, 4, 7])
    s0, s1 = rdist.shard_bounds(10, 1, 3)
    assert np.array_equal(rdist.shard_indices(10, 1, 3, "block"), np.arange(s0, s1))
    with pytest.raises(ValueError):
        rdist.shard_indices(10, 0, 2, "zigzag")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    u, v = synthetic.clim(144, 73)
    sl, sa = synthetic.random_sources(24, 17)
    zwn = np.array([1.0, 2.0, 3.0])
    s0, s1 = rdist.shard_bounds(len(sl), rank, world)
    tr, acc = parity.oracle_run(u, v, np.radians(sl[s0:s1]), np.radians(sa[s0:s1]), zwn, 6.0)
    out7 = np.stack([tr[n] for n in parity.NAMES])
    grids = [torch.from_numpy(g.copy()) for g in ogrid.accumulate(out7, 2.5, 2.5)]
    rdist.allreduce_grids(*grids)
    steps = torch.tensor([int(acc.sum())])
    dist.all_reduce(steps)
    # the sweep's form: strided shards, count / fx / fy in ONE (3, ncx, ncy) buffer, one all-reduce
    idx = rdist.shard_indices(len(sl), rank, world, "stride")
    tr2, _ = parity.oracle_run(u, v, np.radians(sl[idx]), np.radians(sa[idx]), zwn, 6.0)
    g3 = torch.from_numpy(np.stack(ogrid.accumulate(np.stack([tr2[n] for n in parity.NAMES]), 2.5, 2.5)))
    rdist.allreduce_grids(g3)
    assert all(np.array_equal(g3[k].numpy(), grids[k].numpy()) for k in range(3))
    if rank == 0:
        q.put(([g.numpy() for g in grids], int(steps.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_grids_equal_single_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    grids, steps = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-rank reference: all sources in one batch
    u, v = synthetic.clim(144, 73)
    sl, sa = synthetic.random_sources(24, 17)
    tr, acc = parity.oracle_run(u, v, np.radians(sl), np.radians(sa), np.array([1.0, 2.0, 3.0]), 6.0)
    ref = ogrid.accumulate(np.stack([tr[n] for n in parity.NAMES]), 2.5, 2.5)
    for a, b in zip(grids, ref):
        assert np.array_equal(a, b)            # integer accumulators: bit exact for any sharding
    assert steps == int(acc.sum())
    assert ref[0].sum() > 1000


def test_gridding_oracle_properties():
    """The NumPy gridding definition: counts equal the number of live samples inside the domain."""
    u, v = synthetic.jet_z()
    tr, acc = parity.oracle_run(u, v, np.radians([120.0, 10.0]), np.radians([30.0, -20.0]), np.array([1.0, 2.0]), 8.0)
    out7 = np.stack([tr[n] for n in parity.NAMES])
    count, fx, fy = ogrid.accumulate(out7, 5.0, 5.0)
    live = np.isfinite(out7[0, 1:]) & np.isfinite(out7[1, 1:]) & np.isfinite(out7[3, 1:])
    lon_deg = np.degrees(out7[0, 1:][live])
    inside = (lon_deg >= -360) & (lon_deg < 720)
    assert count.sum() == inside.sum()
    assert count.shape == ogrid.grid_shape(5.0, 5.0) == (216, 36)
    ug = out7[5, 1:][live][inside]
    assert abs(fx.sum() / 1048576.0 - np.nansum(ug)) < 1e-3 * max(1.0, np.abs(ug).sum())
