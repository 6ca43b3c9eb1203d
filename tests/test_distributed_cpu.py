"""The N>1 host path on CPU: two gloo ranks shard a grid (plain round-robin and the cost-aware deal) and
combine the per-point statistics; every rank must end with the full table in grid order
(rvmc_b200/grid.py).  Also the deal itself: every point owned once, cost spread below 5 %."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_points, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rvmc_b200 import grid
    assert grid.dist_info() == (rank, world)
    local = grid.shard_indices(n_points, rank, world)
    st = np.zeros(len(local), dtype=grid.STATS_DTYPE)
    st["median"] = local * 1.5            # what a rank would have computed for its points
    st["mode"] = local + 0.25
    st["n_hist_bins"] = local
    raw = torch.from_numpy(st.view(np.uint8).copy())
    full = grid.gather_point_stats(raw, n_points, rank, world)
    assert full.shape == (n_points,)
    np.testing.assert_array_equal(full["median"], np.arange(n_points) * 1.5)
    np.testing.assert_array_equal(full["mode"], np.arange(n_points) + 0.25)
    np.testing.assert_array_equal(full["n_hist_bins"], np.arange(n_points))
    np.save(os.path.join(out_dir, "r%d.npy" % rank), full["median"])
    # the cost-aware deal of a (3 x 5 x 4) grid whose last axis sets the cost, rows of 4 records per point
    shape, fb = (3, 5, 4), np.array([0.25, 1.0, 0.5, 0.75])
    n = int(np.prod(shape))
    cost = fb[np.unravel_index(np.arange(n), shape)[2]]
    mine = grid.shard_grid_indices(shape, 2, fb, rank, world)
    np.testing.assert_array_equal(mine, grid.shard_indices(n, rank, world, cost))
    st = np.zeros((len(mine), 4), dtype=grid.STATS_DTYPE)
    st["p84"] = mine[:, None] * 10.0 + np.arange(4)[None, :]
    st["n_nonconverged"] = mine[:, None]
    raw = torch.from_numpy(st.view(np.uint8).reshape(-1).copy())
    full = grid.combine_point_stats(raw, mine, n, 4, world)
    assert full.shape == (n, 4)
    np.testing.assert_array_equal(full["p84"], np.arange(n)[:, None] * 10.0 + np.arange(4)[None, :])
    np.testing.assert_array_equal(full["n_nonconverged"], np.repeat(np.arange(n), 4).reshape(n, 4))
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    for n_points in (7, 64):
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_points, str(tmp_path)), nprocs=2, join=True)
        a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
        np.testing.assert_array_equal(a, b)


def test_cost_aware_deal_balances_every_world_size():
    """grid points dealt by cost (cost = sample_size * fbin).  On the C5 grid (64 x 64 x 16, f_bin
    fastest) plain round-robin over 8 ranks aliases with the f_bin axis (rank 7 would own f_bin {0.5, 1.0}
    only); the cost-aware deal keeps the per-rank cost within 5 % for every world size, owns every point
    exactly once, and does not depend on which rank computes it."""
    sys.path.insert(0, ROOT)
    from rvmc_b200 import grid
    shape = (64, 64, 16)
    fb = np.linspace(0.0625, 1, 16)
    n = int(np.prod(shape))
    cost = 12 * fb[np.unravel_index(np.arange(n), shape)[2]]
    for world in (1, 2, 3, 4, 5, 7, 8):
        parts = [grid.shard_grid_indices(shape, 2, fb, r, world) for r in range(world)]
        np.testing.assert_array_equal(np.sort(np.concatenate(parts)), np.arange(n))
        tot = np.array([cost[p].sum() for p in parts])
        assert (tot.max() - tot.min()) / tot.mean() < 0.05, (world, tot)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
        for r in (0, world - 1):
            np.testing.assert_array_equal(parts[r], grid.shard_indices(n, r, world, cost))
    # the aliasing the plain deal has at 8 ranks (what round 1 shipped): > 40 % spread
    plain = np.array([cost[grid.shard_indices(n, r, 8)].sum() for r in range(8)])
    assert (plain.max() - plain.min()) / plain.mean() > 0.4
    # arbitrary costs (no grid structure): spread below the cost range of one point
    rng = np.random.RandomState(0)
    c = rng.uniform(1, 12, 1000)
    tot = np.array([c[grid.shard_indices(1000, r, 8, c)].sum() for r in range(8)])
    assert tot.max() - tot.min() <= c.max() - c.min()
