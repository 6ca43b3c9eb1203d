"""The N>1 host path on CPU: two gloo ranks shard a grid round-robin and all-gather the per-point
statistics; every rank must end with the full table in grid order (rvmc_b200/grid.py)."""
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
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather(tmp_path):
    for n_points in (7, 64):
        port = _free_port()
        mp.spawn(_worker, args=(2, port, n_points, str(tmp_path)), nprocs=2, join=True)
        a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
        np.testing.assert_array_equal(a, b)
