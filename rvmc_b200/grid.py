"""Grid-fit entry points of the reference's findBestDistributions.py on the GPU.

The reference file is Python 2, imports a module that is not in its repository and cannot run
(SURVEY.md top box); what is kept here is its public function names, argument meaning, return
tuples, statistics rules (findBestDistributions.py:47-60), best-fit rule (:138-158) and on-disk
formats (:197-201, :276-280, :472-487), restated on top of the B200 kernels:

  * every grid point is one `rvmc_grid_point` (a population + seed); `rvmc_grid_run` evaluates a
    batch of points with the fused sample+Kepler+RV+dispersion kernel (np.std ddof=0, as at :29-30)
    and reduces each point's sigma_1D distribution to mode / mean / median / p05 / p16 / p84 / p95 on
    the device (exact order statistics, NumPy's linear-interpolation percentile rule);
  * with torch.distributed initialised the points are dealt round-robin to the ranks (one process
    per GPU, no data-path collective) and only the 80-byte per-point statistics are all-gathered
    (NCCL over NVLink on GPUs; the same code runs over gloo in the CPU tests).

`create_sig1D_distribution` is the infinite-pool limit of the reference's
choice-from-a-10^5-star-pool bootstrap (`pool=True` gives the finite-pool version).
"""
import ctypes as C
import os
import pickle

import numpy as np

from . import _abi
from . import _device as dv
from ._trace import span

M17_ERRORS = np.array([0.8, 2.3, 1.0, 1.8, 0.8, 1.0, 1.4, 2.5, 2.0, 0.4, 1.5, 0.6])   # findBestDistributions.py:292

STATS_DTYPE = np.dtype([("mode", "f8"), ("mean", "f8"), ("median", "f8"), ("p05", "f8"), ("p16", "f8"),
                        ("p84", "f8"), ("p95", "f8"), ("sigma_max", "f8"), ("n_guard", "u4"),
                        ("n_nonconverged", "u4"), ("n_hist_bins", "u4"), ("reserved", "u4")])
assert STATS_DTYPE.itemsize == C.sizeof(_abi.PointStats) == 80

# rvmc_grid_point as a NumPy record, so that 10^4-10^5 grid points are built by broadcasting instead
# of one ctypes object per point
POINT_DTYPE = np.dtype([(n, "f8") for n, t in _abi.PopParams._fields_ if t is C.c_double] +
                       [("kepler_iters", "i4"), ("substream", "u4"), ("seed", "u8")])
assert POINT_DTYPE.itemsize == C.sizeof(_abi.GridPoint)
POP_FIELDS = tuple(n for n, t in _abi.PopParams._fields_ if t is C.c_double)


def points_array(n, **columns):
    """Record array of n grid points filled with the reference defaults (RVMC.py:28-43), then the
    given columns (scalars or length-n arrays of any rvmc_pop_params field, or `seed`)."""
    d = _abi.PopParams.defaults()
    one = np.zeros(1, dtype=POINT_DTYPE)
    for f in POP_FIELDS:
        one[f] = getattr(d, f)
    one["kepler_iters"] = d.kepler_iters
    a = np.empty(int(n), dtype=POINT_DTYPE)
    a[:] = one[0]                                   # one broadcast of the default record
    for k, v in columns.items():
        if k not in POINT_DTYPE.names:
            raise TypeError("unknown population parameter %r" % k)
        a[k] = v
    return a


def _as_points(pops, seeds):
    if isinstance(pops, np.ndarray) and pops.dtype == POINT_DTYPE:
        a = pops.copy()
        if seeds is not None:
            a["seed"] = np.asarray(seeds, dtype=np.uint64)
        return a
    n = len(pops)
    a = np.empty(n, dtype=POINT_DTYPE)
    if n:
        packed = (_abi.PopParams * n)(*pops)                      # one contiguous ctypes array
        raw = np.frombuffer(packed, dtype=np.uint8).reshape(n, C.sizeof(_abi.PopParams))
        a.view(np.uint8).reshape(n, POINT_DTYPE.itemsize)[:, :C.sizeof(_abi.PopParams)] = raw
    a["seed"] = np.asarray(seeds, dtype=np.uint64)
    return a


# ---------------------------------------------------------------------------------------------
# sharding + gathering (host logic; exercised on CPU with gloo in tests/test_distributed_cpu.py)
# ---------------------------------------------------------------------------------------------
def dist_info(group=None):
    """(rank, world_size) of the current torch.distributed job, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
    except Exception:          # pragma: no cover
        return 0, 1
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_indices(n_points, rank, world, cost=None):
    """Grid points owned by `rank` (int64, in the order the rank evaluates them).

    `cost=None`: all points cost the same -- rank r owns r, r+world, ...
    `cost` = per-point relative cost (one entry per point; the Monte-Carlo work of a point is
    proportional to sample_size * fbin): the points are ordered by descending cost (ties by
    grid index) and dealt round-robin in THAT order.  Every rank then holds the same number of points of
    every cost level to within one, and the total cost of two ranks differs by less than the cost range
    of a single point -- whatever the grid axes' lengths are.  (Dealing the C-order index round-robin is
    balanced only while `world` does not divide the fastest axis: 16 binary fractions over 8 ranks gave
    rank 7 the fractions {0.5, 1.0} and rank 0 {0.0625, 0.5625}, 1.5 : 0.6 in work.)
    The owner of a point depends on (n_points, world, cost) only, so every rank derives the same deal."""
    if cost is None:
        return np.arange(rank, n_points, world, dtype=np.int64)
    cost = np.asarray(cost, dtype=np.float64).ravel()
    if cost.size != n_points:
        raise ValueError("need one cost per grid point")
    order = np.argsort(-cost, kind="stable")
    return order[rank::world].astype(np.int64, copy=False)


def shard_grid_indices(shape, cost_axis, axis_cost, rank, world):
    """shard_indices for a Cartesian grid whose cost depends on ONE axis (the binary fraction): the same
    deal, computed in O(points / world) without touching the other ranks' points.  `shape`: grid shape
    (C order), `cost_axis`: index of the axis that sets the cost, `axis_cost`: cost of each value along it.
    Returns this rank's flat C-order indices."""
    shape = tuple(int(n) for n in shape)
    n_points = int(np.prod(shape))
    n_class = shape[cost_axis]
    n_other = n_points // n_class
    axis_cost = np.asarray(axis_cost, dtype=np.float64).ravel()
    if axis_cost.size != n_class:
        raise ValueError("need one cost per value of the cost axis")
    classes = np.argsort(-axis_cost, kind="stable")            # cost levels, dearest first
    pos = np.arange(rank, n_points, world, dtype=np.int64)      # positions of this rank in the dealt order
    c = classes[pos // n_other]                                 # value index along the cost axis
    j = pos % n_other                                           # rank among the points of that level (C order)
    inner = int(np.prod(shape[cost_axis + 1:], dtype=np.int64))
    return (j // inner) * (n_class * inner) + c * inner + (j % inner)


def _records(byte_tensor):
    """uint8 tensor holding rvmc_point_stats records -> writable NumPy structured array that owns its
    memory (one device-to-host copy, no further host copies)."""
    t = byte_tensor.detach()
    if t.is_cuda:
        return dv.to_host(t).view(STATS_DTYPE)      # pinned destination above 1 MB: one DMA, no staging copy
    return t.clone().numpy().view(STATS_DTYPE)


def combine_point_stats(local_bytes, local, n_points, row_len, world, group=None):
    """The per-point statistics of every rank -> one table in grid order on every rank.

    `local_bytes`: uint8 tensor [n_local * row_len * 80] (CPU for gloo, CUDA for NCCL) with this rank's
    rvmc_point_stats in the order of `local` (its grid indices).  Each rank scatters its records into a
    zeroed [n_points, row_len] table ON ITS DEVICE and the tables are summed as 32-bit integers (one
    all-reduce: every record has exactly one non-zero contributor, so the sum is exact and in grid order);
    one device-to-host copy of the finished table follows.  Host work per rank is O(points / world): no
    per-rank Python loop, no fancy indexing of 80-byte records (7 ms for 65 536 of them)."""
    import torch
    words = STATS_DTYPE.itemsize * row_len // 4
    nl = len(local)
    if world == 1 and nl == n_points and (nl == 0 or (local[0] == 0 and local[-1] == nl - 1 and
                                                      np.all(np.diff(local) == 1))):
        return _records(local_bytes).reshape(n_points, row_len)
    dev = local_bytes.device
    table = torch.zeros((n_points, words), dtype=torch.int32, device=dev)
    if nl:
        idx = torch.from_numpy(np.ascontiguousarray(local, dtype=np.int64)).to(dev)
        table.index_copy_(0, idx, local_bytes[:nl * words * 4].view(torch.int32).view(nl, words))
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(table, group=group)
    return _records(table.view(torch.uint8).view(-1)).reshape(n_points, row_len)


def gather_point_stats(local_bytes, n_points, rank, world, group=None, local=None):
    """combine_point_stats for one record per point; `local` defaults to the uniform-cost deal."""
    if local is None:
        local = shard_indices(n_points, rank, world)
    return combine_point_stats(local_bytes, local, n_points, 1, world, group).reshape(n_points)


def gather_point_stats_rows(local_bytes, n_points, row_len, rank, world, group=None, local=None):
    """combine_point_stats for `row_len` consecutive records per point -> [n_points, row_len]."""
    if local is None:
        local = shard_indices(n_points, rank, world)
    return combine_point_stats(local_bytes, local, n_points, row_len, world, group)


_scratch_cache = {}


def _scratch(lib, dev, ns):
    """The sigma_1D scratch of rvmc_grid_run* (scratch_slots x number_of_samples floats, ~0.5 GB at 10^5
    realizations), kept per (device, stream) between calls: a grid fit calls run_points many times and
    every call would otherwise allocate it afresh.  Calls on one stream are ordered, so reuse is safe."""
    t = dv.torch()
    key = (dev.index, t.cuda.current_stream(dev).cuda_stream)
    need = int(lib.rvmc_grid_scratch_slots(dev.index)) * int(ns)
    buf = _scratch_cache.get(key)
    if buf is None or buf.numel() < need:
        buf = dv.empty((need,), "float32", dev)
        _scratch_cache[key] = buf
    return buf


def point_cost(pops, sample_size):
    """Relative Monte-Carlo cost of grid points: the expected number of binary orbits per realization,
    sample_size * fbin (phases A and C of the fused kernel cost the same for every point)."""
    return float(sample_size) * np.asarray(pops["fbin"], dtype=np.float64)


def _check_indices(pops):
    for f in ("mass_power", "period_pi", "mass_ratio_kappa", "e_power"):
        if np.any(pops[f] == -1.0):
            raise ZeroDivisionError("float division by zero")          # RVMC.py:112


def table_summary(stats, observed_sigma=None, usemode=False):
    """One native pass (rvmc_grid_best_fit_host) over a statistics table in C order -> (flat index of the
    best-fit point, its distance, (sum reserved, sum n_nonconverged, points with n_nonconverged, sum n_guard)).
    The best fit is the rule of findBestDistributions.py:138-158 (first strict minimum of |median - observed|
    below 100, or of |mode - observed| with `usemode`; NaNs never win; index 0 / distance 100 when nothing
    qualifies); with `observed_sigma=None` only the totals are meaningful.  Reading one field of every 80-byte
    record with NumPy is a strided gather per field (~0.2 ms each for 65 536 records); this reads the table once.
    Returns None for a table that is not C-contiguous (the caller falls back to NumPy)."""
    if not (isinstance(stats, np.ndarray) and stats.dtype == STATS_DTYPE and stats.flags.c_contiguous):
        return None
    idx, dist, tot = C.c_int64(0), C.c_double(100.0), (C.c_uint64 * 4)()
    have = observed_sigma is not None
    _abi.check(_abi.lib().rvmc_grid_best_fit_host(stats.ctypes.data, stats.size, int(bool(usemode)), int(have),
                                                  float(observed_sigma) if have else 0.0, C.byref(idx),
                                                  C.byref(dist), tot, dv.host_threads()))
    return int(idx.value), float(dist.value), tuple(int(v) for v in tot)


def _warn_nonconverged(stats, totals=None):
    bad, pts = (totals[1], totals[2]) if totals is not None else \
        (int(stats["n_nonconverged"].sum()), int(np.count_nonzero(stats["n_nonconverged"])))
    if bad:
        import warnings
        # scipy.optimize.newton's RuntimeWarning for array input (RVMC.py:177-180)
        warnings.warn("some failed to converge after 50 iterations (%d orbits in %d grid points)" % (bad, pts),
                      RuntimeWarning)


# ---------------------------------------------------------------------------------------------
# core: evaluate a list of grid points
# ---------------------------------------------------------------------------------------------
def with_streams(pops, seed, common_random_numbers=False):
    """The default random streams of a grid: every point under the SAME seed, told apart by its
    `substream` (the Philox counter word, 0, 1, 2, ... in grid order) -- statistically independent
    Monte-Carlo runs like successive calls of the reference, while the kernels read one launch-uniform
    key schedule instead of deriving ten round keys per point.  `common_random_numbers=True` gives every
    point the same draws (substream 0), which removes most of the point-to-point noise from the statistic
    surfaces.  Returns `pops` (a POINT_DTYPE record array) with both columns set."""
    if not (isinstance(pops, np.ndarray) and pops.dtype == POINT_DTYPE):
        pops = _as_points(pops, np.zeros(len(pops), dtype=np.uint64))          # a list of PopParams
    pops["seed"] = np.uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)
    pops["substream"] = 0 if common_random_numbers else np.arange(len(pops), dtype=np.uint32)
    return pops


def point_seeds(seed, n_points, common_random_numbers=False):
    """Per-point seeds.  Distinct by default (independent Monte-Carlo runs, like successive calls of
    the reference); `common_random_numbers=True` gives every point the same draws, which removes
    most of the point-to-point noise from the statistic surfaces."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    if common_random_numbers:
        return np.full(n_points, seed, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = np.uint64(seed) + np.uint64(0x632BE59BD9B4E019) * np.arange(1, n_points + 1, dtype=np.uint64)
        x = x + np.uint64(0x9E3779B97F4A7C15)                       # splitmix64, vectorised (wraps mod 2^64)
        z = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _local_share(pops, seeds, sample_size, rank, world, local, n_points):
    """(records of this rank's points, their grid indices, n_points).  `pops` is the FULL list / record
    array unless `local` (grid indices of the records in `pops`, which is then the rank's share already)
    and `n_points` are given -- the form grid_search uses so that no rank ever builds the whole grid."""
    if local is not None:
        if not (isinstance(pops, np.ndarray) and pops.dtype == POINT_DTYPE) or len(pops) != len(local):
            raise ValueError("with `local`, pops must be this rank's POINT_DTYPE records")
        _check_indices(pops)
        return np.ascontiguousarray(pops), np.asarray(local, dtype=np.int64), int(n_points)
    if not (isinstance(pops, np.ndarray) and pops.dtype == POINT_DTYPE):
        pops = _as_points(pops, seeds)
        seeds = None
    n_points = len(pops)
    if seeds is not None and len(seeds) != n_points:
        raise ValueError("need one seed per grid point")
    _check_indices(pops)                 # every rank checks the list it was handed, so all ranks raise together
    cost = point_cost(pops, sample_size) if world > 1 else None
    if cost is not None and np.all(cost == cost[0]):
        cost = None
    local = shard_indices(n_points, rank, world, cost)
    pts = np.ascontiguousarray(pops[local])
    if seeds is not None:
        pts["seed"] = np.asarray(seeds, dtype=np.uint64)[local]
    return pts, local, n_points


def _check_errors(measured_errors, sample_size):
    err = np.ascontiguousarray(measured_errors, dtype=np.float32)
    if err.ndim != 1 or err.size != int(sample_size):
        raise ValueError("measured_errors must have sample_size = %d entries (got shape %s)"
                         % (int(sample_size), err.shape))
    return err


def run_points(pops, seeds, measured_errors, sample_size, number_of_samples=10**5, bin=0.5, ddof=0,
               weighted=False, keep_sigma=False, keep_hist=False, hist_bins=2048, device=None, group=None,
               distributed=True, local=None, n_points=None, finish=True):
    """Evaluates grid points (list of PopParams + list of seeds, or POINT_DTYPE records) on the GPU(s).

    With torch.distributed initialised the points are dealt to the ranks by cost (shard_indices:
    cost = sample_size * fbin), each point wholly on one GPU; pass `local` + `n_points` when `pops` already
    is this rank's share.  Returns a dict: 'stats' (structured array [n_points] in grid order, every
    rank), 'sigma' (float32 [n_local, number_of_samples] or None, rows in the order of 'local'), 'hist'
    (uint32 [n_local, hist_bins] or None, same order), 'counters' (binary RVs, guard, non-converged of
    THIS rank), 'local' (grid indices this rank computed)."""
    err = _check_errors(measured_errors, sample_size)
    rank, world = dist_info(group) if distributed else (0, 1)
    pts_np, local, n_points = _local_share(pops, seeds, sample_size, rank, world, local, n_points)
    nl = len(local)
    if not nl:
        pts_np = np.zeros(1, dtype=POINT_DTYPE)
    dev = dv.resolve_device(device)
    ns = int(number_of_samples)
    pts = pts_np.ctypes.data_as(C.POINTER(_abi.GridPoint))
    stats_d = dv.zeros((max(nl, 1) * STATS_DTYPE.itemsize,), "uint8", dev)
    errd = dv.to_device(err, "float32", dev)
    counters = dv.zeros((4,), "int64", dev)
    sigma_d = dv.empty((nl, ns), "float32", dev) if keep_sigma and nl else None
    hist_d = dv.zeros((nl, int(hist_bins)), "int32", dev) if keep_hist and nl else None
    with dv.DeviceGuard(dev):
        lib = _abi.lib()
        scratch = _scratch(lib, dev, ns) if sigma_d is None and nl else None
        with span("grid.points.kernels"):
            _abi.check(lib.rvmc_grid_run(pts, nl, dv.ptr(errd), int(sample_size), int(bool(weighted)), int(ddof), ns,
                                         float(bin), int(hist_bins), dv.ptr(stats_d), dv.ptr(sigma_d), dv.ptr(hist_d),
                                         dv.ptr(scratch), dv.ptr(counters), dv.stream_ptr(dev)))
        with span("grid.combine_stats"):
            stats = combine_point_stats(stats_d, local, n_points, 1, world, group).reshape(n_points)
    out = dict(stats=stats, local=local, counters=dv.to_host(counters), sigma=None, hist=None)
    if sigma_d is not None:
        out["sigma"] = dv.to_host(sigma_d)
    if hist_d is not None:
        out["hist"] = dv.to_host(hist_d).view(np.uint32)
    if finish:               # grid_search folds these checks into its one pass over the table
        _finish(stats)
    return out


def _finish(stats, totals=None):
    """The checks every evaluated table goes through: histograms the statistics kernel could not hold raise,
    non-converged orbits warn.  `totals` = table_summary(stats)[2] when the caller has it already."""
    if totals is None:
        summary = table_summary(stats)
        totals = summary[2] if summary is not None else None
    bad = totals[0] if totals is not None else int(stats["reserved"].sum())
    if bad:
        raise NotImplementedError("%d grid points have sigma_1D histograms wider than the %d bins the statistics "
                                  "kernel holds; use a larger `bin`" % (bad, 4096))
    _warn_nonconverged(stats, totals)


def run_points_fbins(pops, seeds, fbins, measured_errors, sample_size, number_of_samples=10**5, bin=0.5, ddof=0,
                     weighted=False, keep_sigma=False, keep_hist=False, hist_bins=2048, device=None, group=None,
                     distributed=True, local=None, n_points=None, finish=True):
    """Like run_points for the Cartesian product (points x fbins) with each point's seed shared along
    the f_bin axis: rvmc_grid_run_fbins samples and solves every star slot once and reduces once per
    binary fraction -- bit-identical to run_points on the expanded list, at about the cost of the
    largest fraction alone (so every point costs the same and the deal is plain round-robin).  Returns
    'stats' [n_points, n_fbins] (every rank), 'sigma' [n_local, n_fbins, n_real] or None, 'hist'
    [n_local, n_fbins, hist_bins] or None, 'counters', 'local'."""
    fb = np.ascontiguousarray(fbins, dtype=np.float64).ravel()
    if fb.size < 1:
        raise ValueError("need at least one binary fraction")
    err = _check_errors(measured_errors, sample_size)
    rank, world = dist_info(group) if distributed else (0, 1)
    if local is None and isinstance(pops, np.ndarray) and pops.dtype == POINT_DTYPE:
        pops = pops.copy()
        pops["fbin"] = 1.0               # unused along the fused axis: every point costs the same
    pts_np, local, n_points = _local_share(pops, seeds, sample_size, rank, world, local, n_points)
    nl = len(local)
    if not nl:
        pts_np = np.zeros(1, dtype=POINT_DTYPE)
    dev = dv.resolve_device(device)
    ns = int(number_of_samples)
    item = STATS_DTYPE.itemsize
    stats_all = np.empty((n_points, fb.size), dtype=STATS_DTYPE)
    sigma_all = np.empty((nl, fb.size, ns), dtype=np.float32) if keep_sigma and nl else None
    hist_all = np.zeros((nl, fb.size, int(hist_bins)), dtype=np.uint32) if keep_hist and nl else None
    errd = dv.to_device(err, "float32", dev)
    counters = dv.zeros((4,), "int64", dev)
    pts = pts_np.ctypes.data_as(C.POINTER(_abi.GridPoint))
    with dv.DeviceGuard(dev):
        lib = _abi.lib()
        for k0 in range(0, fb.size, 32):                      # the kernel takes <= 32 fractions per pass
            sub = np.ascontiguousarray(fb[k0:k0 + 32])
            nf = sub.size
            stats_d = dv.zeros((max(nl, 1) * nf * item,), "uint8", dev)
            sigma_d = dv.empty((nl, nf, ns), "float32", dev) if keep_sigma and nl else None
            hist_d = dv.zeros((nl, nf, int(hist_bins)), "int32", dev) if keep_hist and nl else None
            scratch = _scratch(lib, dev, ns) if sigma_d is None and nl else None
            with span("grid.fbins.kernels"):
                _abi.check(lib.rvmc_grid_run_fbins(pts, nl, sub.ctypes.data_as(C.POINTER(C.c_double)), nf,
                                                   dv.ptr(errd), int(sample_size), int(bool(weighted)), int(ddof), ns,
                                                   float(bin), int(hist_bins), dv.ptr(stats_d), dv.ptr(sigma_d),
                                                   dv.ptr(hist_d), dv.ptr(scratch), dv.ptr(counters),
                                                   dv.stream_ptr(dev)))
            with span("grid.combine_stats"):
                part = combine_point_stats(stats_d, local, n_points, nf, world, group)
            if nf == fb.size:
                stats_all = part
            else:
                stats_all[:, k0:k0 + nf] = part
            if sigma_d is not None:
                sigma_all[:, k0:k0 + nf] = dv.to_host(sigma_d)
            if hist_d is not None:
                hist_all[:, k0:k0 + nf] = dv.to_host(hist_d).view(np.uint32)
    if finish:
        _finish(stats_all)
    return dict(stats=stats_all, local=local, counters=dv.to_host(counters), sigma=sigma_all, hist=hist_all)


def _pop_for(min_mass, max_mass, fbin, min_period, max_period, sigma_dyn, **extra):
    return _abi.PopParams.defaults(min_mass=float(min_mass), max_mass=float(max_mass), fbin=float(fbin),
                                   min_period=float(min_period), max_period=float(max_period),
                                   sigma_dyn=float(sigma_dyn), **extra)


# ---------------------------------------------------------------------------------------------
# the reference's entry points
# ---------------------------------------------------------------------------------------------
def create_sig1D_distribution(measured_errors=np.zeros(12), number_of_samples=10 ** 5, sample_size=12,
                              number_of_stars=10 ** 5, min_mass=6, max_mass=20, binary_fraction=0.7, min_period=1.4,
                              max_period=3500, sigma_dyn=2.0, seed=None, device=None, pool=False):
    """sigma_1D distribution of `number_of_samples` clusters of `sample_size` stars
    (findBestDistributions.py:21-32; np.std with ddof=0).  `pool=False`: infinite-pool fused kernel
    (`number_of_stars` is then irrelevant); `pool=True`: a `number_of_stars` pool is materialised
    and bootstrapped like the reference does."""
    from .rvmc import RVMC
    err = np.broadcast_to(np.asarray(measured_errors, dtype=np.float64), (int(sample_size),))
    if pool:
        c = RVMC(number_of_stars=number_of_stars, min_mass=min_mass, max_mass=max_mass, fbin=binary_fraction,
                 min_period=min_period, max_period=max_period, sigma_dyn=sigma_dyn, seed=seed, device=device)
        c.calculate_binary_velocities(materialize=False)
        c.calculate_radial_velocity()
        dev = c._device
        poold = c._on_device("radial_velocities")
        out = dv.empty((int(number_of_samples),), "float64", dev)
        errd = dv.to_device(err, "float64", dev)
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_sigma1d_pool(c._next_seed(), 0, dv.ptr(poold), poold.numel(), dv.ptr(errd),
                                                    int(sample_size), 0, 0, 0, int(number_of_samples), dv.ptr(out),
                                                    dv.stream_ptr(dev)))
        return dv.to_host(out)
    dev = dv.resolve_device(device)
    pop = _pop_for(min_mass, max_mass, binary_fraction, min_period, max_period, sigma_dyn)
    out = dv.empty((int(number_of_samples),), "float32", dev)
    errd = dv.to_device(err, "float32", dev)
    sd = dv.fresh_seed() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
    with dv.DeviceGuard(dev):
        _abi.check(_abi.lib().rvmc_sigma1d_fused(C.byref(pop), sd, dv.ptr(errd), int(sample_size), 0, 0, 0,
                                                 int(number_of_samples), dv.ptr(out), None, dv.stream_ptr(dev)))
    return dv.to_host(out).astype(np.float64)


def _stats_tuple(res, axis_values, bin, number_of_samples, keep_sigma):
    st = res["stats"]
    n = len(st)
    all_sig = [None] * n
    values, bins = [None] * n, [None] * n
    for k, i in enumerate(res["local"]):
        nb = int(st["n_hist_bins"][i])
        edges = np.arange(0, st["sigma_max"][i] + bin, bin)           # :47
        bins[i] = edges
        if res["hist"] is not None:
            cnt = res["hist"][k, :nb].astype(np.float64)
            values[i] = cnt / (cnt.sum() * np.diff(edges)) if nb and cnt.sum() else cnt      # density=True
        if keep_sigma and res["sigma"] is not None:
            all_sig[i] = res["sigma"][k].astype(np.float64)
    return (all_sig, list(st["mode"]), list(st["mean"]), list(st["median"]), values, bins, list(st["p05"]),
            list(st["p16"]), list(st["p84"]), list(st["p95"]), list(axis_values))


def get_Pdistr_stats(measured_errors=np.zeros(12), sample_size=12, pmin=1.4, pmax=3300, Npoints=5, bin=0.5,
                     fbin=0.7, min_mass=6, max_mass=20, number_of_samples=10 ** 5, max_period=3500, sigma_dyn=2.0,
                     seed=None, device=None, keep_sigma=True, common_random_numbers=False, verbose=False):
    """Statistics of the sigma_1D distribution for `Npoints` minimum periods log-spaced in
    [pmin, pmax] (findBestDistributions.py:35-71).  Returns the reference's 11-tuple
    (allSig1D, mode, mean, median, Allvalues, Allbins, p05, p16, p84, p95, periods)."""
    periods = np.logspace(np.log10(pmin), np.log10(pmax), Npoints)             # :40
    sd = dv.fresh_seed() if seed is None else seed
    pops = [_pop_for(min_mass, max_mass, fbin, p, max_period, sigma_dyn) for p in periods]
    res = run_points(with_streams(pops, sd, common_random_numbers), None, _errs(measured_errors, sample_size),
                     sample_size, number_of_samples, bin=bin, ddof=0, keep_sigma=keep_sigma, keep_hist=True,
                     device=device)
    if verbose:
        for p, s in zip(periods, res["stats"]):
            print('fbin = %5.2f, Pmin = %5.2f: mode = %5.2f, median = %5.2f, mean = %5.2f, p84 = %5.2f, p16 = %5.2f'
                  % (fbin, p, s["mode"], s["median"], s["mean"], s["p84"], s["p16"]))
    return _stats_tuple(res, periods, bin, number_of_samples, keep_sigma)


def get_fbinDist_stats(measured_errors=np.zeros(12), sample_size=12, fmin=0., fmax=1., Npoints=100, bin=0.5,
                       min_period=1.4, min_mass=6, max_mass=20, number_of_samples=10 ** 5, max_period=3500,
                       sigma_dyn=2.0, seed=None, device=None, keep_sigma=True, common_random_numbers=False,
                       verbose=False):
    """Same for `Npoints` binary fractions in [fmin, fmax] (findBestDistributions.py:74-110)."""
    fbins = np.linspace(fmin, fmax, Npoints)                                   # :78
    sd = dv.fresh_seed() if seed is None else seed
    pops = [_pop_for(min_mass, max_mass, f, min_period, max_period, sigma_dyn) for f in fbins]
    res = run_points(with_streams(pops, sd, common_random_numbers), None, _errs(measured_errors, sample_size),
                     sample_size, number_of_samples, bin=bin, ddof=0, keep_sigma=keep_sigma, keep_hist=True,
                     device=device)
    if verbose:
        for f, s in zip(fbins, res["stats"]):
            print('fbin = %5.2f, Pmin = %5.2f: mode = %5.2f, median = %5.2f, mean = %5.2f, p84 = %5.2f, p16 = %5.2f'
                  % (f, min_period, s["mode"], s["median"], s["mean"], s["p84"], s["p16"]))
    return _stats_tuple(res, fbins, bin, number_of_samples, keep_sigma)


def _errs(measured_errors, sample_size):
    return np.broadcast_to(np.asarray(measured_errors, dtype=np.float64), (int(sample_size),))


def gaussian(x, mean, amplitude, standard_deviation):
    return amplitude * np.exp(- ((x - mean) / standard_deviation) ** 2)


def find_nearest(array, value):
    n = [abs(i - value) for i in array]
    idx = n.index(min(n))
    return idx, array[idx]


def func(x, a, b, c):
    return a + (b * x) + (c * x * x) + (x * x * x)


def best_fit_indices(observed_sigma, mode, median, p05, p16, p84, p95, usemode=False):
    """findBestDistributions.py:129-158: first strict minimum of |statistic - observed| (initial
    distance 100, index 0) for the central statistic and the four percentile curves.
    Returns (index, index05, index16, index84, index95)."""
    central = mode if usemode else median
    out = []
    for curve in (central, p05, p16, p84, p95):
        diff, index = 100, 0
        for i, m in enumerate(curve):
            d = np.abs(m - observed_sigma)
            if d < diff:
                diff, index = d, i
        out.append(index)
    return tuple(out)


def _write_best(path, names, row):
    """The one-row table astropy's ascii.write produced at findBestDistributions.py:197-201."""
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        f.write(" ".join(names) + "\n")
        f.write(" ".join(repr(float(v)) for v in row) + "\n")


def find_best_Pmin(observed_sigma, min_mass, max_mass, sample_size, mode, median, p05, p16, p84, p95, bins, values,
                   period, usemode=False, outdir="results", write=True, verbose=False):
    """Best-fitting minimum period and its percentile-crossing error bars (:127-201).  Writes
    `<outdir>/Pmin_ObsSig_<sigma>_Nstars_<n>_Mass_<a>_<b>.dat` and returns the five periods
    (best, 0.05, 0.16, 0.84, 0.95)."""
    index, i05, i16, i84, i95 = best_fit_indices(observed_sigma, mode, median, p05, p16, p84, p95, usemode)
    row = [period[index], period[i05], period[i16], period[i84], period[i95]]
    if verbose:
        print('Pmin = %5.2f' % row[0])
    if write:
        _write_best(os.path.join(outdir, 'Pmin_ObsSig_' + str(observed_sigma) + '_Nstars_' + str(sample_size) +
                                 '_Mass_' + str(min_mass) + '_' + str(max_mass) + '.dat'),
                    ['best_Pmin', '0.05perc_Pmin', '0.16perc_Pmin', '0.84perc_Pmin', '0.95perc_Pmin'], row)
    return row


def find_best_fbin(observed_sigma, min_mass, max_mass, sample_size, mode, median, p05, p16, p84, p95, bins, values,
                   fbins, usemode=False, outdir="results", write=True, verbose=False):
    """Best-fitting binary fraction (:204-282).  Writes `<outdir>/fbin_ObsSig_...dat`; returns
    `observed_sigma` like the reference.  The five fractions are left in `find_best_fbin.last`."""
    index, i05, i16, i84, i95 = best_fit_indices(observed_sigma, mode, median, p05, p16, p84, p95, usemode)
    row = [fbins[index], fbins[i05], fbins[i16], fbins[i84], fbins[i95]]
    find_best_fbin.last = row
    if verbose:
        print('fbin = %5.2f' % row[0])
    if write:
        _write_best(os.path.join(outdir, 'fbin_ObsSig_' + str(observed_sigma) + '_Nstars_' + str(sample_size) +
                                 '_Mass_' + str(min_mass) + '_' + str(max_mass) + '.dat'),
                    ['best_fbin', '0.05perc_fbin', '0.16perc_fbin', '0.84perc_fbin', '0.95perc_fbin'], row)
    return observed_sigma


def find_Pmin_fbin_individual_clulsters(name_cluster, observed_dispersion, number_stars, min_mass, max_mass,
                                        number_of_samples=10 ** 5, n_pmin=100, n_fbin=200, seed=None, device=None,
                                        outdir="results", verbose=False):
    """Per-cluster best P_min (at fbin 0.7) and best fbin (at P_min 1.4 d) (:285-313).
    Returns {name: dict(Pmin=[5 values], fbin=[5 values])}."""
    out = {}
    sd = dv.fresh_seed() if seed is None else int(seed)
    for k, (n, sig, ns, m0, m1) in enumerate(zip(name_cluster, observed_dispersion, number_stars, min_mass, max_mass)):
        RV_errors = M17_ERRORS if n == 'M17' else np.ones(ns) * 1.2          # :291-294
        t = get_Pdistr_stats(measured_errors=RV_errors, sample_size=ns, pmin=1.4, pmax=3500, Npoints=n_pmin, bin=0.5,
                             fbin=0.7, min_mass=m0, max_mass=m1, number_of_samples=number_of_samples,
                             seed=sd + 2 * k, device=device, keep_sigma=False, verbose=verbose)
        _, mode, mean, median, values, bins, p05, p16, p84, p95, period = t
        pm = find_best_Pmin(sig, m0, m1, ns, mode, median, p05, p16, p84, p95, bins, values, period, usemode=False,
                            outdir=outdir, verbose=verbose)
        t = get_fbinDist_stats(measured_errors=RV_errors, sample_size=ns, fmin=0.00, fmax=1., Npoints=n_fbin, bin=0.5,
                               min_mass=m0, max_mass=m1, number_of_samples=number_of_samples, seed=sd + 2 * k + 1,
                               device=device, keep_sigma=False, verbose=verbose)
        _, mode, mean, median, values, bins, p05, p16, p84, p95, fbins = t
        find_best_fbin(sig, m0, m1, ns, mode, median, p05, p16, p84, p95, bins, values, fbins, usemode=False,
                       outdir=outdir, verbose=verbose)
        out[n] = dict(Pmin=pm, fbin=list(find_best_fbin.last))
    return out


def run_bigGrid(number_stars=20, min_mass=6, max_mass=20, n_fbin=100, n_pmin=100, number_of_samples=10 ** 5,
                seed=None, device=None, outdir="results", write=True, common_random_numbers=False,
                share_fbin_draws=True):
    """The (f_bin x P_min) grid of findBestDistributions.py:457-489 in ONE sharded launch sequence.
    Pickles (binary mode) the reference's list of per-f_bin dicts {values, bins, period, fbin, mean,
    median, mode, p05, p16, p84, p95} and returns it (rank 0 writes the file)."""
    fbins = np.linspace(0, 1, n_fbin)
    periods = np.logspace(np.log10(1.4), np.log10(3500), n_pmin)
    RV_errors = np.ones(number_stars) * 1.2
    sd = dv.fresh_seed() if seed is None else seed
    single = dist_info()[1] == 1
    if share_fbin_draws:
        # one seed per P_min column, shared by its 100 binary fractions: the whole f_bin axis of a
        # column comes out of one pass over its star slots (rvmc_grid_run_fbins)
        pops = points_array(n_pmin, min_mass=float(min_mass), max_mass=float(max_mass), min_period=periods,
                            max_period=3500.0, sigma_dyn=2.0)
        res = run_points_fbins(with_streams(pops, sd, common_random_numbers), None, fbins, RV_errors, number_stars,
                               number_of_samples, bin=0.5, ddof=0, keep_hist=single, device=device)
        st = res["stats"].T                                              # [n_fbin, n_pmin]
        hist = None if res["hist"] is None else np.transpose(res["hist"], (1, 0, 2))
    else:
        pops = [_pop_for(min_mass, max_mass, f, p, 3500, 2.0) for f in fbins for p in periods]
        res = run_points(with_streams(pops, sd, common_random_numbers), None, RV_errors, number_stars,
                         number_of_samples, bin=0.5, ddof=0, keep_hist=single, device=device)
        st = res["stats"].reshape(n_fbin, n_pmin)
        hist = None if res["hist"] is None else res["hist"].reshape(n_fbin, n_pmin, -1)
    l_res = []
    for a, fbin in enumerate(fbins):
        result = dict()
        row = st[a]
        if hist is not None:
            h = hist[a]
            result['values'] = [h[b, :row["n_hist_bins"][b]] / max(1.0, 0.5 * h[b].sum()) for b in range(n_pmin)]
        else:
            result['values'] = None
        result['bins'] = [np.arange(0, row["sigma_max"][b] + 0.5, 0.5) for b in range(n_pmin)]
        result['period'] = list(periods)
        result['fbin'] = fbin
        for key in ('mean', 'median', 'mode', 'p05', 'p16', 'p84', 'p95'):
            result[key] = list(row[key])
        l_res.append(result)
    if write and dist_info()[0] == 0:
        os.makedirs(outdir, exist_ok=True)
        output_name = os.path.join(outdir, "grid_Pmin_fbin_Nstars_{}_Mass_{}_{}.pkl".format(number_stars, min_mass,
                                                                                           max_mass))
        with open(output_name, "wb") as f:
            pickle.dump(l_res, f)
    return l_res


def grid_search(axes, sample_size=12, measured_errors=M17_ERRORS, number_of_samples=10 ** 5, observed_sigma=None,
                base=None, bin=0.5, ddof=0, seed=1, device=None, common_random_numbers=False, group=None,
                usemode=False, share_fbin_draws=True, distributed=True):
    """Generalised grid (BASELINE config 5): `axes` maps population parameter names (any field of
    rvmc_pop_params: 'period_pi', 'mass_ratio_kappa', 'fbin', 'min_period', ...) to 1-D value
    arrays; the Cartesian product is evaluated, sharded over the ranks of the current
    torch.distributed job.  Returns a dict with the per-point statistics reshaped to the grid, the
    aggregated counters, and -- when `observed_sigma` is given -- the best-fit multi-index
    (argmin |median - observed|, first minimum in C order, the rule of :138-158).

    `share_fbin_draws=True` (default): grid points that differ only in `fbin` use the same seed, which
    lets one kernel pass evaluate the whole f_bin axis (rvmc_grid_run_fbins; bit-identical to evaluating
    those points one by one with that seed).  Each cell still gets `number_of_samples` realizations with
    the exact per-cell distribution of the reference; only the noise of neighbouring f_bin cells is
    correlated (which steadies the surface).  `False` gives every cell its own seed."""
    names = list(axes.keys())
    vals = [np.asarray(axes[k], dtype=np.float64) for k in names]
    shape = tuple(len(v) for v in vals)
    fused_axis = names.index("fbin") if (share_fbin_draws and "fbin" in names) else None
    rank, world = dist_info(group) if distributed else (0, 1)
    # indices of -1 raise before any rank builds anything (RVMC.py:112); O(sum of the axis lengths)
    for k, v in list(zip(names, vals)) + list((base or {}).items()):
        if k in ("mass_power", "period_pi", "mass_ratio_kappa", "e_power") and np.any(np.asarray(v) == -1.0):
            raise ZeroDivisionError("float division by zero")
    o_names = [k for k in names if fused_axis is None or k != "fbin"]
    o_vals = [v for k, v in zip(names, vals) if fused_axis is None or k != "fbin"]
    o_shape = tuple(len(v) for v in o_vals)
    n_points = int(np.prod(o_shape)) if o_shape else 1
    _build = span("grid_search.build_share")
    _build.__enter__()
    # this rank's share of the points, by cost: only an un-fused f_bin axis makes points differ in cost
    # (cost = sample_size * fbin); the deal is computed without enumerating the other ranks' points
    if fused_axis is None and "fbin" in o_names and world > 1:
        local = shard_grid_indices(o_shape, o_names.index("fbin"), o_vals[o_names.index("fbin")], rank, world)
    else:
        local = shard_indices(n_points, rank, world)
    # ... and only this rank's records are ever built: column k of point i is axis value multi_index_k(i)
    cols = dict(base or {})
    if o_shape:
        for k, v, mi in zip(o_names, o_vals, np.unravel_index(local, o_shape)):
            cols[k] = v[mi]
    pops = points_array(len(local), **cols)
    pops["seed"] = np.uint64(int(seed) & 0xFFFFFFFFFFFFFFFF)
    pops["substream"] = 0 if common_random_numbers else local.astype(np.uint32)     # == with_streams on the full grid
    _build.__exit__(None, None, None)
    if fused_axis is None:
        res = run_points(pops, None, _errs(measured_errors, sample_size), sample_size, number_of_samples, bin=bin,
                         ddof=ddof, device=device, group=group, distributed=distributed, local=local,
                         n_points=n_points, finish=False)
        stats = res["stats"].reshape(shape)
    else:
        # the f_bin axis shares each point's draws and is evaluated in one pass (rvmc_grid_run_fbins)
        res = run_points_fbins(pops, None, vals[fused_axis], _errs(measured_errors, sample_size), sample_size,
                               number_of_samples, bin=bin, ddof=ddof, device=device, group=group,
                               distributed=distributed, local=local, n_points=n_points, finish=False)
        stats = np.moveaxis(res["stats"].reshape(o_shape + (shape[fused_axis],)), -1, fused_axis)
    _post = span("grid_search.best_fit")
    _post.__enter__()
    out = dict(axes=dict(zip(names, vals)), shape=shape, stats=stats, counters=res["counters"],
               n_local=len(res["local"]), fbin_axis_fused=fused_axis is not None)
    # the table's checks (_finish) and the best fit in ONE native pass over the 80-byte records when the table
    # is in C order (always, unless a fused f_bin axis is not the last one)
    want = observed_sigma is not None
    summary = table_summary(stats, observed_sigma, usemode)
    if summary is None:
        _finish(res["stats"])
    else:
        _finish(stats, summary[2])
    if want:
        # findBestDistributions.py:138-158: first strict minimum below the initial distance of 100, in C
        # order; the strict `<` never selects a NaN (an all-zero sigma_1D distribution has no mode), so
        # NaNs are masked out rather than left to np.argmin, which would return the first of them
        if summary is not None:
            idx, best = summary[0], summary[1]
        else:
            central = np.ascontiguousarray(out["stats"]["mode" if usemode else "median"])   # one strided gather of the field
            d = np.abs(central - observed_sigma)
            ok = d < 100                                   # False for NaN
            best, idx = 100.0, 0
            if ok.any():
                d[~ok] = np.inf
                idx = int(np.argmin(d))                    # first minimum in C order
                best = float(d.flat[idx])
        out["best_index"] = tuple(int(v) for v in np.unravel_index(idx, shape))
        out["best_params"] = {k: float(v[j]) for k, v, j in zip(names, vals, out["best_index"])}
        out["best_distance"] = best
    _post.__exit__(None, None, None)
    return out


def read_Nstars_massRange(path):
    """data_RT20/Nstars_massRange.txt -> list of dicts (name, sigma, Nstars, minMass, maxMass, refs);
    the table plot_grid.py:76 reads with astropy."""
    rows = []
    with open(path) as f:
        header = f.readline().split()
        for line in f:
            parts = line.split(None, 5)
            if len(parts) < 5:
                continue
            rows.append(dict(name=parts[0], sigma=float(parts[1]), Nstars=int(parts[2]), minMass=float(parts[3]),
                             maxMass=float(parts[4]), refs=parts[5].strip() if len(parts) > 5 else ""))
    assert header[:5] == ['name', 'sigma', 'Nstars', 'minMass', 'maxMass']
    return rows


def read_weighted_RVdisp_vs_age(path):
    """data_RT20/weighted_RVdisp_vs_age.txt (';'-separated, '#' comments) -> list of dicts
    (name, age, age_err, sig1D, sig1D_err); the table findBestDistributions.py:338-341 reads."""
    rows = []
    with open(path) as f:
        header = [h.strip() for h in f.readline().split(";")]
        for line in f:
            if line.lstrip().startswith("#") or not line.strip():
                continue
            parts = [p.strip() for p in line.split(";")]
            rows.append(dict(name=parts[0], age=float(parts[1]), age_err=float(parts[2]), sig1D=float(parts[3]),
                             sig1D_err=float(parts[4])))
    assert header == ['name', 'age', 'age_err', 'sig1D', 'sig1D_err']
    return rows


# ---------------------------------------------------------------------------------------------
# P_min(age) relation (findBestDistributions.py:316-397) -- host post-processing of <= 10 points
# ---------------------------------------------------------------------------------------------
def quad_func(p, x):
    """findBestDistributions.py:322-325 (despite its name: a * exp(-x / b) + c)."""
    a, b, c = p
    return a * np.exp(- x / b) + c


def read_best_fit_table(path):
    """One-row table written by find_best_Pmin / find_best_fbin -> dict name -> float."""
    with open(path) as f:
        names = f.readline().split()
        vals = [float(v) for v in f.readline().split()]
    return dict(zip(names, vals))


def plot_Pmin_vs_age(fit=False, results_dir="results", age_table="data_RT20/weighted_RVdisp_vs_age.txt", plot=False):
    """Joins the per-cluster best P_min files with the cluster ages (sigma rounded to 2 decimals,
    findBestDistributions.py:350-359) and, with `fit=True`, fits P_min(t) = a exp(-t/b) + c by
    orthogonal distance regression with a fixed to 1e5, 1e4 and 1e6 (:384-397).  Returns a dict with
    the joined columns and the three (beta, sd_beta) pairs; `plot=True` draws the reference's figure."""
    import glob
    files = sorted(glob.glob(os.path.join(results_dir, 'Pmin_ObsSig_*_Nstars_*.dat')))
    tabs = [read_best_fit_table(f) for f in files]
    sig_file = np.array([float(os.path.basename(f).split('ObsSig_')[1].split('_Nstars')[0]) for f in files])
    ages = read_weighted_RVdisp_vs_age(age_table)
    cols = dict(name=[], age=[], age_err=[], sig1D=[], Pmin=[], Pmin016=[], Pmin084=[])
    for i in range(len(files)):
        for row in ages:
            if round(sig_file[i], 2) == round(row["sig1D"], 2):
                cols["name"].append(row["name"])
                cols["age"].append(row["age"])
                cols["age_err"].append(row["age_err"])
                cols["sig1D"].append(row["sig1D"])
                cols["Pmin"].append(tabs[i]["best_Pmin"])
                cols["Pmin016"].append(tabs[i]["0.16perc_Pmin"])
                cols["Pmin084"].append(tabs[i]["0.84perc_Pmin"])
    out = {k: (np.array(v) if k != "name" else v) for k, v in cols.items()}
    out["Pmin_errs"] = [out["Pmin"] - out["Pmin016"], out["Pmin084"] - out["Pmin"]]
    if fit and len(out["age"]) >= 3:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from scipy import odr
            sy = np.abs(out["Pmin_errs"][0] - out["Pmin_errs"][1])
            sy = np.where(sy > 0, sy, np.min(sy[sy > 0]) if np.any(sy > 0) else 1.0)
            data = odr.RealData(out["age"], out["Pmin"], sx=out["age_err"], sy=sy)
            model = odr.Model(quad_func)
            fits = {}
            for a0 in (1e5, 1e4, 1e6):
                res = odr.ODR(data, model, beta0=[a0, 0.2, 1.3], ifixb=[0, 1, 1]).run()
                fits[a0] = dict(beta=res.beta, sd_beta=res.sd_beta)
        out["fits"] = fits
    if plot:
        import matplotlib.pyplot as plt
        fig = plt.figure(figsize=(4, 4))
        plt.errorbar(out["age"], out["Pmin"], xerr=out["age_err"], yerr=out["Pmin_errs"], fmt='.', color='grey')
        if "fits" in out:
            x = np.linspace(0., 6, 1000)
            plt.plot(x, quad_func(out["fits"][1e5]["beta"], x))
        plt.yscale('log')
        plt.xlabel('age (Myr)')
        plt.ylabel(r'$\rm{P_{min}}$ (log days)')
        out["figure"] = fig
    return out
