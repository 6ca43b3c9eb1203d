"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/rvmc_b200.h
declares, the ctypes mirror matches the C structs, the host-side API validates like the reference,
and the product path refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import rvmc_oracle as orc
from rvmc_b200 import _abi, build, grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rvmc_b200.h")


@pytest.fixture(scope="module")
def lib():
    build.build()                        # no-op when up to date; nvcc cross-compiles without a GPU
    return _abi.lib()


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"RVMC_API\s+[\w\s\*]+?\b(rvmc_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib):
    declared = _declared_symbols()
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_abi.SIGNATURES) == declared          # the ctypes table covers the header, no more no less
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    exported = sorted(set(re.findall(r" T (rvmc_\w+)", out)))
    assert exported == declared                          # -fvisibility=hidden: nothing else leaks


def test_struct_layouts_match_the_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with ctypes."""
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rvmc_b200.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(rvmc_pop_params), offsetof(rvmc_pop_params, fbin),'
                   'offsetof(rvmc_pop_params, kepler_iters), sizeof(rvmc_orbit_soa), sizeof(rvmc_trace_soa),'
                   'sizeof(rvmc_point_stats), sizeof(rvmc_grid_point), offsetof(rvmc_grid_point, seed));return 0;}')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_abi.PopParams), _abi.PopParams.fbin.offset, _abi.PopParams.kepler_iters.offset,
            C.sizeof(_abi.OrbitSoA), C.sizeof(_abi.TraceSoA), C.sizeof(_abi.PointStats), C.sizeof(_abi.GridPoint),
            _abi.GridPoint.seed.offset]
    assert got == want


def test_defaults_match_reference_constructor(lib):
    p = _abi.PopParams()
    lib.rvmc_pop_defaults(C.byref(p))
    q = _abi.PopParams.defaults()
    for name, _ in _abi.PopParams._fields_:
        assert getattr(p, name) == getattr(q, name), name
    # RVMC.py:28-43 / :155-159
    assert (p.min_mass, p.max_mass, p.mass_power, p.period_pi, p.mass_ratio_kappa, p.e_power) == (6, 20, -2.35, -0.5, 0, -0.5)
    assert (p.min_period, p.max_period) == (10 ** 0.15, 10 ** 3.5) and (p.sigma_dyn, p.fbin) == (2.0, 0.7)
    assert (p.e_min, p.e_max_mid, p.e_max_long, p.p_circ_days, p.p_mid_days) == (1e-5, 0.5, 0.9, 4.0, 6.0)
    assert lib.rvmc_abi_version() == 2


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    buf = (C.c_uint32 * 4)()
    rc = lib.rvmc_philox_blocks(1, 0, 1, 0, 0, buf, None)
    assert rc == _abi.RVMC_ERR_NO_DEVICE and b"no CPU fallback" in lib.rvmc_last_error()
    with pytest.raises(_abi.RvmcError, match="no CPU fallback"):
        _abi.check(rc)
    assert lib.rvmc_device_info(0, None, None, None, None) == _abi.RVMC_ERR_NO_DEVICE
    from rvmc_b200.rvmc import RVMC
    from rvmc_b200.bias_corr import BiasCorr
    with pytest.raises(_abi.RvmcError):
        RVMC(number_of_stars=10)
    with pytest.raises(_abi.RvmcError):
        BiasCorr([[0.0, 1.0]], number_of_samples=4)
    with pytest.raises(_abi.RvmcError):
        grid.create_sig1D_distribution(number_of_samples=4)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under rvmc_b200/ may import or execute it."""
    pkg = os.path.join(ROOT, "rvmc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text and "oracle." not in text.replace("the oracle.", ""), f
    code = "import sys; import rvmc_b200, rvmc_b200.rvmc, rvmc_b200.bias_corr, rvmc_b200.grid; " \
           "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_argument_errors_come_before_any_device_work():
    """Validation mirrors the reference and must not need a GPU (RVMC.py:112; RVMC_bias_corr.py:177,201,216,234)."""
    from rvmc_b200.rvmc import RVMC
    from rvmc_b200.bias_corr import BiasCorr
    for kw in (dict(mass_power=-1), dict(period_pi=-1.0), dict(mass_ratio_kappa=-1), dict(e_power=-1)):
        with pytest.raises(ZeroDivisionError):
            RVMC(number_of_stars=10, **kw)
    with pytest.raises(ValueError):
        RVMC(number_of_stars=10, period_dist=np.ones(3))
    t = [np.arange(3.0), np.arange(4.0)]
    with pytest.raises(ValueError, match=r"The number of masses \(1\) and epoch sequences \(2\) are not equal!"):
        BiasCorr(t, masses=[1e31])
    with pytest.raises(ValueError, match="not as expected"):
        BiasCorr(t, mass_dist=np.ones((2, 3)), number_of_samples=4)
    with pytest.raises(ValueError, match="not as expected"):
        BiasCorr(t, period_dist=np.ones((3, 4)), number_of_samples=4)
    with pytest.raises(ValueError, match=r"uncertainties \(3\) does not match"):
        BiasCorr(t, rv_errors=[1.0, 2.0, 3.0])
    with pytest.raises(ZeroDivisionError):
        BiasCorr(t, period_pi=-1)
    with pytest.raises(ZeroDivisionError):
        grid.run_points([_abi.PopParams.defaults(e_power=-1.0)], [1], np.ones(12), 12, 10)
    with pytest.raises(ValueError):
        grid.run_points([_abi.PopParams.defaults()], [1], np.ones(5), 12, 10)


def test_best_fit_rule_and_seeds_and_shards():
    rng = np.random.RandomState(0)
    for _ in range(20):
        curves = [list(rng.uniform(0, 60, 30)) for _ in range(6)]
        obs = rng.uniform(0, 60)
        assert grid.best_fit_indices(obs, *curves) == orc.best_fit_indices(obs, *curves)
        assert grid.best_fit_indices(obs, *curves, usemode=True) == orc.best_fit_indices(obs, *curves, usemode=True)
    far = [500.0] * 4
    assert grid.best_fit_indices(10.0, far, far, far, far, far, far) == (0, 0, 0, 0, 0)    # nothing within 100
    w = grid.with_streams(grid.points_array(5), 2 ** 64 + 7)
    assert w["seed"].tolist() == [7] * 5 and w["substream"].tolist() == [0, 1, 2, 3, 4]
    assert grid.with_streams(grid.points_array(3), 9, True)["substream"].tolist() == [0, 0, 0]
    s = grid.point_seeds(1, 1000)
    assert s.dtype == np.uint64 and len(set(s.tolist())) == 1000 and grid.point_seeds(1, 3, True).tolist() == [1, 1, 1]
    assert np.array_equal(s, grid.point_seeds(1, 1000))
    from rvmc_b200._device import splitmix64
    assert int(s[4]) == splitmix64((1 + 0x632BE59BD9B4E019 * 5) & 0xFFFFFFFFFFFFFFFF)
    for n, w in ((65536, 8), (10, 4), (3, 8), (0, 2)):
        parts = [grid.shard_indices(n, r, w) for r in range(w)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert grid.find_nearest([0.1, 0.69, 0.72], 0.70) == (1, 0.69)


def test_data_rt20_readers(tmp_path):
    a = tmp_path / "Nstars_massRange.txt"
    a.write_text("name sigma Nstars minMass maxMass refs\nIC1805 65.45 8 15 60 \\citet{x}\nM17 5.5 12 6 20 \\citet{paper1}\n")
    rows = grid.read_Nstars_massRange(str(a))
    assert rows[1] == dict(name="M17", sigma=5.5, Nstars=12, minMass=6.0, maxMass=20.0, refs="\\citet{paper1}")
    b = tmp_path / "weighted.txt"
    b.write_text("name ; age ; age_err ; sig1D ; sig1D_err\nIC1805  ; 2.55 ; 0.95 ; 65.4   ; 3.05\n#G333 ; 1 ; 1 ; 1 ; 1\n"
                 "R136    ; 1.5 ; 0.5   ; 25.0                ; 5.9\n")
    rows = grid.read_weighted_RVdisp_vs_age(str(b))
    assert [r["name"] for r in rows] == ["IC1805", "R136"] and rows[1]["sig1D_err"] == 5.9
    ref = "/root/reference/data_RT20"
    if os.path.isdir(ref):          # build container only; the GPU box has no /root/reference
        assert [r["Nstars"] for r in grid.read_Nstars_massRange(ref + "/Nstars_massRange.txt")] == \
            [8, 5, 14, 13, 9, 44, 16, 22, 332, 12]
        assert len(grid.read_weighted_RVdisp_vs_age(ref + "/weighted_RVdisp_vs_age.txt")) == 10


def test_pmin_vs_age_join_and_fit(tmp_path):
    """findBestDistributions.py:328-397: join on sigma rounded to 2 dp, ODR fit of a exp(-t/b) + c."""
    age = tmp_path / "ages.txt"
    rows = [("A", 1.0, 0.5, 10.123), ("B", 2.0, 0.5, 20.456), ("C", 3.0, 0.5, 30.789), ("D", 4.0, 0.6, 41.5),
            ("E", 0.5, 0.2, 7.25)]
    age.write_text("name ; age ; age_err ; sig1D ; sig1D_err\n" +
                   "".join("%s ; %g ; %g ; %r ; 1.0\n" % r for r in rows) + "#F ; 9 ; 1 ; 99 ; 1\n")
    for name, t, _, sig in rows:
        pm = 1e5 * np.exp(-t / 0.4) + 2.0
        grid._write_best(str(tmp_path / ("Pmin_ObsSig_%r_Nstars_12_Mass_6_20.dat" % sig)),
                         ['best_Pmin', '0.05perc_Pmin', '0.16perc_Pmin', '0.84perc_Pmin', '0.95perc_Pmin'],
                         [pm, pm * 0.5, pm * 0.8, pm * 1.3, pm * 2])
    out = grid.plot_Pmin_vs_age(fit=True, results_dir=str(tmp_path), age_table=str(age))
    assert sorted(out["name"]) == ["A", "B", "C", "D", "E"]
    beta = out["fits"][1e5]["beta"]
    assert beta[0] == 1e5 and abs(beta[1] - 0.4) < 1e-3 and abs(beta[2] - 2.0) < 1e-2
    assert grid.quad_func([2.0, 1.0, 3.0], 0.0) == 5.0
    assert grid.read_best_fit_table(str(tmp_path / "Pmin_ObsSig_10.123_Nstars_12_Mass_6_20.dat"))["best_Pmin"] > 0


def test_biascorr_campaign_tables_host_logic():
    """BiasCorr._tables packs n_epochs / t_obs / rv_err into one upload: regular and ragged campaigns,
    scalar, per-star and per-epoch errors (check_error_shape, RVMC_bias_corr.py:374-393).  Pure host
    logic -- exercised on a CPU torch device, no library call involved."""
    import torch
    from rvmc_b200.bias_corr import BiasCorr

    def tables(t_obs, rv_errors):
        bc = object.__new__(BiasCorr)
        object.__setattr__(bc, "_host", {})
        object.__setattr__(bc, "_dev", {})
        object.__setattr__(bc, "_device", torch.device("cpu"))
        bc.__dict__.update(t_obs=t_obs, number_of_stars=len(t_obs), rv_errors=rv_errors, current_star=0)
        return bc._tables()

    reg = [np.arange(6.0) * 86400 + i for i in range(5)]
    nep, max_ep, d_nep, d_tobs, d_err = tables(reg, 2.0)
    assert max_ep == 6 and nep.tolist() == [6] * 5 and d_nep.dtype == torch.int32 and d_tobs.dtype == torch.float64
    np.testing.assert_array_equal(d_tobs.numpy(), np.array(reg))
    np.testing.assert_array_equal(d_err.numpy(), np.full((5, 6), 2.0, np.float32))
    assert tables(np.array(reg), None)[4] is None                         # a 2-D array works as well
    rag = [np.array([0.0, 5.0]), np.array([1.0, 2.0, 3.0, 4.0]), [7.0]]
    nep, max_ep, d_nep, d_tobs, d_err = tables(rag, [1.0, np.array([0.1, 0.2, 0.3, 0.4]), 3.0])
    assert max_ep == 4 and d_nep.tolist() == [2, 4, 1]
    np.testing.assert_array_equal(d_tobs.numpy(), [[0, 5, 0, 0], [1, 2, 3, 4], [7, 0, 0, 0]])
    np.testing.assert_allclose(d_err.numpy(), [[1, 1, 0, 0], [0.1, 0.2, 0.3, 0.4], [3, 0, 0, 0]], rtol=1e-7)
    np.testing.assert_array_equal(tables(rag, 1.5)[4].numpy(), [[1.5, 1.5, 0, 0], [1.5] * 4, [1.5, 0, 0, 0]])
    with pytest.raises(ValueError, match="does not match the number of epochs"):
        tables(rag, [1.0, np.array([0.1, 0.2]), 3.0])
    with pytest.raises(ValueError, match="at least one observing epoch"):
        tables([np.array([1.0]), np.array([])], 1.0)


def test_hot_kernels_resource_usage_and_instruction_mix(lib):
    """Regression guard from the round-1 tuning: the FAST multi-epoch kernel and the single-population
    fused sigma kernel must stay call-free (no stack frame: a callee such as pow() silently raised the
    launch to 128 registers once) and within their occupancy budget; the library must contain no
    tensor-core instruction (the path has no dense contraction) and must be sm_100a code."""
    out = subprocess.run(["cuobjdump", "-res-usage", _abi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    usage = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+) STACK:(\d+)", line)
        if m and name:
            usage[name] = (int(m.group(1)), int(m.group(2)))
    assert "sm_100a" in out
    fast_bc = [k for k in usage if "biascorr_fused_kernelILi6ELb1" in k or "biascorr_fused_kernelILi8ELb1" in k]
    assert len(fast_bc) == 16         # {6, 8 epochs} x {quirk, physical phase} x {RVs written or not} x {errors or not}
    for k in fast_bc:
        assert usage[k][0] <= 85 and usage[k][1] == 0, (k, usage[k])            # 3 blocks of 256 threads per SM
    single = [k for k in usage if "sigma1d_fused_kernelILb1ELb1ELi1ELb0" in k]
    assert len(single) == 1 and usage[single[0]][0] <= 64 and usage[single[0]][1] == 0, usage[single[0]]
    multi = [k for k in usage if "sigma1d_fused_kernelILb0ELb0ELi1ELb0" in k or "sigma1d_fused_kernelILb0ELb1ELi1ELb0" in k]
    assert len(multi) == 2 and all(usage[k][0] <= 64 and usage[k][1] == 0 for k in multi)   # per-point / shared key
    # the f_bin-axis kernel of the grid: 4 blocks of 256 threads per SM need <= 64 registers, and its fast
    # instantiations (shared / per-point key) must not spill (a per-row form of its reduction phase did, and was slower)
    fbins = [k for k in usage if "sigma1d_fused_fbins_kernelILb1ELi1ELb0" in k or "sigma1d_fused_fbins_kernelILb0ELi1ELb0" in k]
    assert len(fbins) == 2 and all(usage[k][0] <= 64 and usage[k][1] == 0 for k in fbins), [usage[k] for k in fbins]
    sass = subprocess.run(["cuobjdump", "-sass", _abi.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert "IMAD.WIDE.U32" in sass and "MUFU.EX2" in sass and "MUFU.LG2" in sass and "DFMA" in sass
    for tensor_op in ("HMMA", "UTCHMMA", "UTCQMMA", "IMMA", "DMMA"):
        assert tensor_op not in sass, tensor_op


def test_host_bit_helpers_match_numpy_unpackbits():
    """rvmc_unpack_bits_host / rvmc_count_bits_host are host code: checked here against NumPy for ragged
    widths, strides wider than the row, sub-blocks of a larger array and several thread counts."""
    lib = _abi.lib()
    rng = np.random.default_rng(3)
    for nr, nc, threads in [(1, 1, 1), (3, 100, 1), (5, 300_007, 4), (1, 64, 2), (7, 31, 3), (2, 4096 * 3 + 5, 8),
                            (4, 262144 * 2 + 17, 5)]:
        nw = (nc + 31) // 32 + 2
        bits = rng.integers(0, 2 ** 32, size=(nr, nw), dtype=np.uint32)
        out = np.full((nr, nc + 5), 7, np.uint8)
        assert lib.rvmc_unpack_bits_host(bits.ctypes.data, nw, nr, nc, out.ctypes.data, nc + 5, threads) == 0
        want = np.unpackbits(bits.view(np.uint8).reshape(nr, -1), axis=1, bitorder="little")[:, :nc]
        np.testing.assert_array_equal(out[:, :nc], want)
        assert (out[:, nc:] == 7).all()                        # nothing written past the row
        counts = np.zeros(nr, np.int64)
        assert lib.rvmc_count_bits_host(bits.ctypes.data, nw, nr, nc, counts.ctypes.data, threads) == 0
        np.testing.assert_array_equal(counts, want.sum(axis=1))
    # a sub-block: rows 1..2, columns 64..(64+1000) of a 4-row array
    nr, nc = 4, 5000
    nw = (nc + 31) // 32
    bits = rng.integers(0, 2 ** 32, size=(nr, nw), dtype=np.uint32)
    out = np.zeros((nr, nc), np.uint8)
    assert lib.rvmc_unpack_bits_host(bits.ctypes.data + 4 * (1 * nw + 2), nw, 2, 1000, out.ctypes.data + nc + 64, nc,
                                     2) == 0
    want = np.unpackbits(bits.view(np.uint8).reshape(nr, -1), axis=1, bitorder="little")[:, :nc]
    np.testing.assert_array_equal(out[1:3, 64:1064], want[1:3, 64:1064])
    assert out.sum() == want[1:3, 64:1064].sum()
    # argument errors
    assert lib.rvmc_unpack_bits_host(bits.ctypes.data, 1, nr, nc, out.ctypes.data, nc, 1) == _abi.RVMC_ERR_INVALID
    assert lib.rvmc_unpack_bits_host(None, nw, nr, nc, out.ctypes.data, nc, 1) == _abi.RVMC_ERR_INVALID


def test_call_seeds_of_neighbouring_user_seeds_are_disjoint():
    """with splitmix64(seed + call) object `s` at call k reused the seed of object s+1 at call
    k-1 (replicate studies over seed = 0..N were silently correlated).  The call counter now enters in a
    different domain: the sequences of 64 neighbouring seeds over 64 calls do not share a single value."""
    from rvmc_b200._device import call_seed
    seen = {}
    for s in list(range(64)) + [2 ** 64 - 1, 2 ** 63, 1234, 1235]:
        for k in range(64):
            v = call_seed(s, k)
            assert 0 <= v < 2 ** 64
            assert v not in seen, (s, k, seen[v])
            seen[v] = (s, k)
    assert call_seed(5, 3) == call_seed(5 + 2 ** 64, 3)           # the seed is taken mod 2^64


def test_trace_spans_are_free_when_off_and_recorded_when_on(caplog):
    """rvmc_b200._trace (NVTX ranges + logging where the reference has bare prints): a span
    costs nothing when tracing is off; when on it is timed, accumulated and logged at DEBUG level."""
    import logging
    import time

    from rvmc_b200 import _trace
    _trace.enable("")
    _trace.totals(reset=True)
    assert _trace.span("x") is _trace.span("y")              # one shared null context
    with _trace.span("off"):
        pass
    assert _trace.totals() == {}
    _trace.enable("1")
    try:
        with caplog.at_level(logging.DEBUG, logger="rvmc_b200"):
            for _ in range(3):
                with _trace.span("phase.a"):
                    time.sleep(0.002)
        tot = _trace.totals(reset=True)
        assert tot["phase.a"][0] == 3 and 0.005 < tot["phase.a"][1] < 0.5
        assert sum("phase.a" in r.getMessage() for r in caplog.records) == 3
        assert _trace.totals() == {}
    finally:
        _trace.enable("")


def test_flag_transfer_resolution_range_weights_and_host_threads(monkeypatch):
    """Host policy of the end-to-end pass (no GPU needed): `auto` resolves to the byte route for a rank that
    has the host to itself and to the packed route when ranks share it; the launch ranges are small at both
    ends; the helper-thread count follows RVMC_HOST_THREADS / the rank's share of the cores."""
    from rvmc_b200 import _device as dv
    from rvmc_b200.bias_corr import BiasCorr
    bc = BiasCorr.__new__(BiasCorr)                      # policy only: no device is touched
    monkeypatch.delenv("LOCAL_WORLD_SIZE", raising=False)
    monkeypatch.setattr(BiasCorr, "flag_transfer", "auto")
    assert bc._flag_mode() == "bytes"
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    assert bc._flag_mode() == "bits"
    for mode in ("bits", "bytes", "hybrid"):
        monkeypatch.setattr(BiasCorr, "flag_transfer", mode)
        assert bc._flag_mode() == mode
    monkeypatch.setattr(BiasCorr, "flag_transfer", "carrier pigeon")
    with pytest.raises(ValueError):
        bc._flag_mode()
    w = BiasCorr._range_weights(12)
    assert len(w) == 12 and np.allclose(w, w[::-1]) and w[0] == w.min() and w[0] / w.sum() < 0.03
    assert np.all(np.diff(w[:6]) > 0) and np.all(np.diff(w[6:]) < 0)
    assert np.allclose(BiasCorr._range_weights(1), [1.0])
    monkeypatch.setenv("RVMC_HOST_THREADS", "5")
    assert dv.host_threads() == 5
    monkeypatch.delenv("RVMC_HOST_THREADS")
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1000")
    assert dv.host_threads() == 1
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "1")
    assert 1 <= dv.host_threads() <= 8


def test_grid_search_builds_only_the_ranks_share(monkeypatch):
    """grid_search must hand each rank exactly its own records (SURVEY 8e / VERDICT r1: O(points / world) host
    work): the records a rank builds equal the rows of the full Cartesian table it owns, with the point's grid
    index as its random substream, for the fused f_bin axis and for independent cells, at several world sizes.
    No GPU: run_points / run_points_fbins are replaced by recorders."""
    from rvmc_b200 import grid
    axes = dict(period_pi=np.linspace(-0.9, 0.4, 5), mass_ratio_kappa=np.linspace(-0.9, 0.9, 3),
                fbin=np.linspace(0.25, 1.0, 4))
    names = list(axes)
    full_mesh = np.meshgrid(*axes.values(), indexing="ij")
    full = grid.points_array(60, **{k: m.ravel() for k, m in zip(names, full_mesh)})
    other = np.meshgrid(axes["period_pi"], axes["mass_ratio_kappa"], indexing="ij")
    calls = {}

    def fake_points(pops, seeds, errs, S, n, **kw):
        calls["points"] = (pops.copy(), kw["local"].copy(), kw["n_points"])
        return dict(stats=np.zeros(kw["n_points"], dtype=grid.STATS_DTYPE), counters=np.zeros(4, np.int64),
                    local=kw["local"])

    def fake_fbins(pops, seeds, fbins, errs, S, n, **kw):
        calls["fbins"] = (pops.copy(), kw["local"].copy(), kw["n_points"], np.asarray(fbins))
        return dict(stats=np.zeros((kw["n_points"], len(fbins)), dtype=grid.STATS_DTYPE),
                    counters=np.zeros(4, np.int64), local=kw["local"])

    monkeypatch.setattr(grid, "run_points", fake_points)
    monkeypatch.setattr(grid, "run_points_fbins", fake_fbins)
    owned_indep, owned_fused = [], []
    for world in (1, 3, 8):
        for rank in range(world):
            monkeypatch.setattr(grid, "dist_info", lambda group=None, r=rank, w=world: (r, w))
            out = grid.grid_search(axes, number_of_samples=10, seed=7, share_fbin_draws=False, observed_sigma=5.0)
            pops, local, n_points = calls["points"]
            assert n_points == 60 and out["shape"] == (5, 3, 4) and len(pops) == len(local)
            for k in names:
                np.testing.assert_array_equal(pops[k], full[k][local])
            np.testing.assert_array_equal(pops["substream"], local.astype(np.uint32))
            assert np.all(pops["seed"] == 7)
            if world == 8:
                owned_indep.append(local)
                # every rank holds each binary fraction (the cost level) the same number of times, to within one
                counts = np.bincount(np.searchsorted(axes["fbin"], pops["fbin"]), minlength=4)
                assert counts.max() - counts.min() <= 1
            out = grid.grid_search(axes, number_of_samples=10, seed=7, observed_sigma=5.0)
            pops, local, n_points, fb = calls["fbins"]
            assert n_points == 15 and out["fbin_axis_fused"]
            np.testing.assert_array_equal(fb, axes["fbin"])
            np.testing.assert_array_equal(pops["period_pi"], other[0].ravel()[local])
            np.testing.assert_array_equal(pops["mass_ratio_kappa"], other[1].ravel()[local])
            np.testing.assert_array_equal(pops["substream"], local.astype(np.uint32))
            if world == 8:
                owned_fused.append(local)
    np.testing.assert_array_equal(np.sort(np.concatenate(owned_indep)), np.arange(60))
    np.testing.assert_array_equal(np.sort(np.concatenate(owned_fused)), np.arange(15))
    # an index of -1 on an axis raises before anything is built or launched (RVMC.py:112)
    calls.clear()
    with pytest.raises(ZeroDivisionError):
        grid.grid_search(dict(period_pi=np.array([-1.0, 0.0]), fbin=np.array([0.5])), number_of_samples=10)
    assert not calls


def test_table_summary_is_the_reference_best_fit_rule_in_one_pass(monkeypatch):
    """rvmc_grid_best_fit_host (host code, no GPU): the best-fit rule of findBestDistributions.py:138-158 -- first
    strict minimum below the initial distance of 100, NaNs never win -- and the table's sanity totals, in one pass
    over the 80-byte records; checked against best_fit_indices (the Python statement of the rule) and against the
    NumPy gathers it replaces, for ties, NaNs, nothing-below-100 and every thread count; then through grid_search
    for a C-ordered table (native pass) and for one whose fused f_bin axis is not the last (NumPy fallback)."""
    from rvmc_b200 import grid
    rng = np.random.default_rng(11)
    for n in (1, 7, 4096, 4097, 65536 + 13):
        st = np.zeros(n, dtype=grid.STATS_DTYPE)
        st["median"] = rng.uniform(5, 60, n)
        st["mode"] = np.round(rng.uniform(5, 60, n) * 2) / 2 + 0.25          # many exact ties, like real modes
        st["median"][rng.random(n) < 0.1] = np.nan
        st["n_nonconverged"] = rng.integers(0, 3, n) * (rng.random(n) < 0.01)
        st["n_guard"] = rng.integers(0, 5, n)
        z = np.zeros(n)
        for usemode in (False, True):
            for obs in (25.0, 25.25, -200.0, 1e4):
                idx, dist, tot = grid.table_summary(st, obs, usemode)
                want = grid.best_fit_indices(obs, st["mode"], st["median"], z, z, z, z, usemode=usemode)[0]
                assert idx == want, (n, usemode, obs)
                c = st["mode" if usemode else "median"][idx]
                assert dist == (abs(c - obs) if abs(c - obs) < 100 else 100.0)
                assert tot == (0, int(st["n_nonconverged"].sum()), int(np.count_nonzero(st["n_nonconverged"])),
                               int(st["n_guard"].sum()))
        for threads in (1, 2, 5):
            monkeypatch.setenv("RVMC_HOST_THREADS", str(threads))
            assert grid.table_summary(st, 25.0)[0] == grid.best_fit_indices(25.0, st["mode"], st["median"], z, z, z, z)[0]
        monkeypatch.delenv("RVMC_HOST_THREADS")
    st = np.zeros(10, dtype=grid.STATS_DTYPE)
    st["median"] = np.nan
    assert grid.table_summary(st, 25.0)[:2] == (0, 100.0)                   # nothing qualifies: index 0, distance 100
    assert grid.table_summary(st)[2] == (0, 0, 0, 0)                         # totals only
    assert grid.table_summary(st[::2]) is None                               # not C-contiguous: caller falls back
    assert grid.table_summary(np.zeros(0, dtype=grid.STATS_DTYPE), 1.0)[:2] == (0, 100.0)
    st["reserved"][3] = 1
    with pytest.raises(NotImplementedError):
        grid._finish(st)
    st["reserved"][3] = 0
    st["n_nonconverged"][[2, 5]] = (3, 4)
    with pytest.warns(RuntimeWarning, match=r"7 orbits in 2 grid points"):
        grid._finish(st)
    with pytest.warns(RuntimeWarning, match=r"7 orbits in 2 grid points"):
        grid._finish(st[::1], totals=None)

    # through grid_search: fbin last (C-ordered table, native pass) and fbin first (moveaxis -> NumPy fallback)
    def fake_fbins(pops, seeds, fbins, errs, S, n, **kw):
        assert kw["finish"] is False                  # grid_search folds the checks into its own pass
        t = np.zeros((kw["n_points"], len(fbins)), dtype=grid.STATS_DTYPE)
        t["median"] = 10.0 + np.arange(t.size).reshape(t.shape) % 17
        t["median"][0, 0] = np.nan
        return dict(stats=t, counters=np.zeros(4, np.int64), local=kw["local"])
    monkeypatch.setattr(grid, "run_points_fbins", fake_fbins)
    monkeypatch.setattr(grid, "dist_info", lambda group=None: (0, 1))
    for order in (("period_pi", "fbin"), ("fbin", "period_pi")):
        axes = {k: (np.linspace(-0.9, 0.4, 6) if k == "period_pi" else np.linspace(0.25, 1.0, 4)) for k in order}
        out = grid.grid_search(axes, number_of_samples=10, seed=7, observed_sigma=13.0)
        med = out["stats"]["median"]
        assert med.shape == out["shape"] == tuple(len(axes[k]) for k in order)
        d = np.abs(med - 13.0)
        d[np.isnan(d)] = np.inf
        assert out["best_index"] == tuple(int(v) for v in np.unravel_index(np.argmin(d), med.shape))
        assert out["best_distance"] == 0.0
