"""GPU parity, part 1: generator, materialised population, Kepler and RV -- through the C ABI.

Oracles: oracle/philox_ref.py (Random123 KATs), oracle/rvmc_oracle.py on the same Philox draws,
and the reference's own outputs stored in tests/golden/*.npz (made by oracle/make_golden.py)."""
import ctypes as C

import numpy as np
import pytest

from oracle import philox_ref as ph
from oracle import rvmc_oracle as orc
from rvmc_b200 import _abi
from rvmc_b200 import _device as dv
from tests import gpu_util as g
from tests.oracle_from_draws import near_regime_edge, orbits_from_draws, rv_from_orbits
from tests.test_philox import KAT

pytestmark = pytest.mark.gpu


def test_device_is_b200_and_library_loaded():
    sms, maj, mnr, khz = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    g.ok(g.lib().rvmc_device_info(g.dev().index, C.byref(sms), C.byref(maj), C.byref(mnr), C.byref(khz)))
    assert (maj.value, mnr.value) == (10, 0) and sms.value >= 100


def test_philox_kat_and_layout_on_gpu():
    # KAT 1: ctr = 0, key = 0 is (seed 0, item 0, stream 0, sub 0)
    out = g.zeros((4,), "int32")
    g.ok(g.lib().rvmc_philox_blocks(0, 0, 1, 0, 0, g.dv.ptr(out), g.stream()))
    assert tuple(g.host(out).view(np.uint32)) == KAT[0][2]
    # KAT 2 / 3 through the counter layout (item_lo, item_hi, stream, sub), key = (seed_lo, seed_hi)
    for ctr, key, want in KAT[1:]:
        item = ctr[0] | (ctr[1] << 32)
        seed = key[0] | (key[1] << 32)
        g.ok(g.lib().rvmc_philox_blocks(seed, item, 1, ctr[2], ctr[3], g.dv.ptr(out), g.stream()))
        assert tuple(g.host(out).view(np.uint32)) == want
    n, seed, first = 5000, 0x1234_5678_9ABC_DEF0, 2 ** 32 - 100
    out = g.zeros((n, 4), "int32")
    g.ok(g.lib().rvmc_philox_blocks(seed, first, n, 6, 3, g.dv.ptr(out), g.stream()))
    want = ph.philox_block(seed, np.arange(first, first + n, dtype=np.uint64), 6, 3)
    np.testing.assert_array_equal(g.host(out).view(np.uint32), want)


def test_normal_f64_matches_box_muller_model():
    n, seed, first = 10001, 99, 7
    out = g.zeros((n,), "float64")
    g.ok(g.lib().rvmc_normal_f64(seed, first, n, 3, 0, 1.5, 2.0, 0, g.dv.ptr(out), g.stream()))
    items = np.arange(first, first + n, dtype=np.uint64)
    w = ph.philox_block(seed, items >> np.uint64(1), 3, 0)
    z0, z1 = ph.box_muller(w[:, 0], w[:, 1])
    want = 1.5 + 2.0 * np.where(items & np.uint64(1), z1, z0)
    np.testing.assert_allclose(g.host(out), want, rtol=0, atol=1e-12)
    g.ok(g.lib().rvmc_normal_f64(seed, first, n, 3, 0, 0.0, 0.0, 1, g.dv.ptr(out), g.stream()))   # accumulate 0
    np.testing.assert_allclose(g.host(out), want, rtol=0, atol=1e-12)


@pytest.mark.parametrize("variant", [{}, dict(min_mass=15., max_mass=60., min_period=1.4, max_period=3500.,
                                              period_pi=0.2, mass_ratio_kappa=-0.3, mass_power=-1.9, min_q=0.2,
                                              max_q=0.95, e_power=0.5)])
def test_sample_orbits_matches_oracle_on_same_draws(variant):
    p = _abi.PopParams.defaults(**variant)
    seed, first, n = 4242, 2 ** 33 + 11, 50000
    names = ("primary_mass", "period", "time", "mass_ratio", "eccentricity", "inclination", "orbit_rotation",
             "eccentric_anomaly")
    tens = {k: g.empty((n,), "float64") for k in names}
    bad = g.zeros((1,), "int32")
    g.ok(g.lib().rvmc_sample_orbits(C.byref(p), seed, first, n, 0, dv.soa(**tens), dv.ptr(bad), g.stream()))
    o = orbits_from_draws(p, seed, np.arange(first, first + n, dtype=np.uint64))
    keep = ~near_regime_edge(o, 1e-12)
    for k in names[:-1]:
        np.testing.assert_allclose(g.host(tens[k])[keep], o[k][keep], rtol=2e-14, atol=1e-300, err_msg=k)
    E, rv, K = rv_from_orbits(o)
    assert np.max(np.abs(g.host(tens["eccentric_anomaly"]) - E)[keep]) < 1e-9
    assert int(g.host(bad)[0]) == 0
    # RVMC.py:182-212 on the GPU arrays
    rvd, Kd = g.empty((n,), "float64"), g.empty((n,), "float64")
    g.ok(g.lib().rvmc_rv_replay_f64(n, dv.soa(**tens), dv.ptr(rvd), dv.ptr(Kd), g.stream()))
    assert g.rel_err(g.host(rvd), rv, K)[keep].max() < 1e-9


def test_sample_orbits_is_partition_invariant():
    p = _abi.PopParams.defaults()
    n = 10000
    full = g.empty((n,), "float64")
    g.ok(g.lib().rvmc_sample_orbits(C.byref(p), 5, 100, n, 0, dv.soa(period=full), None, g.stream()))
    parts = []
    for a, b in ((0, 3333), (3333, 10000)):
        t = g.empty((b - a,), "float64")
        g.ok(g.lib().rvmc_sample_orbits(C.byref(p), 5, 100 + a, b - a, 0, dv.soa(period=t), None, g.stream()))
        parts.append(g.host(t))
    np.testing.assert_array_equal(np.concatenate(parts), g.host(full))


def test_given_period_drives_time_and_eccentricity():
    p = _abi.PopParams.defaults()
    n = 4096
    per = np.full(n, 5.0 * 86400.0)
    per[::2] = 3.0 * 86400.0
    per[1::4] = 100 * 86400.0
    pd = g.up(per, "float64")
    t, e = g.empty((n,), "float64"), g.empty((n,), "float64")
    g.ok(g.lib().rvmc_sample_orbits(C.byref(p), 1, 0, n, 2, dv.soa(period=pd, time=t, eccentricity=e), None,
                                    g.stream()))
    t, e = g.host(t), g.host(e)
    np.testing.assert_array_equal(g.host(pd), per)
    assert np.all((t >= 0) & (t < per))
    assert np.all(e[::2] == 0) and np.all((e[1::4] >= 1e-5) & (e[1::4] <= 0.9))
    mid = per == 5.0 * 86400.0
    assert np.all((e[mid] >= 1e-5) & (e[mid] <= 0.5))


@pytest.mark.parametrize("name", ["rvmc_default_n2048", "rvmc_variant_n1536"])
def test_replay_of_reference_arrays_f64_and_f32(name):
    """The reference's own exported arrays -> GPU Kepler + RV vs the reference's E and RV."""
    z, _ = g.golden(name)
    n = int(z["number_of_stars"])
    tens, soa = g.orbit_soa_from(z)
    K = orc.semi_amplitude(z["mass_ratio"], z["primary_mass"], z["period"], z["eccentricity"], z["inclination"])
    # Kepler, fp64
    E = g.empty((n,), "float64")
    bad = g.zeros((1,), "int32")
    g.ok(g.lib().rvmc_kepler_f64(n, dv.ptr(tens["time"]), dv.ptr(tens["period"]), dv.ptr(tens["eccentricity"]),
                                 dv.ptr(E), dv.ptr(bad), g.stream()))
    assert np.max(np.abs(g.host(E) - z["eccentric_anomaly"])) < 1e-9 and int(g.host(bad)[0]) == 0
    # RV, fp64, from the reference's E: 1e-9 relative (north_star)
    rv = g.empty((n,), "float64")
    g.ok(g.lib().rvmc_rv_replay_f64(n, soa, dv.ptr(rv), None, g.stream()))
    assert g.rel_err(g.host(rv), z["rv_orbit"], K).max() < 1e-9
    # the two intermediate methods
    a, vm = g.empty((n,), "float64"), g.empty((n,), "float64")
    g.ok(g.lib().rvmc_orbit_scales_f64(n, soa, None, dv.ptr(a), dv.ptr(vm), g.stream()))
    np.testing.assert_allclose(g.host(a), z["semi_major_axis"], rtol=1e-13)
    np.testing.assert_allclose(g.host(vm), z["v_max"], rtol=1e-13)
    rv2 = g.empty((n,), "float64")
    g.ok(g.lib().rvmc_rv_from_vmax_f64(n, soa, dv.ptr(vm), dv.ptr(rv2), g.stream()))
    assert g.rel_err(g.host(rv2), z["rv_orbit"], K).max() < 1e-9
    # RV, fp32 fused arithmetic (own Kepler solve): 1e-5 relative
    p = g.pop_from_golden(z)
    rv32, E32 = g.empty((n,), "float32"), g.empty((n,), "float32")
    cnt = g.zeros((2,), "int32")
    g.ok(g.lib().rvmc_rv_replay_f32(n, soa, C.byref(p), dv.ptr(rv32), dv.ptr(E32), dv.ptr(cnt), g.stream()))
    assert g.rel_err(g.host(rv32), z["rv_orbit"], K).max() < 1e-5
    dE = np.abs(g.host(E32).astype(np.float64) - z["eccentric_anomaly"])
    assert np.max(np.minimum(dE, 2 * np.pi - dE)) < 5e-6
    assert tuple(g.host(cnt)) == (0, 0)


def test_kepler_f64_flags_bad_input():
    t = g.up([1.0, np.nan], "float64")
    P = g.up([10.0, 10.0], "float64")
    e = g.up([0.3, 0.3], "float64")
    E = g.empty((2,), "float64")
    bad = g.zeros((1,), "int32")
    g.ok(g.lib().rvmc_kepler_f64(2, dv.ptr(t), dv.ptr(P), dv.ptr(e), dv.ptr(E), dv.ptr(bad), g.stream()))
    assert int(g.host(bad)[0]) == 1


def test_error_conventions():
    p = _abi.PopParams.defaults(period_pi=-1.0)
    out = g.empty((4,), "float64")
    rc = g.lib().rvmc_sample_orbits(C.byref(p), 1, 0, 4, 0, dv.soa(period=out), None, g.stream())
    assert rc == _abi.RVMC_ERR_ZERODIV
    with pytest.raises(ZeroDivisionError):
        _abi.check(rc)
    rc = g.lib().rvmc_sample_orbits(C.byref(_abi.PopParams.defaults()), 1, 0, -1, 0, dv.soa(period=out), None,
                                    g.stream())
    assert rc == _abi.RVMC_ERR_INVALID and b"number of stars" in g.lib().rvmc_last_error()
    # empty input is fine
    g.ok(g.lib().rvmc_sample_orbits(C.byref(_abi.PopParams.defaults()), 1, 0, 0, 0, dv.soa(), None, g.stream()))


def test_to_host_staged_path_equals_the_direct_copy(monkeypatch):
    """Results above PIN_DIRECT_MAX are staged through two fixed pinned buffers into pageable memory
    (reading a multi-GB attribute must not grow the locked memory of the process)."""
    import torch
    x = torch.arange(7 * 428_573, dtype=torch.float32, device=g.dev()).reshape(7, -1) * 0.5
    want = x.cpu().numpy()
    monkeypatch.setattr(dv, "PIN_DIRECT_MAX", 1 << 20)
    monkeypatch.setattr(dv, "PIN_STAGE_BYTES", (1 << 20) + 4096)
    monkeypatch.setattr(dv, "_stage", {})
    got = dv.to_host(x)
    assert got.shape == want.shape and got.dtype == want.dtype
    np.testing.assert_array_equal(got, want)
