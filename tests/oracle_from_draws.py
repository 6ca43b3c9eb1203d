"""Test helper: regenerate on the CPU, from (seed, item), the Philox draws a kernel consumed and
push them through the oracle -- the 'same exported uniform draws' leg of the parity protocol."""
import numpy as np

from oracle import philox_ref as ph
from oracle import rvmc_oracle as orc


def pop_kwargs(p):
    return {n: getattr(p, n) for n, _ in p._fields_}


def orbits_from_draws(p, seed, items):
    """Oracle orbit arrays (reference units) for global items under population `p` (PopParams)."""
    wa = ph.philox_block(seed, items, ph.STREAM_ORBIT_A)
    wb = ph.philox_block(seed, items, ph.STREAM_ORBIT_B)
    ua = ph.u01(wa)
    ub = ph.u01(wb)
    assert (p.e_min, p.e_max_mid, p.e_max_long, p.p_circ_days, p.p_mid_days) == (1e-5, 0.5, 0.9, 4.0, 6.0)
    o = {}
    o["primary_mass"] = orc.masses_from_uniform(ua[:, 0], p.min_mass, p.max_mass, p.mass_power)
    o["period"] = orc.period_from_uniform(ua[:, 1], p.min_period, p.max_period, p.period_pi)
    o["time"] = orc.time_from_uniform(ua[:, 2], o["period"])
    o["mass_ratio"] = orc.mass_ratio_from_uniform(ua[:, 3], p.min_q, p.max_q, p.mass_ratio_kappa)
    o["eccentricity"] = orc.eccentricity_from_uniform_slotted(o["period"], ub[:, 0], p.e_power)
    o["inclination"] = orc.inclination_from_uniform(ub[:, 1])
    o["orbit_rotation"] = orc.orbit_rotation_from_uniform(ub[:, 2])
    o["u"] = np.concatenate([ua, ub], axis=1)
    return o


def rv_from_orbits(o):
    """Reference E (SciPy-style secant) and orbital RV / K for oracle orbit arrays."""
    E = orc.eccentric_anomaly_single_epoch(o["time"], o["period"], o["eccentricity"])
    rv = orc.orbital_rv(o["primary_mass"], o["period"], o["mass_ratio"], o["eccentricity"],
                        o["inclination"], o["orbit_rotation"], E)
    K = orc.semi_amplitude(o["mass_ratio"], o["primary_mass"], o["period"], o["eccentricity"],
                           o["inclination"])
    return E, rv, K


def near_regime_edge(o, tol=2e-6):
    """Orbits whose log10 P sits within `tol` of the 4 d / 6 d eccentricity switches: fp32 and
    fp64 may legitimately land on different sides (documented exclusion)."""
    x = np.log10(o["period"] / 86400.0)
    return (np.abs(x - np.log10(4.0)) < tol) | (np.abs(x - np.log10(6.0)) < tol)


# ---- bootstrap / fused-slot / multi-epoch draws ------------------------------------------------
def pool_bootstrap_from_draws(seed, call_id, pool, err, S, first_real, n_real):
    """The (n_real, S) matrix `choice(pool) + normal(0, err)` of RVMC.py:264-271 for the draws
    rvmc_sigma1d_pool consumes: one STREAM_BOOT block per pair of star slots."""
    slots = np.arange(first_real * S, (first_real + n_real) * S, dtype=np.uint64)
    w = ph.philox_block(seed, slots >> np.uint64(1), ph.STREAM_BOOT, call_id)
    odd = (slots & np.uint64(1)).astype(bool)
    word = np.where(odd, w[:, 1], w[:, 0]).astype(np.uint64)
    idx = (word * np.uint64(len(pool))) >> np.uint64(32)
    z0, z1 = ph.box_muller(w[:, 2], w[:, 3])
    j = (slots % np.uint64(S)).astype(np.int64)
    v = np.asarray(pool)[idx.astype(np.int64)] + np.where(odd, z1, z0) * np.asarray(err, dtype=np.float64)[j]
    return v.reshape(n_real, S)


def fused_slots_from_draws(p, seed, err, S, first_slot, n):
    """What rvmc_sigma1d_fused computes per star slot, in fp64 from the same Philox draws:
    coin, noise N(0, sqrt(sigma_dyn^2 + err_j^2)) and the orbital RV through the oracle."""
    slots = np.arange(first_slot, first_slot + n, dtype=np.uint64)
    w = ph.philox_block(seed, slots >> np.uint64(1), ph.STREAM_SLOT, 0)
    odd = (slots & np.uint64(1)).astype(bool)
    coin_word = np.where(odd, w[:, 1], w[:, 0])
    thresh = int(round(p.fbin * 2 ** 24))
    is_binary = (coin_word >> np.uint32(8)) < thresh
    z0, z1 = ph.box_muller(w[:, 2], w[:, 3])
    j = (slots % np.uint64(S)).astype(np.int64)
    scale = np.sqrt(p.sigma_dyn ** 2 + np.asarray(err, dtype=np.float64)[j] ** 2)
    o = orbits_from_draws(p, seed, slots)
    E, rv, K = rv_from_orbits(o)
    return dict(is_binary=is_binary, noise=np.where(odd, z1, z0) * scale, rv=rv, K=K, orbits=o,
                value=np.where(odd, z1, z0) * scale + np.where(is_binary, rv, 0.0))


def biascorr_item(star, sample):
    return (np.uint64(star) << np.uint64(40)) + np.asarray(sample, dtype=np.uint64)


def biascorr_noise_from_draws(noise_seed, star, samples, n_epochs, err, stream=None):
    """(n_epochs, n_samples) measurement errors: one STREAM_EPOCH block per six epochs (noise6)."""
    stream = ph.STREAM_EPOCH if stream is None else stream
    items = biascorr_item(star, samples)
    out = np.zeros((n_epochs, len(items)))
    for g in range((n_epochs + 5) // 6):
        z6 = ph.noise6(ph.philox_block(noise_seed, items, stream, g))
        for k in range(6 * g, min(6 * g + 6, n_epochs)):
            out[k] = z6[k - 6 * g] * np.asarray(err, dtype=np.float64).reshape(-1)[k if np.size(err) > 1 else 0]
    return out
