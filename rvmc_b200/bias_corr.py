"""Drop-in for the reference's `BiasCorr` class (RVMC_bias_corr.py:22-476): multi-epoch
Delta-RV distributions and binary detection probabilities, computed on the GPU through the C ABI.

Two execution modes, chosen automatically:
  * FUSED (default): nothing is materialised.  `calculate_radial_velocities()` launches one kernel
    (rvmc_biascorr_fused) that samples each binary, evaluates its RV at every epoch, adds the
    measurement noise and writes only Delta-RV (fp32) and the detection flag.  The attribute arrays
    (`period`, `eccentricity`, ...) and `radial_velocities` are regenerated from the counter-based
    generator if and when they are read -- same (seed, star, sample) -> same numbers.
  * REPLAY: as soon as an attribute array has been read or assigned on the host (the user may have
    edited it, as scripts written for the reference do), or `mass_dist` / `period_dist` were given,
    the fp64 replay kernel (rvmc_biascorr_replay_f64) evaluates RVMC_bias_corr.py:406-422 on the
    arrays as they are.

The phase quirk of RVMC_bias_corr.py:357 (mean anomaly = ((2 pi t) mod P)/P) is
reproduced by default; `physical_phase=True` uses 2 pi frac(t/P) instead.  Additive kwargs:
`seed`, `device`, `physical_phase`.

Precision of what the FUSED mode hands back: `delta_rv`, `radial_velocities` and the detection flags come
from the fp32 kernel (per-orbit RVs within 1e-5 of the reference's fp64 arithmetic relative to
max(|RV|, K)); `delta_rv` is a float32 array, `radial_velocities` float64 arrays holding float32-accurate
values.  The REPLAY mode evaluates RVMC_bias_corr.py:406-422 in fp64 (1e-9) and returns float64.
`get_detected_binaries()` returns the reference's `bool (n_stars, n_samples)`; on the way from the GPU the
flags travel bit-packed (1/8 of the PCIe bytes) and are widened on the host (`flag_transfer`).
"""
import ctypes as C
import os
import warnings

import numpy as np

from . import _abi
from . import _device as dv
from ._trace import span
from .rvmc import _check_powers

_ORBIT = ("primary_mass", "period", "time", "mass_ratio", "eccentricity", "inclination", "orbit_rotation")


_FUSED_MAX_EPOCHS = 64          # rvmc_biascorr_fused: epochs per star (include/rvmc_b200.h)


class BiasCorr(dv.LazyArrays):
    """Determines the binary-detection bias of a multi-epoch RV campaign by simulating a population
    of binaries (GPU implementation of RVMC_bias_corr.py:22-476)."""

    _ARRAY_ATTRS = _ORBIT + ("semi_major_axis", "max_orbital_velocity", "delta_rv", "delta_rv_single")
    # star ranges a large fused pass is launched in (the D2H copy of one range overlaps the next ranges)
    _d2h_chunks = 12
    # how the detection flags of a fused pass travel to the host: "bits" = bit-packed by the kernel
    # (one __ballot_sync word per warp, 1/8 of the PCIe bytes) and widened to NumPy bool on the host by
    # the library's threaded helper; "bytes" = one uint8 per binary copied as is
    # "hybrid" = both at once: a share of the ranges as bytes by DMA while the CPU widens the packed rest;
    # "auto" (default) = "bytes" for a rank that has the host to itself (one copy engine at 56 GB/s hides
    # behind the kernel and costs no CPU), "bits" as soon as ranks share the host (LOCAL_WORLD_SIZE > 1: the
    # DMA route of the whole box saturates at 121 GB/s, the widening route at 205 GB/s)
    flag_transfer = os.environ.get("RVMC_FLAG_TRANSFER", "auto")
    flag_hybrid_byte_share = float(os.environ.get("RVMC_FLAG_BYTE_SHARE", "0.4"))

    def __init__(self, t_obs, masses=None, mass_errors=None, number_of_samples=10**4, min_mass=6,
                 max_mass=20, min_period=10**0.15, max_period=10**3.5, period_pi=-0.5, mass_ratio_kappa=0,
                 mass_power=-2.35, rv_errors=None, fbin=0.7, period_dist=None, mass_dist=None, min_q=0.1,
                 max_q=1.0, e_power=-0.5, seed=None, device=None, physical_phase=False):
        # RVMC_bias_corr.py:160-172
        self.current_star = 0
        self.min_q = min_q
        self.max_q = max_q
        self.kappa = mass_ratio_kappa
        self.pi = period_pi
        self.mass_power = mass_power
        self.number_of_samples = number_of_samples
        self.t_obs = t_obs
        self.number_of_stars = len(t_obs)
        self.e_power = e_power
        self.physical_phase = bool(physical_phase)
        self.seed = dv.fresh_seed() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF
        self._calls = 0
        self.nonconverged = 0
        self.counters = None
        nstars, ns = self.number_of_stars, int(number_of_samples)

        # ---- masses (RVMC_bias_corr.py:175-207) ------------------------------------------------
        self._mass_mode = 0
        self._masses = self._mass_errors = None
        given = 0
        given_arrays = {}
        lazy_mass_range = False
        if masses is not None:
            if len(masses) != nstars:
                raise ValueError(f"The number of masses ({len(masses)}) and epoch "
                                 f"sequences ({len(t_obs)}) are not equal!")
            self._masses = np.ascontiguousarray(np.reshape(masses, -1), dtype=np.float64)
            if mass_errors is not None:
                self._mass_mode = 1
                self._mass_errors = np.ascontiguousarray(np.reshape(mass_errors, -1), dtype=np.float64)
                if self._mass_errors.size != nstars:
                    raise ValueError("operands could not be broadcast together: mass_errors has %d elements, "
                                     "expected %d" % (self._mass_errors.size, nstars))
                lazy_mass_range = True           # min/max of the normal draws, computed when asked for
            else:
                self._mass_mode = 2
                self.min_mass = np.min(self._masses) / 2.e30
                self.max_mass = np.max(self._masses) / 2.e30
        elif mass_dist is not None:
            if np.shape(mass_dist) == (len(t_obs), number_of_samples):
                self.min_mass = np.min(mass_dist) / 2.e30
                self.max_mass = np.max(mass_dist) / 2.e30
                given |= 1
                given_arrays["primary_mass"] = mass_dist
            else:
                raise ValueError(f"The shape of mass distributions ({np.shape(mass_dist)})"
                                 f" not as expected {(len(t_obs), number_of_samples)}.")
        else:
            self.min_mass = min_mass
            self.max_mass = max_mass
            _check_powers(mass_power=mass_power)
        self._imf_range = (float(min_mass), float(max_mass))      # what the IMF draw at :207 used

        # ---- periods (:209-223) -----------------------------------------------------------------
        if period_dist is not None:
            if np.shape(period_dist) == (len(t_obs), number_of_samples):
                # the reference overwrites min_mass/max_mass here; kept
                self.min_mass = np.min(period_dist) / (3600 * 24)
                self.max_mass = np.max(period_dist) / (3600 * 24)
                lazy_mass_range = False
                given |= 2
                given_arrays["period"] = period_dist
            else:
                raise ValueError(f"The shape of period distribution ({np.shape(period_dist)})"
                                 f" not as expected: {(len(t_obs), number_of_samples)}.")
        else:
            self.min_period = min_period
            self.max_period = max_period
            _check_powers(period_pi=period_pi)

        # ---- RV errors (:225-235) -----------------------------------------------------------------
        if rv_errors is None:
            self.rv_errors = None
        else:
            if isinstance(rv_errors, (float, int)):
                self.rv_errors = rv_errors
            elif isinstance(rv_errors, (list, np.ndarray)):
                if len(rv_errors) == nstars:
                    self.rv_errors = rv_errors
                else:
                    raise ValueError(f"The number of radial velocity uncertainties ({len(rv_errors)}) does not match"
                                     f"the number of stars ({self.number_of_stars}).")
            # any other type: the reference silently leaves the attribute unset
        _check_powers(mass_ratio_kappa=mass_ratio_kappa, e_power=e_power)

        self._lazy_init(dv.resolve_device(device))
        self._given = given
        self._lazy_mass_range = lazy_mass_range
        for k, v in given_arrays.items():
            setattr(self, k, v)                       # host-authoritative -> replay mode
        self._generated = set()                       # orbit attributes regenerated from the seed so far
        self._seed0 = self._next_seed()
        self._rv_seed = None
        self._rv_cache = None                         # device results of the last RV pass
        self._rv_user = None                          # user-assigned radial_velocities
        self.fbin = fbin
        self.delta_rv_single = None
        # delta_rv = np.zeros(shape) at :251 -- materialised only if it is read before the first RV pass

    # ---- plumbing -------------------------------------------------------------------------------
    def _next_seed(self):
        s = dv.call_seed(self.seed, self._calls)
        self._calls += 1
        return s

    def __getattr__(self, name):
        d = self.__dict__
        if "_host" in d:
            if name in _ORBIT and name not in d["_host"] and name not in d["_dev"]:
                self._materialize((name,))
            elif name in ("semi_major_axis", "max_orbital_velocity") and not self._has(name):
                if d.get("_rv_cache") is None:
                    return None                                           # RVMC_bias_corr.py:246-247
                # calculate_radial_velocities() leaves both behind in the reference (:412-413); the
                # fused pass never needs them, so they are computed when first asked for
                self.calculate_semi_major_axis()
                self.calculate_max_orbital_velocity()
            elif name in ("min_mass", "max_mass") and d.get("_lazy_mass_range"):
                pm = self._orbit_on_device(("primary_mass",))["primary_mass"]
                object.__setattr__(self, "min_mass", float(pm.min().item()) / 2.e30)      # :186-187
                object.__setattr__(self, "max_mass", float(pm.max().item()) / 2.e30)
                object.__setattr__(self, "_lazy_mass_range", False)
                return d[name]
            elif name == "radial_velocities":
                return self._radial_velocities_list()
            elif name == "delta_rv" and not self._has(name):
                self._set_device("delta_rv", dv.zeros((self.number_of_stars, int(self.number_of_samples)), "float32",
                                                      self._device))
        return dv.LazyArrays.__getattr__(self, name)

    def __setattr__(self, name, value):
        if name == "radial_velocities" and "_host" in self.__dict__:
            object.__setattr__(self, "_rv_user", value)
            return
        dv.LazyArrays.__setattr__(self, name, value)

    def _pop(self):
        has_p = "min_period" in self.__dict__
        kw = dict(mass_power=float(self.mass_power), period_pi=float(self.pi), min_q=float(self.min_q),
                  max_q=float(self.max_q), mass_ratio_kappa=float(self.kappa), e_power=float(self.e_power),
                  fbin=float(self.fbin) if "fbin" in self.__dict__ else 0.7)
        if self._mass_mode == 0 and not (self._given & 1):
            kw.update(min_mass=self._imf_range[0], max_mass=self._imf_range[1])
        if has_p:
            kw.update(min_period=float(self.min_period), max_period=float(self.max_period))
        return _abi.PopParams.defaults(**kw)

    def _tables(self):
        """Device tables of the observing campaign: n_epochs[nstars], t_obs / rv_err [nstars][max_epochs]."""
        nstars = self.number_of_stars
        tob = None
        if nstars:
            try:                                  # a regular campaign (same number of epochs for every star)
                tob = np.asarray(self.t_obs, dtype=np.float64)
            except ValueError:                    # ragged
                tob = None
            if tob is not None and (tob.ndim != 2 or tob.shape[0] != nstars):
                tob = None
        if tob is not None:
            ragged = False
            nep = np.full(nstars, tob.shape[1], dtype=np.int32)
        else:
            rows = [np.ravel(t) for t in self.t_obs]
            nep = np.fromiter((r.size for r in rows), dtype=np.int32, count=nstars)
            ragged = True
        if nstars and nep.min() < 1:
            raise ValueError("every star needs at least one observing epoch")
        max_ep = int(nep.max()) if nstars else 1
        if ragged:
            tob = np.zeros((nstars, max_ep), dtype=np.float64)
            for i, r in enumerate(rows):
                tob[i, :nep[i]] = r
        err = None
        rv_errors = self.__dict__.get("rv_errors", None)
        if rv_errors is not None and not isinstance(rv_errors, (np.ndarray, list)):
            # one scalar for every star and epoch (check_error_shape's last branch, :391-392)
            err = np.full((nstars, max_ep), rv_errors, dtype=np.float32)
            if ragged:
                err[np.arange(max_ep)[None, :] >= nep[:, None]] = 0.0
        elif rv_errors is not None:
            err = np.zeros((nstars, max_ep), dtype=np.float32)
            for i in range(nstars):
                self.current_star = i
                e = self.check_error_shape(int(nep[i]))
                err[i, :nep[i]] = np.ravel(e) if isinstance(e, np.ndarray) else e
        # one host-to-device copy for the three tables (fp64 first: every view stays aligned)
        nb = (tob.nbytes, nep.nbytes, 0 if err is None else err.nbytes)
        packed = np.empty(sum(nb), dtype=np.uint8)
        packed[:nb[0]] = tob.reshape(-1).view(np.uint8)
        packed[nb[0]:nb[0] + nb[1]] = nep.view(np.uint8)
        if err is not None:
            packed[nb[0] + nb[1]:] = err.reshape(-1).view(np.uint8)
        t = dv.torch()
        d = t.from_numpy(packed).to(self._device)
        d_tobs = d[:nb[0]].view(t.float64).view(nstars, max_ep)
        d_nep = d[nb[0]:nb[0] + nb[1]].view(t.int32)
        d_err = None if err is None else d[nb[0] + nb[1]:].view(t.float32).view(nstars, max_ep)
        return nep, max_ep, d_nep, d_tobs, d_err

    def _mass_tables(self):
        dev = self._device
        m = None if self._masses is None else dv.to_device(self._masses, "float64", dev)
        me = None if self._mass_errors is None else dv.to_device(self._mass_errors, "float64", dev)
        return m, me

    def _generate(self, names):
        """Regenerates orbit attributes from (seed, star, sample) on the device -> {name: tensor}."""
        dev = self._device
        nstars, ns = self.number_of_stars, int(self.number_of_samples)
        tens = {k: dv.empty((nstars, ns), "float64", dev) for k in names}
        given = 0
        if "period" in self._host or "period" in self._dev:
            # time and eccentricity follow the CURRENT periods (RVMC_bias_corr.py:237,240)
            tens["period"] = self._on_device("period")
            given |= 2
        m, me = self._mass_tables()
        mode = self._mass_mode
        out_names = dict(tens)
        if mode == 2 and "primary_mass" in names:
            # masses at face value have shape (nstars, 1) in the reference (:190)
            del out_names["primary_mass"]
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_biascorr_sample(C.byref(self._pop()), self._seed0, nstars, mode, dv.ptr(m),
                                                       dv.ptr(me), 0, ns, given, dv.soa(**out_names),
                                                       dv.stream_ptr(dev)))
        if mode == 2 and "primary_mass" in names:
            tens["primary_mass"] = dv.to_device(self._masses.reshape(-1, 1), "float64", dev)
        return {k: tens[k] for k in names}

    def _materialize(self, names):
        for k, v in self._generate(tuple(names)).items():
            self._set_device(k, v)
            self._generated.add(k)

    def _orbit_on_device(self, names=_ORBIT):
        missing = [k for k in names if not self._has(k)]
        fresh = self._generate(tuple(missing)) if missing else {}
        return {k: (fresh[k] if k in fresh else self._on_device(k)) for k in names}

    def _replay_mode(self):
        """True once any orbit attribute is host-authoritative (read, edited or supplied)."""
        return any(k in self._host for k in _ORBIT)

    # ---- samplers: fresh draws, as each call of the reference's methods gives (:256-328) ----------
    def _draw(self, field, size=None, given=0):
        nstars, ns = self.number_of_stars, int(self.number_of_samples)
        dev = self._device
        shape = (nstars, ns) if size is None else (tuple(size) if not np.isscalar(size) else (int(size),))
        n = int(np.prod(shape))
        out = dv.empty((n,), "float64", dev)
        tens = {field: out}
        pop = self._pop()
        if given & 2:
            tens["period"] = self._on_device("period") if self._has("period") else \
                self._orbit_on_device(("period",))["period"]
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_sample_orbits(C.byref(pop), self._next_seed(), 0, n, given, dv.soa(**tens),
                                                     None, dv.stream_ptr(dev)))
        return dv.to_host(out).reshape(shape)

    def sample_powerlaw(self, min_val, max_val, power, size=None):
        _check_powers(power=power)
        nstars, ns = self.number_of_stars, int(self.number_of_samples)
        shape = (nstars, ns) if size is None else (tuple(size) if not np.isscalar(size) else (int(size),))
        n = int(np.prod(shape))
        dev = self._device
        out = dv.empty((n,), "float64", dev)
        pop = _abi.PopParams.defaults(min_q=float(min_val), max_q=float(max_val), mass_ratio_kappa=float(power))
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_sample_orbits(C.byref(pop), self._next_seed(), 0, n, 0,
                                                     dv.soa(mass_ratio=out), None, dv.stream_ptr(dev)))
        return dv.to_host(out).reshape(shape)

    def sample_masses(self):
        return self.sample_powerlaw(self.min_mass, self.max_mass, self.mass_power) * 2e30

    def sample_inclination(self):
        return self._draw("inclination")

    def sample_period(self):
        _check_powers(period_pi=self.pi)
        return self._draw("period")

    def sample_period_old(self):
        """RVMC_bias_corr.py:290-293 (unused by the reference itself): uniform in sqrt(log10 P)."""
        lo, hi = np.log10(self.min_period) ** 0.5, np.log10(self.max_period) ** 0.5
        u = self.sample_powerlaw(lo, hi, 0.0)        # index 0 == uniform between lo and hi
        return 10 ** (u ** 2) * 24 * 3600

    def sample_mass_ratio(self):
        _check_powers(mass_ratio_kappa=self.kappa)
        return self._draw("mass_ratio")

    def sample_orbit_rotation(self):
        return self._draw("orbit_rotation")

    def sample_eccentricity(self):
        _check_powers(e_power=self.e_power)
        return self._draw("eccentricity", given=2)

    def sample_time(self):
        return self._draw("time", given=2)

    # ---- per-star building blocks (:330-372) ------------------------------------------------------
    def calculate_semi_major_axis(self):
        t = self._orbit_on_device(("mass_ratio", "primary_mass", "period"))
        n = t["period"].numel()
        pm = t["primary_mass"].expand_as(t["period"]).contiguous()
        out = dv.empty(tuple(t["period"].shape), "float64", self._device)
        with dv.DeviceGuard(self._device):
            _abi.check(_abi.lib().rvmc_orbit_scales_f64(n, dv.soa(mass_ratio=t["mass_ratio"], primary_mass=pm,
                                                                  period=t["period"]), None, dv.ptr(out), None,
                                                        dv.stream_ptr(self._device)))
        self._set_device("semi_major_axis", out)

    def calculate_max_orbital_velocity(self):
        if not self._has("semi_major_axis"):
            raise TypeError("unsupported operand type(s) for *: 'float' and 'NoneType'")     # :345 with a = None
        t = self._orbit_on_device(("mass_ratio", "primary_mass", "period", "eccentricity"))
        n = t["period"].numel()
        pm = t["primary_mass"].expand_as(t["period"]).contiguous()
        a = self._on_device("semi_major_axis")
        out = dv.empty(tuple(t["period"].shape), "float64", self._device)
        with dv.DeviceGuard(self._device):
            _abi.check(_abi.lib().rvmc_orbit_scales_f64(
                n, dv.soa(mass_ratio=t["mass_ratio"], primary_mass=pm, period=t["period"],
                          eccentricity=t["eccentricity"]), dv.ptr(a), None, dv.ptr(out), dv.stream_ptr(self._device)))
        self._set_device("max_orbital_velocity", out)

    def calculate_eccentric_anomaly(self, time_deltas):
        """Eccentric anomalies of the current star at the given times (:348-358); `time_deltas` has
        shape (n_epochs, n_samples)."""
        i = self.current_star
        dev = self._device
        t = self._orbit_on_device(("period", "eccentricity"))
        td = dv.to_device(np.asarray(time_deltas, dtype=np.float64), "float64", dev)
        nep, ns = td.shape
        out = dv.empty((nep, ns), "float64", dev)
        nonconv = dv.zeros((1,), "int32", dev)
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_kepler_epochs_f64(nep, ns, dv.ptr(td), dv.ptr(t["period"][i].contiguous()),
                                                         dv.ptr(t["eccentricity"][i].contiguous()),
                                                         0 if self.physical_phase else 1, dv.ptr(out),
                                                         dv.ptr(nonconv), dv.stream_ptr(dev)))
        bad = int(nonconv.item())
        self.nonconverged += bad
        if bad:
            warnings.warn("some failed to converge after 50 iterations (%d orbits)" % bad, RuntimeWarning)
        return dv.to_host(out)

    def binary_radial_velocity(self, eccentric_anomalies_i):
        """RVs of the current star for the given eccentric anomalies (:360-372)."""
        i = self.current_star
        dev = self._device
        if not self._has("max_orbital_velocity"):
            raise TypeError("'NoneType' object is not subscriptable")        # :369 with max_orbital_velocity None
        t = self._orbit_on_device(("eccentricity", "inclination", "orbit_rotation"))
        vm = self._on_device("max_orbital_velocity")
        E = dv.to_device(np.asarray(eccentric_anomalies_i, dtype=np.float64), "float64", dev)
        nep, ns = E.shape
        out = dv.empty((nep, ns), "float64", dev)
        with dv.DeviceGuard(dev):
            _abi.check(_abi.lib().rvmc_rv_epochs_f64(nep, ns, dv.ptr(E), dv.ptr(vm[i].contiguous()),
                                                     dv.ptr(t["eccentricity"][i].contiguous()),
                                                     dv.ptr(t["inclination"][i].contiguous()),
                                                     dv.ptr(t["orbit_rotation"][i].contiguous()), dv.ptr(out),
                                                     dv.stream_ptr(dev)))
        return dv.to_host(out)

    def check_error_shape(self, number_of_epochs):
        """The error(s) that apply to the current star (RVMC_bias_corr.py:374-393): the one number given
        for everything, this star's number, or this star's per-epoch values as a column (n_epochs, 1)."""
        sequence = (np.ndarray, list)
        errors = self.rv_errors
        if not isinstance(errors, sequence):
            return errors
        mine = errors[self.current_star]
        if not isinstance(mine, sequence):
            return mine
        if len(mine) != number_of_epochs:
            raise ValueError(f"The number of radial velocity uncertainties "
                             f"({len(mine)}) does not match the number of epochs,"
                             f"({number_of_epochs})")
        return np.reshape(mine, (-1, 1))

    def generate_rv_errors(self, number_of_epochs, number_of_samples):
        """Fresh measurement errors for the current star, shape (n_epochs, n_samples) (:395-404)."""
        use_err = self.check_error_shape(number_of_epochs)
        err = np.broadcast_to(np.asarray(use_err, dtype=np.float64).reshape(-1), (number_of_epochs,))
        dev = self._device
        out = dv.empty((number_of_epochs, number_of_samples), "float64", dev)
        with dv.DeviceGuard(dev):
            lib = _abi.lib()
            for k in range(number_of_epochs):
                _abi.check(lib.rvmc_normal_f64(self._next_seed(), 0, int(number_of_samples), 7, k, 0.0, float(err[k]),
                                               0, dv.ptr(out[k]), dv.stream_ptr(dev)))
        return dv.to_host(out)

    # ---- the hot path (:406-476) ---------------------------------------------------------------------
    def calculate_radial_velocities(self, delta_RV=20, n_sigma_RV=4):
        """Radial velocities at every epoch for every star and sample; updates `delta_rv`
        (RVMC_bias_corr.py:406-422).  The detection flags for (delta_RV, n_sigma_RV) are computed in
        the same pass and cached for `get_detected_binaries`."""
        dev = self._device
        nstars, ns = self.number_of_stars, int(self.number_of_samples)
        with span("biascorr.tables"):
            nep, max_ep, d_nep, d_tobs, d_err = self._tables()
        self._rv_seed = self._next_seed()
        self._rv_user = None
        quirk = 0 if self.physical_phase else 1
        if self._replay_mode() or max_ep > _FUSED_MAX_EPOCHS:
            # host-authoritative attributes, or more epochs per star than the fused kernel keeps in
            # registers: the materialised fp64 path (any number of epochs, RVMC_bias_corr.py:406-422 as written)
            t = self._orbit_on_device()
            pm = t["primary_mass"]
            rv = dv.zeros((nstars, max_ep, ns), "float64", dev)
            drv = dv.empty((nstars, ns), "float64", dev)
            with dv.DeviceGuard(dev):
                _abi.check(_abi.lib().rvmc_biascorr_replay_f64(
                    nstars, ns, max_ep, dv.ptr(d_nep), dv.ptr(d_tobs),
                    dv.soa(**{k: t[k].contiguous() for k in _ORBIT}), int(pm.shape[1]) if pm.dim() == 2 else ns,
                    dv.ptr(d_err), self._rv_seed, 0, quirk, dv.ptr(rv), dv.ptr(drv), dv.stream_ptr(dev)))
            self._rv_cache = dict(mode="replay", rv=rv, nep=nep, max_ep=max_ep, d_nep=d_nep, d_err=d_err, det={})
            self._set_device("delta_rv", drv)
            # the reference also leaves these two behind (:412-413)
            self.calculate_semi_major_axis()
            self.calculate_max_orbital_velocity()
            return
        drv = dv.empty((nstars, ns), "float32", dev)
        det = bits = None
        if d_err is not None:
            if self._flag_mode() in ("bits", "hybrid"):
                bits = dv.empty((nstars, (ns + 31) // 32), "int32", dev)
            if self._flag_mode() in ("bytes", "hybrid"):
                det = dv.empty((nstars, ns), "uint8", dev)
        counters = dv.zeros((4,), "int64", dev)
        per_star = dv.zeros((max(nstars, 1),), "int32", dev)
        # large problems go out as a few launches over star ranges (sample ranges when there are only
        # a few stars): the copy of one range's results to the host (get_detected_binaries) then overlaps
        # the compute of the next ranges.  The ranges are small at both ends (_range_weights); everything
        # launch-invariant is prepared once.
        ranges = [(0, nstars, 0, ns)]
        if nstars * ns >= 4_000_000 and self._d2h_chunks > 1:
            by_star = nstars >= self._d2h_chunks // 2
            total = nstars if by_star else ns
            n_chunks = min(total, self._d2h_chunks)
            w = self._range_weights(n_chunks)
            bounds = np.concatenate([[0], np.round(np.cumsum(w) / w.sum() * total)]).astype(np.int64)
            if not by_star:
                bounds[1:-1] = (bounds[1:-1] + 2047) // 4096 * 4096        # whole blocks of samples
                bounds = np.unique(np.clip(bounds, 0, ns))
            ranges = [(int(a), int(b), 0, ns) if by_star else (0, nstars, int(a), int(b))
                      for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
        chunks = []
        t = dv.torch()
        launch = self._fused_launcher(max_ep, d_nep, d_tobs, d_err, delta_RV, n_sigma_RV, drv, det, None, counters,
                                      per_star, bits=bits)
        # the ranges are independent: they alternate between the current stream and a side stream, so the
        # draining tail of one launch is filled by the first blocks of the next
        main = t.cuda.current_stream(dev)
        streams = [main, dv.side_stream(dev)] if len(ranges) > 1 else [main]
        if len(streams) > 1:
            ready = t.cuda.Event()
            ready.record(main)
            streams[1].wait_event(ready)
        with dv.DeviceGuard(dev), span("biascorr.launch_ranges"):
            for k, (s0, s1, c0, c1) in enumerate(ranges):
                st = streams[k % len(streams)]
                launch(s0, s1, C.c_void_p(st.cuda_stream), c0, c1)
                ev = t.cuda.Event()
                ev.record(st)
                chunks.append((s0, s1, c0, c1, ev))
        if len(streams) > 1:
            done = t.cuda.Event()
            done.record(streams[1])
            main.wait_event(done)        # later work on the current stream sees every range
        self._rv_cache = dict(mode="fused", nep=nep, max_ep=max_ep, d_nep=d_nep, d_tobs=d_tobs, d_err=d_err,
                              det={(float(delta_RV), float(n_sigma_RV)): (det, bits, chunks)} if d_err is not None else {},
                              counters=counters, per_star=per_star, pop=launch.pop, err_fp=self._errors_fingerprint())
        self._set_device("delta_rv", drv)

    @staticmethod
    def _range_weights(n):
        """Relative sizes of the ranges a large pass is launched in: small at BOTH ends.  The host side of a
        range (copy + widening of its flags) starts when its kernel ends and the ranges' host work is
        serial, so the pass takes  kernel(first range) + max(host work, remaining kernels) + host(last
        range): the first range delays the start of the host pipeline (it dominates when the host is the
        slower side -- eight ranks sharing one host's memory system), the last one is the tail nothing
        overlaps (it dominates when the GPU is the slower side -- one rank)."""
        k = np.arange(n, dtype=np.float64)
        return np.minimum(1.6 ** k, 1.6 ** (n - 1 - k))

    def _fused_launcher(self, max_ep, d_nep, d_tobs, d_err, delta_RV, n_sigma_RV, drv, det, rv_out, counters,
                        per_star, bits=None, pop=None):
        """-> launch(s0, s1, stream, c0, c1): one rvmc_biascorr_fused call over stars [s0, s1) and samples
        [c0, c1) (default: all).  Population, mass tables
        and base pointers are resolved once; a star range is pointer arithmetic on the row-major
        [nstars, ...] buffers (the caller holds the device guard)."""
        dev = self._device
        ns = int(self.number_of_samples)
        m, me = self._mass_tables()
        pop = self._pop() if pop is None else pop       # relaunches of a cached pass reuse ITS population
        lib = _abi.lib()
        quirk = 0 if self.physical_phase else 1
        stream = dv.stream_ptr(dev)
        keep = (d_nep, d_tobs, d_err, m, me, drv, det, bits, rv_out, counters, per_star, pop)

        def base(x):
            return (0, 0) if x is None else (x.data_ptr(), x.stride(0) * x.element_size() if x.dim() else 0)
        b_nep, b_tobs, b_err, b_m, b_me = base(d_nep), base(d_tobs), base(d_err), base(m), base(me)
        b_drv, b_det, b_rv, b_star = base(drv), base(det), base(rv_out), base(per_star)
        b_bits = base(bits)
        bits_stride = 0 if bits is None else bits.stride(0)
        p_cnt = dv.ptr(counters)
        args = (float(delta_RV), float(n_sigma_RV))

        def rows(b, s0):
            return (b[0] + s0 * b[1]) if b[0] else None

        def cols(b, s0, c0, itemsize):
            return (b[0] + s0 * b[1] + c0 * itemsize) if b[0] else None

        def launch(s0, s1, stream=stream, c0=0, c1=ns, _keep=keep):
            # the sample index is the Philox counter, so a sample range reproduces the full pass bit for bit
            # (a sample range starts at a multiple of 32, so its flags start on a word of the packed rows)
            _abi.check(lib.rvmc_biascorr_fused_packed(
                C.byref(pop), self._seed0, self._rv_seed, s0, s1 - s0, max_ep, rows(b_nep, s0), rows(b_tobs, s0),
                rows(b_err, s0), self._mass_mode, rows(b_m, s0), rows(b_me, s0), quirk, args[0], args[1], c0, c1 - c0,
                ns, cols(b_drv, s0, c0, 4), cols(b_det, s0, c0, 1), cols(b_bits, s0, c0 // 32, 4), bits_stride,
                cols(b_rv, s0, c0, 4), p_cnt, rows(b_star, s0), stream))
        launch.pop = pop
        return launch

    def _launch_fused(self, nep, max_ep, d_nep, d_tobs, d_err, delta_RV, n_sigma_RV, drv, det, rv_out, counters,
                      per_star, stars=None, bits=None, pop=None):
        s0, s1 = (0, self.number_of_stars) if stars is None else stars
        with dv.DeviceGuard(self._device):
            self._fused_launcher(max_ep, d_nep, d_tobs, d_err, delta_RV, n_sigma_RV, drv, det, rv_out, counters,
                                 per_star, bits=bits, pop=pop)(int(s0), int(s1))

    def _flag_mode(self):
        mode = self.flag_transfer
        if mode == "auto":
            mode = "bits" if int(os.environ.get("LOCAL_WORLD_SIZE", "1")) > 1 else "bytes"
        if mode not in ("bits", "bytes", "hybrid"):
            raise ValueError("flag_transfer must be 'auto', 'bits', 'bytes' or 'hybrid' (got %r)" % (mode,))
        return mode

    def _errors_fingerprint(self):
        """The measurement errors AS THEY ARE NOW.  get_detected_binaries() compares it with the one taken
        by calculate_radial_velocities(): the reference reads `self.rv_errors` again at :459-466, so a user
        who edits the errors between the two calls gets thresholds from the new errors."""
        e = self.__dict__.get("rv_errors", None)
        if e is None:
            return b""
        try:
            return np.asarray(e, dtype=np.float64).tobytes()
        except (ValueError, TypeError):          # ragged per-epoch lists
            return repr([list(np.ravel(x)) for x in e]).encode()

    def _cached_rv_device(self, c):
        """float32 [nstars, max_ep, ns] noisy RVs of the cached fused pass, regenerated from its own seeds,
        population and error table (what the reference keeps in `radial_velocities`, :419-421)."""
        nstars, ns = self.number_of_stars, int(self.number_of_samples)
        rvd = dv.zeros((nstars, c["max_ep"], ns), "float32", self._device)
        self._launch_fused(c["nep"], c["max_ep"], c["d_nep"], c["d_tobs"], c["d_err"], 20, 4, None, None, rvd,
                           None, None, pop=c["pop"])
        return rvd

    def _radial_velocities_list(self):
        """`radial_velocities`: list of (n_epochs_i, n_samples) arrays (RVMC_bias_corr.py:250,419-421)."""
        if self._rv_user is not None:
            return self._rv_user
        c = self._rv_cache
        if c is None:
            return [[]] * self.number_of_stars
        if c["mode"] == "replay":
            rv = dv.to_host(c["rv"])
        else:
            if "rv_host" not in c:
                c["rv_host"] = dv.to_host(self._cached_rv_device(c)).astype(np.float64)
            rv = c["rv_host"]
        out = [rv[i, :c["nep"][i]] for i in range(self.number_of_stars)]
        self._rv_user = out          # host copy is authoritative from now on (the user may edit it)
        return out

    def generate_single_delta_rv(self):
        """Delta-RV of single stars from measurement noise alone (RVMC_bias_corr.py:424-442)."""
        needed_singles = int((self.number_of_samples - self.fbin * self.number_of_samples) / self.fbin)
        nstars = self.number_of_stars
        dev = self._device
        if self.rv_errors is not None:
            nep, max_ep, d_nep, _, d_err = self._tables()
            out = dv.empty((nstars, needed_singles), "float64", dev)
            with dv.DeviceGuard(dev):
                _abi.check(_abi.lib().rvmc_biascorr_singles(self._next_seed(), nstars, max_ep, dv.ptr(d_nep),
                                                            dv.ptr(d_err), 0, needed_singles, dv.ptr(out),
                                                            dv.stream_ptr(dev)))
            self._set_device("delta_rv_single", out)
        else:
            self.delta_rv_single = np.zeros((nstars, needed_singles))
            print("Cannot generate single star radial velocities as there are no uncertainties to work with."
                  "So all are zero. ")

    def get_all_delta_rv(self):
        """RVMC_bias_corr.py:444-449."""
        return np.concatenate([np.asarray(self.delta_rv, dtype=np.float64), self.delta_rv_single], axis=1).ravel()

    def get_detected_binaries(self, delta_RV=20, n_sigma_RV=4):
        """Bool array (n_stars, n_samples): is each simulated observation set detected as a binary
        (RVMC_bias_corr.py:451-476)."""
        nstars, ns = self.number_of_stars, int(self.number_of_samples)
        c = self._rv_cache
        if c is None and self._rv_user is None:
            return np.zeros((nstars, ns), dtype=bool)           # radial_velocities is still [[]]*n: no pairs
        if self.rv_errors is None:
            raise TypeError("'NoneType' object is not subscriptable")      # err[v1] at :472
        dev = self._device
        key = (float(delta_RV), float(n_sigma_RV))
        if self._rv_user is not None or c["mode"] == "replay":
            if self._rv_user is not None:
                nep, max_ep, d_nep, _, d_err = self._tables()
                rvh = np.zeros((nstars, max_ep, ns), dtype=np.float64)
                for i, r in enumerate(self._rv_user):
                    r = np.asarray(r, dtype=np.float64)
                    if r.size:
                        rvh[i, :r.shape[0]] = r
                        nep[i] = r.shape[0]
                    else:
                        nep[i] = 0
                rv = dv.to_device(rvh, "float64", dev)
                d_nep = dv.to_device(nep, "int32", dev)
            else:
                rv, max_ep, d_nep, d_err = c["rv"], c["max_ep"], c["d_nep"], c["d_err"]
            det = dv.empty((nstars, ns), "uint8", dev)
            with dv.DeviceGuard(dev):
                _abi.check(_abi.lib().rvmc_biascorr_detect(nstars, ns, max_ep, dv.ptr(d_nep), dv.ptr(rv), dv.ptr(d_err),
                                                           float(delta_RV), float(n_sigma_RV), dv.ptr(det),
                                                           dv.stream_ptr(dev)))
            return dv.to_host(det).astype(bool)
        if c["err_fp"] != self._errors_fingerprint():
            # rv_errors was edited after the pass: the reference applies the CURRENT errors to the RVs it
            # stored (RVMC_bias_corr.py:459-466).  Same here: the RVs of the cached pass (its own noise) are
            # regenerated and tested against thresholds from the current errors.  Rare path, fp64 test kernel.
            _, max_ep2, _, _, d_err2 = self._tables()
            if max_ep2 != c["max_ep"]:
                raise ValueError("t_obs changed since calculate_radial_velocities()")
            rv = self._cached_rv_device(c).double()
            det = dv.empty((nstars, ns), "uint8", dev)
            with dv.DeviceGuard(dev):
                _abi.check(_abi.lib().rvmc_biascorr_detect(nstars, ns, max_ep2, dv.ptr(c["d_nep"]), dv.ptr(rv),
                                                           dv.ptr(d_err2), float(delta_RV), float(n_sigma_RV),
                                                           dv.ptr(det), dv.stream_ptr(dev)))
            return dv.to_host(det).view(np.bool_)
        if key not in c["det"]:
            det = bits = None
            if self._flag_mode() in ("bits", "hybrid"):
                bits = dv.empty((nstars, (ns + 31) // 32), "int32", dev)
            if self._flag_mode() in ("bytes", "hybrid"):
                det = dv.empty((nstars, ns), "uint8", dev)
            self._launch_fused(c["nep"], c["max_ep"], c["d_nep"], c["d_tobs"], c["d_err"], delta_RV, n_sigma_RV, None,
                               det, None, None, None, bits=bits, pop=c["pop"])
            c["det"][key] = (det, bits, None)
        det, bits, chunks = c["det"][key]
        if chunks is not None and len(chunks) < 2:
            chunks = None
        if bits is not None and det is not None:
            return dv.fetch_flags_hybrid(det, bits, chunks, self.flag_hybrid_byte_share)
        if bits is not None:
            with span("biascorr.fetch_flags"):
                return dv.unpack_bits_overlapped(bits, ns, chunks)
        if chunks:
            return dv.to_host_overlapped(det, chunks).view(np.bool_)
        return dv.to_host(det).view(np.bool_)

    def detection_counts(self):
        """Additive helper: per-star detected counts and the global counters of the last fused pass
        (binary-epoch RVs, detected, fp64-guard lanes, non-converged lanes) without copying the
        per-binary arrays to the host."""
        c = self._rv_cache
        if c is None or c["mode"] != "fused":
            raise RuntimeError("no fused radial-velocity pass has been run")
        cnt = dv.to_host(c["counters"])
        self.counters = dict(binary_epoch_rvs=int(cnt[0]), detected=int(cnt[1]), guard=int(cnt[2]),
                             nonconverged=int(cnt[3]))
        return dv.to_host(c["per_star"])[:self.number_of_stars].astype(np.int64), self.counters
