"""ctypes mirror of include/rvmc_b200.h and loader of librvmc_b200.so.

There is NO CPU fallback: `lib()` raises if the shared library is missing, and the library's
entry points return RVMC_ERR_NO_DEVICE when no sm_100 GPU is present.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RVMC_B200_LIB lets a developer point at an experimental build of the same ABI (tuning sweeps)
LIB_PATH = os.environ.get("RVMC_B200_LIB") or os.path.join(HERE, "librvmc_b200.so")

RVMC_OK = 0
RVMC_ERR_INVALID = -1
RVMC_ERR_ZERODIV = -2
RVMC_ERR_UNSUPPORTED = -3
RVMC_ERR_NO_DEVICE = -4


class PopParams(C.Structure):
    """rvmc_pop_params (include/rvmc_b200.h)."""
    _fields_ = [(n, C.c_double) for n in (
        "min_mass", "max_mass", "mass_power", "min_period", "max_period", "period_pi", "min_q",
        "max_q", "mass_ratio_kappa", "e_power", "e_min", "e_max_mid", "e_max_long", "p_circ_days",
        "p_mid_days", "sigma_dyn", "fbin", "guard_amp")] + [("kepler_iters", C.c_int32),
                                                             ("substream", C.c_uint32)]

    @classmethod
    def defaults(cls, **overrides):
        """RVMC.py:28-43 defaults + the constants hard-coded at RVMC.py:155-159."""
        p = cls(min_mass=6.0, max_mass=20.0, mass_power=-2.35, min_period=10. ** 0.15,
                max_period=10. ** 3.5, period_pi=-0.5, min_q=0.1, max_q=1.0, mass_ratio_kappa=0.0,
                e_power=-0.5, e_min=1e-5, e_max_mid=0.5, e_max_long=0.9, p_circ_days=4.0,
                p_mid_days=6.0, sigma_dyn=2.0, fbin=0.7, guard_amp=64.0, kepler_iters=1, substream=0)
        for k, v in overrides.items():
            if not hasattr(p, k):
                raise TypeError("unknown population parameter %r" % k)
            setattr(p, k, v)
        return p

    def copy(self, **overrides):
        q = PopParams.from_buffer_copy(self)
        for k, v in overrides.items():
            setattr(q, k, v)
        return q


_PD = C.POINTER(C.c_double)
_PF = C.POINTER(C.c_float)


class OrbitSoA(C.Structure):
    """rvmc_orbit_soa: device pointers as integers (0 = NULL)."""
    _fields_ = [(n, C.c_void_p) for n in (
        "primary_mass", "period", "time", "mass_ratio", "eccentricity", "inclination",
        "orbit_rotation", "eccentric_anomaly")]


class TraceSoA(C.Structure):
    """rvmc_trace_soa."""
    _fields_ = [(n, C.c_void_p) for n in (
        "words", "log2m", "x", "uphase", "q", "e", "cosi", "wturn", "K", "rv", "noise", "is_binary")]


class PointStats(C.Structure):
    """rvmc_point_stats."""
    _fields_ = [(n, C.c_double) for n in ("mode", "mean", "median", "p05", "p16", "p84", "p95",
                                          "sigma_max")] + \
               [(n, C.c_uint32) for n in ("n_guard", "n_nonconverged", "n_hist_bins", "reserved")]


class FlagBlock(C.Structure):
    """rvmc_flag_block."""
    _fields_ = [("row0", C.c_int64), ("row1", C.c_int64), ("col0", C.c_int64), ("col1", C.c_int64),
                ("ready_event", C.c_void_p)]


class GridPoint(C.Structure):
    """rvmc_grid_point."""
    _fields_ = [("pop", PopParams), ("seed", C.c_uint64)]


# name -> (restype, argtypes); also the list of symbols the library must export
_VP, _I, _I64, _U64, _U32, _D, _F = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_uint32, C.c_double, C.c_float
SIGNATURES = {
    "rvmc_abi_version": (_I, []),
    "rvmc_last_error": (C.c_char_p, []),
    "rvmc_pop_defaults": (None, [C.POINTER(PopParams)]),
    "rvmc_device_info": (_I, [_I] + [C.POINTER(C.c_int)] * 4),
    "rvmc_pipe_peak": (_I, [_I, _I, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "rvmc_philox_blocks": (_I, [_U64, _U64, _I64, _U32, _U32, _VP, _VP]),
    "rvmc_sample_orbits": (_I, [C.POINTER(PopParams), _U64, _U64, _I64, _U32, OrbitSoA, _VP, _VP]),
    "rvmc_normal_f64": (_I, [_U64, _U64, _I64, _U32, _U32, _D, _D, _I, _VP, _VP]),
    "rvmc_kepler_f64": (_I, [_I64, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rvmc_rv_replay_f64": (_I, [_I64, OrbitSoA, _VP, _VP, _VP]),
    "rvmc_orbit_scales_f64": (_I, [_I64, OrbitSoA, _VP, _VP, _VP, _VP]),
    "rvmc_rv_from_vmax_f64": (_I, [_I64, OrbitSoA, _VP, _VP, _VP]),
    "rvmc_rv_replay_f32": (_I, [_I64, OrbitSoA, C.POINTER(PopParams), _VP, _VP, _VP, _VP]),
    "rvmc_mix_pool": (_I, [_U64, _U32, _I64, _D, _VP, _VP, _VP, _VP]),
    "rvmc_sigma1d_pool": (_I, [_U64, _U32, _VP, _I64, _VP, _I, _I, _I, _U64, _I64, _VP, _VP]),
    "rvmc_sigma1d_fused": (_I, [C.POINTER(PopParams), _U64, _VP, _I, _I, _I, _U64, _I64, _VP, _VP, _VP]),
    "rvmc_fused_trace": (_I, [C.POINTER(PopParams), _U64, _VP, _I, _U64, _I64, TraceSoA, _VP]),
    "rvmc_biascorr_fused": (_I, [C.POINTER(PopParams), _U64, _U64, _I, _I, _I, _VP, _VP, _VP, _I, _VP, _VP, _I, _F,
                                 _F, _U64, _I64, _I64, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rvmc_biascorr_fused_packed": (_I, [C.POINTER(PopParams), _U64, _U64, _I, _I, _I, _VP, _VP, _VP, _I, _VP, _VP, _I,
                                        _F, _F, _U64, _I64, _I64, _VP, _VP, _VP, _I64, _VP, _VP, _VP, _VP]),
    "rvmc_unpack_bits_host": (_I, [_VP, _I64, _I64, _I64, _VP, _I64, _I]),
    "rvmc_count_bits_host": (_I, [_VP, _I64, _I64, _I64, _VP, _I]),
    "rvmc_copy_host": (_I, [_VP, _VP, _I64, _I]),
    "rvmc_fetch_flags_host": (_I, [_VP, _I64, _I64, C.POINTER(FlagBlock), _I, _VP, _VP, _VP, _I64, _I]),
    "rvmc_biascorr_sample": (_I, [C.POINTER(PopParams), _U64, _I, _I, _VP, _VP, _U64, _I64, _U32, OrbitSoA, _VP]),
    "rvmc_kepler_epochs_f64": (_I, [_I, _I64, _VP, _VP, _VP, _I, _VP, _VP, _VP]),
    "rvmc_rv_epochs_f64": (_I, [_I, _I64, _VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "rvmc_biascorr_replay_f64": (_I, [_I, _I64, _I, _VP, _VP, OrbitSoA, _I64, _VP, _U64, _U64, _I, _VP, _VP, _VP]),
    "rvmc_biascorr_singles": (_I, [_U64, _I, _I, _VP, _VP, _U64, _I64, _VP, _VP]),
    "rvmc_biascorr_detect": (_I, [_I, _I64, _I, _VP, _VP, _VP, _D, _D, _VP, _VP]),
    "rvmc_grid_scratch_slots": (_I, [_I]),
    "rvmc_grid_run_fbins": (_I, [C.POINTER(GridPoint), _I64, C.POINTER(C.c_double), _I, _VP, _I, _I, _I, _I64, _D, _I,
                                 _VP, _VP, _VP, _VP, _VP, _VP]),
    "rvmc_grid_run": (_I, [C.POINTER(GridPoint), _I64, _VP, _I, _I, _I, _I64, _D, _I, _VP, _VP, _VP, _VP,
                           _VP, _VP]),
    "rvmc_grid_best_fit_host": (_I, [_VP, _I64, _I, _I, _D, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                                     C.POINTER(C.c_uint64), _I]),
}

_lib = None


def lib():
    """Load librvmc_b200.so (built by __graft_entry__.build() / rvmc_b200/build.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "rvmc_b200: %s is missing -- build it with `python -m rvmc_b200.build`; there is no CPU "
            "fallback for the hot path" % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(l, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if l.rvmc_abi_version() != 2:
        raise RuntimeError("rvmc_b200: ABI version mismatch")
    _lib = l
    return l


class RvmcError(RuntimeError):
    pass


def check(rc):
    """Map a C status to the exception the reference would raise at the same place."""
    if rc == RVMC_OK:
        return
    msg = lib().rvmc_last_error().decode("utf-8", "replace")
    if rc == RVMC_ERR_ZERODIV:
        raise ZeroDivisionError(msg or "float division by zero")     # RVMC.py:112
    if rc == RVMC_ERR_INVALID:
        raise ValueError(msg)
    if rc == RVMC_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == RVMC_ERR_NO_DEVICE:
        raise RvmcError("rvmc_b200 needs an sm_100 (B200) GPU; no CPU fallback exists: " + msg)
    raise RvmcError("CUDA error %d: %s" % (rc, msg))
