"""Helpers for the `-m gpu` parity tests: device buffers via torch, calls through the C ABI."""
import ctypes as C
import os

import numpy as np

from rvmc_b200 import _abi
from rvmc_b200 import _device as dv

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dev():
    return dv.resolve_device(None)


def lib():
    return _abi.lib()


def stream():
    return dv.stream_ptr(dev())


_KEEP = []


def up(a, dtype):
    """Upload and KEEP ALIVE until the end of the test: `dv.ptr(up(x))` inline would otherwise
    free the tensor before the (asynchronous) kernel reads it."""
    t = dv.to_device(np.asarray(a), dtype, dev())
    _KEEP.append(t)
    if len(_KEEP) > 256:
        dv.torch().cuda.synchronize()
        del _KEEP[:128]
    return t


def empty(shape, dtype):
    return dv.empty(shape, dtype, dev())


def zeros(shape, dtype):
    return dv.zeros(shape, dtype, dev())


def host(t):
    return dv.to_host(t)


def ok(rc):
    _abi.check(rc)


def golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    draws = [z[k] for k in sorted(k for k in z.files if k.startswith("draw_"))]
    return z, draws


def orbit_soa_from(z, names=("primary_mass", "period", "time", "mass_ratio", "eccentricity", "inclination",
                             "orbit_rotation", "eccentric_anomaly")):
    tens = {k: up(np.ravel(z[k]), "float64") for k in names if k in z}
    return tens, dv.soa(**tens)


def pop_from_golden(z):
    kw = {}
    for k in z.files:
        if k.startswith("kw_") and k != "kw_names":
            kw[k[3:]] = float(z[k])
    return _abi.PopParams.defaults(**kw)


def rel_err(got, want, scale):
    """|got - want| / max(|want|, scale): the parity metric of SURVEY H1."""
    return np.abs(np.asarray(got, dtype=np.float64) - want) / np.maximum(np.abs(want), scale)
