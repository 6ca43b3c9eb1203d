"""Builds (g++) and loads tests/_build/libhostcheck.so -- test infrastructure only."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
OUT = os.path.join(HERE, "_build", "libhostcheck.so")
_DEPS = [SRC] + [os.path.join(HERE, "..", "rvmc_b200", "csrc", f)
                 for f in ("philox.cuh", "rvmc_math.cuh", "orbit.cuh", "pop_derive.h", "rvmc_constants.h")] + \
        [os.path.join(HERE, "..", "include", "rvmc_b200.h")]
_lib = None


def load_hostcheck():
    global _lib
    if _lib is not None:
        return _lib
    stale = not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in _DEPS)
    if stale:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas",
                               "-o", OUT, SRC])
    _lib = ctypes.CDLL(OUT)
    return _lib
