"""Philox4x32-10 known answers (Random123 kat_vectors: philox4x32 10) for the NumPy model and for
the host-compiled copy of the device source; the GPU copy is checked in test_gpu_parity.py."""
import ctypes

import numpy as np
import pytest

from oracle import philox_ref as ph
from tests.hostcheck_lib import load_hostcheck

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_numpy_model_kat(ctr, key, want):
    got = ph.philox4x32_10(*[np.array([c]) for c in ctr], key[0], key[1])
    assert tuple(int(g[0]) for g in got) == want


@pytest.mark.parametrize("ctr,key,want", KAT)
def test_device_source_on_host_kat(ctr, key, want):
    lib = load_hostcheck()
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    out = (ctypes.c_uint32 * 4)()
    lib.hc_philox_raw(c, k, out)
    assert tuple(out) == want


def test_block_layout_matches_model():
    lib = load_hostcheck()
    seed = 0x1234_5678_9ABC_DEF0
    items = np.array([0, 1, 2 ** 32 - 1, 2 ** 32, 2 ** 40 + 17], dtype=np.uint64)
    want = ph.philox_block(seed, items, 6, 3)
    out = (ctypes.c_uint32 * 4)()
    for i, it in enumerate(items):
        lib.hc_philox(ctypes.c_uint64(seed), ctypes.c_uint64(int(it)), 6, 3, out)
        assert tuple(out) == tuple(int(v) for v in want[i])


def test_uniform_maps():
    w = np.array([0, 255, 256, 0xFFFFFFFF], dtype=np.uint32)
    assert ph.u01(w).tolist() == [0.0, 0.0, 2.0 ** -24, 1 - 2.0 ** -24]
    assert ph.u01_open(w).tolist() == [2.0 ** -24, 2.0 ** -24, 2.0 ** -23, 1.0]
