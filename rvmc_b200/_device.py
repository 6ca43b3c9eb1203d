"""Device plumbing shared by the drop-in classes: PyTorch supplies device memory and streams,
nothing else.  Every computation goes through librvmc_b200.so (`_abi.lib()`); if the library or
the GPU is missing the calls raise -- there is no CPU path in this package.
"""
import ctypes as C

import numpy as np

from . import _abi

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def resolve_device(device=None):
    """`device` kwarg of the drop-in classes -> torch.device('cuda', k).  Raises when there is no GPU."""
    t = torch()
    if not t.cuda.is_available():
        raise _abi.RvmcError("rvmc_b200 needs an sm_100 (B200) GPU; torch.cuda.is_available() is False and "
                             "there is no CPU fallback")
    _abi.lib()      # fail early and loudly when the extension is missing
    if device is None:
        return t.device("cuda", t.cuda.current_device())
    d = t.device(device)
    if d.type != "cuda":
        raise ValueError("rvmc_b200 runs on CUDA devices only (got %r)" % (device,))
    if d.index is None:
        d = t.device("cuda", t.cuda.current_device())
    return d


def stream_ptr(device):
    return C.c_void_p(torch().cuda.current_stream(device).cuda_stream)


def ptr(tensor):
    """Device pointer of a tensor (or NULL for None) as c_void_p."""
    if tensor is None:
        return C.c_void_p(0)
    return C.c_void_p(tensor.data_ptr())


def empty(shape, dtype, device):
    t = torch()
    return t.empty(shape, dtype=getattr(t, dtype), device=device)


def zeros(shape, dtype, device):
    t = torch()
    return t.zeros(shape, dtype=getattr(t, dtype), device=device)


def to_device(array, dtype, device):
    """Host array-like -> contiguous device tensor of `dtype` (a torch tensor passes through)."""
    t = torch()
    if isinstance(array, t.Tensor):
        return array.to(device=device, dtype=getattr(t, dtype)).contiguous()
    a = np.ascontiguousarray(np.asarray(array), dtype=np.dtype(dtype))
    if not a.flags.writeable:          # e.g. np.broadcast_to views: torch.from_numpy wants a writable buffer
        a = a.copy()
    return t.from_numpy(a).to(device)


def to_host(tensor):
    """Device tensor -> NumPy array.  Large copies land in pinned host memory from torch's caching
    host allocator (fast DMA, no per-call page pinning after the first use); the array keeps the
    pinned block alive and returns it to the cache when it is garbage collected."""
    t = torch()
    tensor = tensor.detach()
    if tensor.is_cuda and tensor.numel() * tensor.element_size() >= (1 << 20):
        tensor = tensor.contiguous()
        host = t.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True)
        host.copy_(tensor, non_blocking=True)
        t.cuda.current_stream(tensor.device).synchronize()
        return host.numpy()
    return tensor.cpu().numpy()


_copy_streams = {}
_side_streams = {}


def side_stream(device):
    """A second compute stream per device: independent launches (disjoint star ranges of one pass)
    alternate between it and the current stream so that the tail of one overlaps the head of the next."""
    t = torch()
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = t.cuda.Stream(device=device)
    return _side_streams[key]


def to_host_overlapped(tensor, chunks):
    """Blocks of a [nstars, nsamples] device tensor -> one pinned NumPy array, copied block by block on
    a side stream as soon as the kernel that produced each block has finished.  `chunks` is a list of
    (row0, row1, col0, col1, event) with `event` recorded after that block's kernel: the D2H transfer of
    block k overlaps the compute of blocks k+1.. (the copy engine and the SMs work concurrently)."""
    t = torch()
    dev = tensor.device
    key = (dev.type, dev.index)
    if key not in _copy_streams:
        _copy_streams[key] = t.cuda.Stream(device=dev)
    cs = _copy_streams[key]
    host = t.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True)
    with t.cuda.stream(cs):
        for r0, r1, c0, c1, ev in chunks:
            cs.wait_event(ev)
            host[r0:r1, c0:c1].copy_(tensor[r0:r1, c0:c1], non_blocking=True)
    cs.synchronize()
    return host.numpy()


def soa(**tensors):
    """rvmc_orbit_soa from keyword tensors (missing = NULL)."""
    s = _abi.OrbitSoA()
    for k, v in tensors.items():
        setattr(s, k, v.data_ptr() if v is not None else None)
    return s


class DeviceGuard:
    """`with DeviceGuard(dev):` makes `dev` the current CUDA device for the library calls inside."""

    def __init__(self, device):
        self.device = device

    def __enter__(self):
        self._ctx = torch().cuda.device(self.device)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        return self._ctx.__exit__(*exc)


def splitmix64(x):
    """Seed derivation for successive draws of one object (each call of a sample_* method must see
    fresh numbers, like successive np.random calls in the reference)."""
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def fresh_seed():
    """Seed used when the caller gives none: OS entropy (the reference never seeds np.random)."""
    import os
    return int.from_bytes(os.urandom(8), "little")


class LazyArrays:
    """Mixin: attributes named in `_ARRAY_ATTRS` live on the device and become NumPy arrays when
    read.  Rule: once an attribute has been read or assigned on the host, the HOST copy is
    authoritative (the user may mutate it in place, as scripts written for the reference do) and
    is uploaded again whenever a kernel needs it; untouched attributes never leave the device."""

    _ARRAY_ATTRS = ()

    def _lazy_init(self, device):
        object.__setattr__(self, "_dev", {})
        object.__setattr__(self, "_host", {})
        object.__setattr__(self, "_device", device)

    def __getattr__(self, name):
        # only called when normal lookup fails
        if name in type(self)._ARRAY_ATTRS:
            d = self.__dict__
            if "_host" in d and name in d["_host"]:
                return d["_host"][name]
            if "_dev" in d and name in d["_dev"]:
                arr = to_host(d["_dev"].pop(name))
                d["_host"][name] = arr
                return arr
        raise AttributeError("%r object has no attribute %r" % (type(self).__name__, name))

    def __setattr__(self, name, value):
        if name in type(self)._ARRAY_ATTRS and "_host" in self.__dict__:
            self._dev.pop(name, None)
            self._host.pop(name, None)
            t = torch()
            if isinstance(value, t.Tensor) and value.is_cuda:
                self._dev[name] = value
            else:
                self._host[name] = value        # kept as given (None, list, ndarray): reference semantics
            return
        object.__setattr__(self, name, value)

    def _set_device(self, name, tensor):
        self._host.pop(name, None)
        self._dev[name] = tensor

    def _on_device(self, name, dtype="float64"):
        """Device tensor of an array attribute for a kernel call (uploads host-authoritative data)."""
        if name in self._dev:
            return self._dev[name]
        if name in self._host and self._host[name] is not None:
            return to_device(self._host[name], dtype, self._device)
        raise AttributeError("attribute %r has not been computed yet" % name)

    def _has(self, name):
        return name in self._dev or (name in self._host and self._host[name] is not None)
