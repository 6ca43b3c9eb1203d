"""Device plumbing shared by the drop-in classes: PyTorch supplies device memory and streams,
nothing else.  Every computation goes through librvmc_b200.so (`_abi.lib()`); if the library or
the GPU is missing the calls raise -- there is no CPU path in this package.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._trace import span

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


# Pinned host memory comes from torch's caching host allocator, which rounds sizes up to a power of two
# and never returns blocks to the OS.  Results up to PIN_DIRECT_MAX land in a pinned block directly (the
# NumPy array keeps it alive; it goes back to the cache when the array is collected).  Anything larger
# is staged through two fixed pinned buffers into ordinary pageable memory, so reading a multi-GB
# attribute array cannot grow the locked memory of the process.
PIN_DIRECT_MAX = 512 << 20
PIN_STAGE_BYTES = 32 << 20
_stage = {}


def _stage_buffers(device):
    t = torch()
    key = (device.type, device.index)
    if key not in _stage:
        _stage[key] = ([t.empty(PIN_STAGE_BYTES, dtype=t.uint8, pin_memory=True) for _ in range(2)],
                       [t.cuda.Event(), t.cuda.Event()])
    return _stage[key]


def to_host(tensor):
    """Device tensor -> NumPy array (see PIN_DIRECT_MAX above for where the bytes land)."""
    t = torch()
    tensor = tensor.detach()
    nbytes = tensor.numel() * tensor.element_size()
    if not tensor.is_cuda or nbytes < (1 << 20):
        return tensor.cpu().numpy()
    tensor = tensor.contiguous()
    stream = t.cuda.current_stream(tensor.device)
    if nbytes <= PIN_DIRECT_MAX:
        host = t.empty(tensor.shape, dtype=tensor.dtype, pin_memory=True)
        host.copy_(tensor, non_blocking=True)
        stream.synchronize()
        return host.numpy()
    out = t.empty(tensor.shape, dtype=tensor.dtype)          # pageable
    src = tensor.view(-1).view(t.uint8)
    dst = out.view(-1).view(t.uint8)
    bufs, evs = _stage_buffers(tensor.device)
    lib, threads = _abi.lib(), host_threads()
    n_chunks = (nbytes + PIN_STAGE_BYTES - 1) // PIN_STAGE_BYTES
    for k in range(n_chunks + 1):
        if k < n_chunks:                                     # start the copy of chunk k ...
            lo, hi = k * PIN_STAGE_BYTES, min(nbytes, (k + 1) * PIN_STAGE_BYTES)
            bufs[k & 1][:hi - lo].copy_(src[lo:hi], non_blocking=True)
            evs[k & 1].record(stream)
        if k >= 1:                                           # ... while chunk k-1 moves to its final place
            lo, hi = (k - 1) * PIN_STAGE_BYTES, min(nbytes, k * PIN_STAGE_BYTES)
            evs[(k - 1) & 1].synchronize()
            _abi.check(lib.rvmc_copy_host(dst.data_ptr() + lo, bufs[(k - 1) & 1].data_ptr(), hi - lo, threads))
    return out.numpy()


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


def _copy_stream(dev):
    t = torch()
    key = (dev.type, dev.index)
    if key not in _copy_streams:
        _copy_streams[key] = t.cuda.Stream(device=dev)
    return _copy_streams[key]


def _host_result(shape, dtype):
    """Host tensor for a result the caller keeps: pinned (cached by torch) up to PIN_DIRECT_MAX."""
    t = torch()
    nbytes = t.empty(0, dtype=dtype).element_size()
    for d in shape:
        nbytes *= int(d)
    return t.empty(shape, dtype=dtype, pin_memory=nbytes <= PIN_DIRECT_MAX)


def _copy_block(dst, src, r0, r1, c0, c1):
    """Asynchronous copy of the block [r0:r1, c0:c1]: whole rows go as one contiguous transfer, a column
    range goes row by row (each slice contiguous -- a strided 2-D slice would make torch stage the copy
    through pageable memory and lose the overlap)."""
    if c0 == 0 and c1 == src.shape[1]:
        dst[r0:r1].copy_(src[r0:r1], non_blocking=True)
    else:
        for r in range(r0, r1):
            dst[r, c0:c1].copy_(src[r, c0:c1], non_blocking=True)


def to_host_overlapped(tensor, chunks):
    """Blocks of a [nstars, nsamples] device tensor -> one NumPy array, copied block by block on a
    side stream as soon as the kernel that produced each block has finished.  `chunks` is a list of
    (row0, row1, col0, col1, event) with `event` recorded after that block's kernel: the D2H transfer of
    block k overlaps the compute of blocks k+1.. (the copy engine and the SMs work concurrently)."""
    t = torch()
    cs = _copy_stream(tensor.device)
    host = _host_result(tensor.shape, tensor.dtype)
    with t.cuda.stream(cs):
        for r0, r1, c0, c1, ev in chunks:
            cs.wait_event(ev)
            _copy_block(host, tensor, r0, r1, c0, c1)
    cs.synchronize()
    return host.numpy()


def host_threads():
    """Threads the host-side helpers of the library may use in this process: RVMC_HOST_THREADS, else the
    cores this process may run on divided by the ranks sharing the node (LOCAL_WORLD_SIZE), at most 8."""
    import os
    env = os.environ.get("RVMC_HOST_THREADS")
    if env:
        return max(1, int(env))
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:          # pragma: no cover
        cores = os.cpu_count() or 1
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    # the rank's share of the cores.  The widening is bound by the host's memory system long before that
    # (32-vCPU, 8-GPU box: 16 threads in total stream 205 GB/s, 32 threads 203 GB/s), but inside the real
    # pass, where the helper threads also sleep between blocks, 4 per rank measured best (8.6e11 against
    # 7.6e11 RV/s with 2)
    return max(1, min(8, cores // local_world))


def unpack_bits_overlapped(bits, n_cols, chunks, threads=None):
    """Bit-packed flags [nrows, nwords] (int32 words on the device, bit j of word w = column 32 w + j) ->
    NumPy bool [nrows, n_cols].  Block by block, as the kernels that produced them finish: the packed
    block crosses PCIe on a side stream (1/8 of the bytes of a bool array) and is widened on the host by
    rvmc_unpack_bits_host while later blocks are still being computed and copied.  `chunks` as for
    to_host_overlapped (columns in flags, multiples of 32), or None for one block."""
    t = torch()
    lib = _abi.lib()
    threads = host_threads() if threads is None else int(threads)
    nrows, nwords = bits.shape
    if chunks is None:
        ev = t.cuda.Event()
        ev.record(t.cuda.current_stream(bits.device))
        chunks = [(0, nrows, 0, n_cols, ev)]
    cs = _copy_stream(bits.device)
    packed = t.empty(bits.shape, dtype=bits.dtype, pin_memory=True)
    out = _host_result((nrows, n_cols), t.uint8)
    blocks = (_abi.FlagBlock * len(chunks))()
    for k, (r0, r1, c0, c1, ev) in enumerate(chunks):
        blocks[k].row0, blocks[k].row1, blocks[k].col0, blocks[k].col1 = r0, r1, c0, c1
        blocks[k].ready_event = ev.cuda_event
    # one library call does the rest (waits, copies, widening) without a Python loop in the way; ctypes
    # releases the GIL for its duration
    with DeviceGuard(bits.device), span("flags.fetch"):
        _abi.check(lib.rvmc_fetch_flags_host(bits.data_ptr(), nwords, n_cols, blocks, len(chunks), cs.cuda_stream,
                                             packed.data_ptr(), out.data_ptr(), n_cols, threads))
    return out.numpy().view(np.bool_)


def fetch_flags_hybrid(det, bits, chunks, byte_share, threads=None):
    """The detection flags of a pass whose kernels wrote BOTH forms (det: uint8 [nrows, ncols], bits: packed
    words) -> NumPy bool [nrows, ncols], using the two routes into host memory AT THE SAME TIME: about
    `byte_share` of the flags come as bytes by DMA straight into the result (the copy engines' path:
    PCIe / IOMMU bound), the rest come bit-packed and are widened by the CPU threads (the host memory
    system's path).  With several ranks on one host each route has its own ceiling (measured on an
    8 x B200, 32-vCPU box: 120 GB/s of DMA writes in aggregate, ~200 GB/s of CPU streaming stores), and a
    pass that must leave 100 MB of flags per rank in host memory every 3 ms needs both.  Blocks are assigned
    to a route by error diffusion over their sizes, so the two routes stay busy side by side."""
    t = torch()
    lib = _abi.lib()
    threads = host_threads() if threads is None else int(threads)
    nrows, ncols = det.shape
    nwords = bits.shape[1]
    dev = det.device
    if chunks is None:
        ev = t.cuda.Event()
        ev.record(t.cuda.current_stream(dev))
        chunks = [(0, nrows, 0, ncols, ev)]
    cs_bytes = _copy_stream(dev)
    key = (dev.type, dev.index, "bits")
    if key not in _copy_streams:
        _copy_streams[key] = t.cuda.Stream(device=dev)
    cs_bits = _copy_streams[key]
    out = _host_result((nrows, ncols), t.uint8)
    packed = t.empty(bits.shape, dtype=bits.dtype, pin_memory=True)
    acc, plan = 0.0, []
    for r0, r1, c0, c1, ev in chunks:
        size = (r1 - r0) * (c1 - c0)
        acc += byte_share * size
        as_bytes = acc >= 0.5 * size
        if as_bytes:
            acc -= size
        plan.append(as_bytes)
    done = []
    for (r0, r1, c0, c1, ev), as_bytes in zip(chunks, plan):
        if as_bytes:
            with t.cuda.stream(cs_bytes):
                cs_bytes.wait_event(ev)
                _copy_block(out, det, r0, r1, c0, c1)
            done.append(None)
        else:
            with t.cuda.stream(cs_bits):
                cs_bits.wait_event(ev)
                _copy_block(packed, bits, r0, r1, c0 // 32, (c1 + 31) // 32)
                e = t.cuda.Event()
                e.record(cs_bits)
            done.append(e)
    p_ptr, o_ptr = packed.data_ptr(), out.data_ptr()
    for (r0, r1, c0, c1, _), e in zip(chunks, done):
        if e is None:
            continue
        e.synchronize()
        _abi.check(lib.rvmc_unpack_bits_host(p_ptr + 4 * (r0 * nwords + c0 // 32), nwords, r1 - r0, c1 - c0,
                                             o_ptr + r0 * ncols + c0, ncols, threads))
    cs_bytes.synchronize()
    return out.numpy().view(np.bool_)


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


def call_seed(seed, call):
    """Seed of the `call`-th random draw site of an object created with `seed`.  The user seed is hashed
    once and the call counter enters through a multiplication by an odd constant in a DIFFERENT domain,
    so objects created with neighbouring seeds (replicate studies over seed = 0..N) never walk into each
    other's sequences -- with splitmix64(seed + call) object `s` at call k met object `s + 1` at call k - 1."""
    return splitmix64(splitmix64(int(seed) & 0xFFFFFFFFFFFFFFFF) ^ ((int(call) * 0xD1342543DE82EF95) & 0xFFFFFFFFFFFFFFFF))


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
