"""Tracing of the host side (the reference has bare `print`s and no tracing at all).

`span(name)` brackets a phase of a host routine:
  * with RVMC_TRACE=1 (or `enable()`), every span is an NVTX range (visible in Nsight Systems / ncu
    timelines next to the kernels it launched) and is timed; the timings accumulate in `totals()` and go
    to the `rvmc_b200` logger at DEBUG level;
  * with RVMC_TRACE=sync the device is synchronised at both ends of a span, so a span's time is the time
    of the GPU work it issued (a measuring mode: it serialises the streams);
  * otherwise a span costs one attribute lookup.

`log` is the package logger (`logging.getLogger("rvmc_b200")`); the drop-in classes keep the reference's
user-facing `print`s (scripts parse them) and send everything else there.
"""
import contextlib
import logging
import os
import time

log = logging.getLogger("rvmc_b200")

_mode = os.environ.get("RVMC_TRACE", "")
_totals = {}


def enable(mode="1"):
    """mode: "" off, "1" NVTX ranges + host timings, "sync" also synchronise the device around spans."""
    global _mode
    _mode = mode or ""


def enabled():
    return bool(_mode)


def totals(reset=False):
    """{span name: [calls, seconds]} accumulated since the last reset."""
    out = {k: list(v) for k, v in _totals.items()}
    if reset:
        _totals.clear()
    return out


@contextlib.contextmanager
def _span(name):
    import torch
    sync = _mode == "sync" and torch.cuda.is_available()
    if sync:
        torch.cuda.synchronize()
    nvtx = torch.cuda.is_available()
    if nvtx:
        torch.cuda.nvtx.range_push(name)
    t0 = time.perf_counter()
    try:
        yield
    finally:
        if sync:
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if nvtx:
            torch.cuda.nvtx.range_pop()
        rec = _totals.setdefault(name, [0, 0.0])
        rec[0] += 1
        rec[1] += dt
        log.debug("%s: %.3f ms", name, dt * 1e3)


_NULL = contextlib.nullcontext()


def span(name):
    return _span(name) if _mode else _NULL
