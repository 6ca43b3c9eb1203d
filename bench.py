#!/usr/bin/env python
"""Headline benchmark: binary-epoch RVs per second on the multi-epoch detection workload
(BASELINE.json configs[1]: 10^8 binaries x 6 epochs, Delta-RV > 20 km/s at 4 sigma), plus one leg per
other BASELINE config (C1 example cluster, C3 six data_sana variants, C4 data_RT20 clusters, C5 grid).

    python bench.py --gpus 1 --steps K --warmup W                 # this framework (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference itself on the host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU, weak scaling

A step = one pass of the hot path (rvmc_biascorr_fused_packed: sample + Kepler + RV at 6 epochs + noise +
Delta-RV + significance test) over 10^8 binaries per GPU, writing Delta-RV (fp32) and the detection
flags (bit-packed when ranks share a host, one byte each for a single rank: --flag-transfer auto) for every
binary to HBM -- the outputs of the product path.  `value` is timed with
CUDA events on the launching stream with everything resident on the device; `e2e` is the same metric
through the drop-in Python class (`BiasCorr(...)` -> `calculate_radial_velocities()` ->
`get_detected_binaries()`) with host inputs and the bool detection array back in host memory inside the
timed region.

Only this file's `cpu_baseline` leg and `--impl reference` touch `oracle/` (the unmodified reference
modules under oracle/_ref when present, else the CPU restatement); the product path never does.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DAYS6 = np.array([0, 1, 7, 30, 180, 365.0]) * 86400.0
METRIC = "binary_epoch_rvs_per_sec"
UNIT = "binary-epoch RV/s"

# ALGORITHMIC work per binary-epoch RV
ALG_MULTI = {"fp32_flop": 100.0, "sfu_ops": 17.0, "int_ops": 55.0, "fp64_flop": 6.0}       # multi-epoch path (C2)
ALG_SINGLE = {"fp32_flop": 160.0, "sfu_ops": 30.0, "int_ops": 250.0, "fp64_flop": 0.0}     # single-epoch fused path
# EXECUTED thread-instructions per unit and the issue-slot utilisation ncu saw, from the committed
# captures of the same kernels (profiles/README.md says how they are read); profiles/executed_counts.json
# (written from the newest captures) overrides these round-1 figures.
EXEC_MULTI = {"lane_instr_per_unit": 142.4, "issue_slots_busy": 0.776, "traffic_bytes": 442.62e6,
              "source": "profiles/r01_biascorr_fused_fullsize_ncu.md"}
EXEC_SINGLE = {"lane_instr_per_unit": 436.0, "issue_slots_busy": 0.75, "source": "profiles/r01_sigma1d_fused_ncu.md"}
PROFILE_OVERRIDES = os.path.join(ROOT, "profiles", "executed_counts.json")
if os.path.exists(PROFILE_OVERRIDES):
    with open(PROFILE_OVERRIDES) as _f:
        _o = json.load(_f)
    EXEC_MULTI.update(_o.get("multi", {}))
    EXEC_SINGLE.update(_o.get("single", {}))

M17 = np.array([0.8, 2.3, 1.0, 1.8, 0.8, 1.0, 1.4, 2.5, 2.0, 0.4, 1.5, 0.6])      # findBestDistributions.py:292
# /root/reference/data_RT20/Nstars_massRange.txt:2-11 (name, observed sigma_1D [km/s], N stars, mass range)
RT20 = (("IC1805", 65.45329203402903, 8, 15, 60), ("IC1848", 50.26275741127991, 5, 15, 60),
        ("IC2944", 31.357122212648363, 14, 15, 60), ("NGC6231", 67.61903406110999, 13, 15, 60),
        ("NGC6611", 25.32524131575383, 9, 15, 60), ("Wd2", 14.999765386529885, 44, 6, 60),
        ("M8", 32.69110237188463, 16, 6, 20), ("NGC6357", 26.850667796133674, 22, 6, 30),
        ("R136", 25.0, 332, 15, 60), ("M17", 5.5, 12, 6, 20))
# C3: P_min in {1.4 d, 30 d, 8 yr} at f_bin 0.7 and f_bin in {0.12, 0.28, 0.70} at P_min 1.4 d;
# the data_sana file each variant is named after gives the sanity band (tests/golden/data_sana_band.json)
C3_VARIANTS = (("period_1.4d", 1.4, 0.70), ("period_1mo", 30.0, 0.70), ("period_8yr", 2922.0, 0.70),
               ("fbin_0.12", 1.4, 0.12), ("fbin_0.28", 1.4, 0.28), ("fbin_0.70", 1.4, 0.70))


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference on the host cores
# ---------------------------------------------------------------------------------------------
def load_reference():
    """(RVMC, RVMC_bias_corr) modules of the unmodified reference from oracle/_ref, or None."""
    try:
        from oracle.build_ref import load_reference as _load
        return _load()
    except Exception:          # pragma: no cover
        return None


_REF = None


def _cpu_chunk(args):
    """One bounded sample of the C2 workload on one core: `nstars` stars x `ns` samples x 6 epochs through
    RVMC_bias_corr.py:57-254,406-422,451-476 -- the reference's own BiasCorr when oracle/_ref is there
    (kind "reference"), else the oracle port (NumPy + SciPy-style vectorised secant, kind "port")."""
    seed, nstars, ns, use_ref = args
    if use_ref:
        import contextlib
        import io
        import warnings
        np.random.seed(seed)
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            bc = _REF[1].BiasCorr([DAYS6] * nstars, number_of_samples=ns, rv_errors=2.0)
            bc.calculate_radial_velocities()
            det = bc.get_detected_binaries(delta_RV=20, n_sigma_RV=4)
        return nstars * ns * len(DAYS6), int(det.sum())
    from oracle import rvmc_oracle as orc
    rng = np.random.RandomState(seed)
    shape = (nstars, ns)
    m1 = orc.masses_from_uniform(rng.random_sample(shape), 6, 20, -2.35)
    per = orc.period_from_uniform(rng.random_sample(shape), 10 ** 0.15, 10 ** 3.5, -0.5)
    t = orc.time_from_uniform(rng.random_sample(shape), per)
    q = orc.mass_ratio_from_uniform(rng.random_sample(shape), 0.1, 1.0, 0)
    e = np.stack([orc.eccentricity_from_uniform_slotted(per[i], rng.random_sample(ns), -0.5) for i in range(nstars)])
    inc = orc.inclination_from_uniform(rng.random_sample(shape))
    om = orc.orbit_rotation_from_uniform(rng.random_sample(shape))
    ndet = 0
    for i in range(nstars):
        rv = orc.multi_epoch_rv(DAYS6, t[i], per[i], e[i], q[i], m1[i], inc[i], om[i])
        rv = rv + rng.normal(0, 2.0, rv.shape)
        _ = orc.delta_rv(rv)
        ndet += int(orc.detected_binaries(rv, 2.0, 20, 4).sum())
    return nstars * ns * len(DAYS6), ndet


def cpu_pass(pool, cores, nstars, ns, seed0, use_ref):
    t0 = time.perf_counter()
    res = pool.map(_cpu_chunk, [(seed0 + c, nstars, ns, use_ref) for c in range(cores)])
    dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_kind():
    global _REF
    if _REF is None:
        _REF = load_reference() or False
    if _REF:
        return True, "reference", ("the unmodified reference (oracle/_ref/RVMC_bias_corr.py under oracle/shims): BiasCorr(...)"
                                   ".calculate_radial_velocities(); get_detected_binaries(20, 4), one process per core")
    return False, "port", "oracle port of RVMC_bias_corr.py (NumPy + SciPy-style vectorised secant), one process per core"


def cpu_c1_seconds():
    """BASELINE configs[0] on ONE host core with the reference itself (RVMC_example.py:6-37 shape)."""
    if not _REF:
        return None
    import contextlib
    import io
    import warnings
    best = None
    for rep in range(2):
        np.random.seed(5 + rep)
        with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
            warnings.simplefilter("ignore")
            t0 = time.perf_counter()
            c = _REF[0].RVMC(number_of_stars=10 ** 5)
            c.calculate_binary_velocities()
            c.subsample(sample_size=100, measured_errors=1.2, number_of_samples=10 ** 4)
            dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU implementation on all host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    use_ref, kind, how = cpu_kind()
    cores = host_cores()
    nstars, ns = 2, 50000        # per core and step: 6e5 binary-epoch RVs (~0.7 s at ~9e5 RV/s/core)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            cpu_pass(pool, cores, nstars, ns, 1000 * w, use_ref)
        total, elapsed = 0, 0.0
        for k in range(args.steps):
            n, dt = cpu_pass(pool, cores, nstars, ns, 77 + 1000 * k, use_ref)
            total += n
            elapsed += dt
    value = total / elapsed
    sample = "%d cores x %d stars x %d samples x 6 epochs per step; %s" % (cores, nstars, ns, how)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(n_gpus, nstars=100, ns=10 ** 6, out_bytes=None):
    out_bytes = nstars * ns * 4 + nstars * ((ns + 31) // 32) * 4 if out_bytes is None else out_bytes
    return {"workload": "BiasCorr multi-epoch detection (BASELINE configs[1]): %d stars x %d samples = %.0e binaries "
                        "per GPU x 6 epochs [0,1,7,30,180,365] d, IMF masses 6-20, Sana+2012 defaults, rv_errors 2.0 "
                        "km/s, Delta-RV > 20 km/s at 4 sigma" % (nstars, ns, nstars * ns),
            "binaries_per_gpu": nstars * ns, "epochs": 6, "parallelism": "dp%d (samples sharded, no collective)" % n_gpus,
            "l2": "inputs are < 8 KB of tables; outputs (%.0f MB/step) exceed the 126 MB L2" % (out_bytes / 1e6)}


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz, self.power = [], set(), None, []
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as exc:       # pragma: no cover
            self.err = repr(exc)

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:          # pragma: no cover
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def issue_roofline(units_per_s, alg, executed, sms, clk_hz, kernel, hbm_bytes_per_unit, peaks, peak_src, pipes,
                   launch_ms=None, traffic=None, traffic_src=None):
    """Roofline block of a kernel bound by instruction issue (no dense contraction, < 1 B of HBM per unit:
    neither `tensor` nor `hbm` binds).  peak = SMs x 4 schedulers x 32 lanes x the SM clock NVML
    reported during the timed region.  `frac` is the EXECUTED figure: thread-instructions per unit of the
    committed ncu capture of this kernel x units/s of the live CUDA-event timing / peak -- the share of
    the machine's issue slots the kernel keeps busy.  `frac_algorithmic` (SURVEY 8d's operation count x
    units/s / peak) is kept beside it: it exceeds `frac` when the kernel needs fewer instructions than the
    survey budgeted, which says nothing about utilisation."""
    alg_instr = alg["fp32_flop"] + alg["sfu_ops"] + alg["int_ops"] + alg["fp64_flop"]
    peak = sms * 4 * 32 * clk_hz
    roof = {"bound": "issue (instruction pipes; not hbm, not tensor)", "kernel": kernel,
            "achieved": units_per_s * executed["lane_instr_per_unit"] / 1e12, "peak": peak / 1e12,
            "unit": "T lane-instr/s", "frac": units_per_s * executed["lane_instr_per_unit"] / peak,
            "frac_definition": "executed thread-instructions per unit (ncu capture of this kernel) x units/s (live CUDA "
                               "events) / issue ceiling",
            "executed_lane_instr_per_unit_ncu": executed["lane_instr_per_unit"],
            "issue_slots_busy_ncu": executed["issue_slots_busy"], "ncu": executed["source"],
            "frac_algorithmic": units_per_s * alg_instr / peak,
            "algorithmic_per_unit": dict(alg, lane_instr=alg_instr, hbm_bytes=hbm_bytes_per_unit),
            "peak_source": "%d SMs x 4 schedulers x 32 lanes x %.0f MHz (NVML, sampled during the timed region)"
                           % (sms, clk_hz / 1e6),
            "traffic": traffic, "traffic_source": traffic_src,
            "hbm": {"achieved": units_per_s * hbm_bytes_per_unit / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                    "peak_source": "MEASURED_PEAKS.json" if peak_src == "measured" else "fallback 6650 GB/s",
                    "frac": units_per_s * hbm_bytes_per_unit / 1e9 / peaks.get("hbm_gbs", 6650.0)}}
    if launch_ms is not None:
        roof["launch_ms"] = launch_ms
    if pipes:
        # the individual pipes, against rates measured live on this device by rvmc_pipe_peak
        roof["fp32"] = {"achieved": units_per_s * alg["fp32_flop"] / 1e12, "peak": 2.0 * pipes["ffma"] / 1e12,
                        "unit": "TFLOP/s", "frac": units_per_s * alg["fp32_flop"] / (2.0 * pipes["ffma"])}
        roof["sfu"] = {"achieved": units_per_s * alg["sfu_ops"] / 1e12, "peak": pipes["mufu"] / 1e12,
                       "unit": "T MUFU-op/s", "frac": units_per_s * alg["sfu_ops"] / pipes["mufu"]}
        if alg["fp64_flop"]:
            roof["fp64"] = {"achieved": units_per_s * alg["fp64_flop"] / 1e12, "peak": 2.0 * pipes["dfma"] / 1e12,
                            "unit": "TFLOP/s", "frac": units_per_s * alg["fp64_flop"] / (2.0 * pipes["dfma"])}
    return roof


def run_gpu_arm(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from rvmc_b200 import _abi
    from rvmc_b200 import _device as dv
    from rvmc_b200.bias_corr import BiasCorr
    from rvmc_b200.rvmc import RVMC

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _abi.lib()
    p = _abi.PopParams.defaults()
    nstars, ns = args.stars, args.samples
    first_sample = rank * ns                 # ranks take disjoint sample ranges of the same stars
    seed, noise_seed = 1234, 4321
    st = dv.stream_ptr(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def event_ms(fn, reps, warm):
        """Mean CUDA-event time of `fn` over `reps` calls on the current stream, max over ranks."""
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / reps)

    # ---- C2, the headline: kernel-only `value` -------------------------------------------------------
    d_nep = dv.to_device(np.full(nstars, 6, np.int32), "int32", dev)
    d_tobs = dv.to_device(np.tile(DAYS6, (nstars, 1)), "float64", dev)
    d_err = dv.to_device(np.full((nstars, 6), 2.0, np.float32), "float32", dev)
    drv = dv.empty((nstars, ns), "float32", dev)
    # the outputs of the product path (BiasCorr.calculate_radial_velocities): Delta-RV per binary and the
    # detection flags, bit-packed (or one byte each with --flag-transfer bytes)
    BiasCorr.flag_transfer = args.flag_transfer
    if args.flag_transfer == "auto":          # what the class resolves it to in this job
        args.flag_transfer = "bits" if int(os.environ.get("LOCAL_WORLD_SIZE", "1")) > 1 else "bytes"
    nw = (ns + 31) // 32
    det = dv.empty((nstars, ns), "uint8", dev) if args.flag_transfer in ("bytes", "hybrid") else None
    bits = dv.empty((nstars, nw), "int32", dev) if args.flag_transfer in ("bits", "hybrid") else None
    out_bytes = nstars * ns * 4 + (nstars * ns if det is not None else 0) + (nstars * nw * 4 if bits is not None else 0)
    cnt = dv.zeros((4,), "int64", dev)

    def step():
        _abi.check(lib.rvmc_biascorr_fused_packed(C.byref(p), seed, noise_seed, 0, nstars, 6, dv.ptr(d_nep),
                                                  dv.ptr(d_tobs), dv.ptr(d_err), 0, None, None, 1, 20.0, 4.0,
                                                  first_sample, ns, ns, dv.ptr(drv), dv.ptr(det), dv.ptr(bits), nw,
                                                  None, dv.ptr(cnt), None, st))

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    cnt.zero_()
    sampler = ClockSampler(local)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for k in range(args.steps):
        step()
        ev[k + 1].record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    counters = dv.to_host(cnt)
    units_per_step = nstars * ns * 6
    assert int(counters[0]) == units_per_step * args.steps, "kernel did not process the whole workload"
    assert int(counters[3]) == 0, "non-converged Kepler lanes"
    total_ms_max = max_over_ranks(total_ms)
    value = world * units_per_step * args.steps / (total_ms_max * 1e-3)
    launches = args.steps

    # ---- e2e: the drop-in class with host inputs and the bool detection array back on the host ---------
    t_obs_host = [DAYS6.copy() for _ in range(nstars)]
    h2d = nstars * 6 * 8 + nstars * 6 * 4 + nstars * 4
    share = BiasCorr.flag_hybrid_byte_share
    d2h = {"bytes": nstars * ns, "bits": nstars * nw * 4,
           "hybrid": int(share * nstars * ns + (1 - share) * nstars * nw * 4)}[args.flag_transfer]
    e2e_steps = max(3, min(args.steps, 10))

    from rvmc_b200 import _trace

    def e2e_step(k, read_delta_rv=False):
        with _trace.span("biascorr.construct"):
            bc = BiasCorr(t_obs_host, number_of_samples=ns, rv_errors=2.0, seed=1000 + k + 7919 * rank, device=dev)
        bc.calculate_radial_velocities(delta_RV=20, n_sigma_RV=4)
        d = bc.get_detected_binaries(delta_RV=20, n_sigma_RV=4)
        if read_delta_rv:
            return d, bc.delta_rv
        return d, None

    def e2e_time(steps, read_delta_rv):
        for k in range(2):
            d, x = e2e_step(-1 - k, read_delta_rv)
        barrier()
        t0 = time.perf_counter()
        for k in range(steps):
            d, x = e2e_step(k, read_delta_rv)
            assert d.shape == (nstars, ns) and d.dtype == np.bool_
            assert x is None or x.shape == (nstars, ns)
        barrier()
        return max_over_ranks(time.perf_counter() - t0), float(d.mean())

    e2e_s, frac_det = e2e_time(e2e_steps, False)
    e2e_value = world * units_per_step * e2e_steps / e2e_s
    # where the host side of a step goes (rank 0): a second, traced run of the same steps -- host timers and NVTX
    # ranges only, no extra synchronisation; not the timed run above
    _trace.enable("1")
    _trace.totals(reset=True)
    e2e_time(e2e_steps, False)
    e2e_phases = {k: round(1e3 * v[1] / (e2e_steps + 2), 3) for k, v in _trace.totals(reset=True).items()}
    _trace.enable("")
    # what a user who also reads `.delta_rv` (RVMC_bias_corr.py:422) pays: 4 B per binary more over PCIe
    e2e_drv_steps = 3
    e2e_drv_s, _ = e2e_time(e2e_drv_steps, True)
    e2e_drv_value = world * units_per_step * e2e_drv_steps / e2e_drv_s

    # ---- C5: grid points per second (BASELINE configs[4]) ------------------------------------------------
    grid_info = grid_indep = None
    if not args.no_grid:
        from rvmc_b200 import grid as rgrid
        axes = dict(period_pi=np.linspace(-0.95, 0.5, 64), mass_ratio_kappa=np.linspace(-0.95, 1.0, 64),
                    fbin=np.linspace(0.0625, 1, 16))
        rgrid.grid_search({k: v[:4] for k, v in axes.items()}, number_of_samples=1000, seed=1, device=dev)

        def time_grid(reps, **kw):
            times, res = [], None
            rgrid.grid_search(axes, number_of_samples=10 ** 5, seed=2023, observed_sigma=25.0, device=dev, **kw)   # warm-up
            for rep in range(reps):
                barrier()
                t0 = time.perf_counter()
                res = rgrid.grid_search(axes, number_of_samples=10 ** 5, seed=2024, observed_sigma=25.0, device=dev, **kw)
                barrier()
                times.append(time.perf_counter() - t0)
            return max_over_ranks(min(times)), res

        timing = ("host wall clock, barrier + synchronize on both sides, best of %d, max over ranks; includes building "
                  "the rank's share of the grid, the all-reduce and the host copy of the statistics table")
        tg, gres = time_grid(5)
        npts = int(np.prod(gres["shape"]))
        grid_info = {"workload": "findBestDistributions grid (BASELINE configs[4]): 64 x 64 x 16 (pi, kappa, fbin) x 10^5 "
                                 "realizations x 12 stars (M17 errors), statistics + best fit; f_bin axis evaluated in one "
                                 "pass per (pi, kappa) point (shared draws); points dealt to the ranks, 80-byte statistics "
                                 "all-reduced",
                     "points": npts, "seconds": tg, "points_per_s": npts / tg, "scaling": "strong",
                     "timing": timing % 5, "best_index": list(gres["best_index"]),
                     "binary_rvs_this_rank": int(gres["counters"][0])}
        launches += 6 * 2 * int(np.ceil(4096 / world / 37))
        tgi, gres_i = time_grid(2, share_fbin_draws=False)
        grid_indep = {"workload": "the same grid with independent draws in every cell (share_fbin_draws=False): 65 536 "
                                  "points dealt by cost (sample_size x fbin) so that every rank holds the same number of "
                                  "points of every binary fraction",
                      "points": npts, "seconds": tgi, "points_per_s": npts / tgi, "scaling": "strong",
                      "timing": timing % 2, "best_index": list(gres_i["best_index"]),
                      "n_local_points": int(gres_i["n_local"])}
        launches += 3 * 2 * int(np.ceil(65536 / world / 592))

    # ---- C3: six data_sana-named variants, 10^9 binary orbits each per GPU (weak) -------------------------
    configs = {}
    if not args.no_configs:
        d_m17 = dv.to_device(M17, "float32", dev)
        band_path = os.path.join(ROOT, "tests", "golden", "data_sana_band.json")
        band = json.load(open(band_path)) if os.path.exists(band_path) else {}
        c3 = []
        c3_total_rv, c3_total_ms = 0, 0.0
        target = int(args.c3_orbits)
        for vi, (name, pmin, fbin) in enumerate(C3_VARIANTS):
            pop = _abi.PopParams.defaults(min_period=pmin, fbin=fbin)
            n_real = int(np.ceil(target / (12 * fbin)))
            sig = dv.empty((n_real,), "float32", dev)
            c3cnt = dv.zeros((4,), "int64", dev)
            first_real = rank * n_real

            def run(pop=pop, n_real=n_real, sig=sig, c3cnt=c3cnt, first_real=first_real, vi=vi):
                _abi.check(lib.rvmc_sigma1d_fused(C.byref(pop), vi + 1, dv.ptr(d_m17), 12, 0, 1, first_real, n_real,
                                                  dv.ptr(sig), dv.ptr(c3cnt), st))
            ms = event_ms(run, 2, 1)
            n_rv = int(dv.to_host(c3cnt)[0]) // 3
            launches += 3
            med = float(torch.median(sig).item())          # reporting only, outside the timed region
            c3.append({"variant": name, "min_period_d": pmin, "fbin": fbin, "realizations": n_real, "binary_rvs": n_rv,
                       "ms": ms, "binary_rvs_per_s": world * n_rv / (ms * 1e-3), "median_sigma_1d": med,
                       "data_sana_median": band.get(name, {}).get("median"),
                       "data_sana_p16_p84": [band.get(name, {}).get("p16"), band.get(name, {}).get("p84")]})
            c3_total_rv += n_rv
            c3_total_ms += ms
            del sig, run
        configs["C3"] = {"workload": "BASELINE configs[2]: six variants named after data_sana (P_min 1.4 d / 30 d / 8 yr at "
                                     "f_bin 0.7; f_bin 0.12 / 0.28 / 0.70 at P_min 1.4 d), clusters of 12 stars with the M17 "
                                     "error vector, ddof 1, >= %.0e binary orbits per variant and GPU (weak scaling: ranks "
                                     "take consecutive realization ranges), fused infinite-pool kernel; data_sana medians are "
                                     "a sanity band only (SURVEY Q16: the reference's current code does not reproduce them)"
                                     % target,
                         "variants": c3, "binary_rvs_per_s": world * c3_total_rv / (c3_total_ms * 1e-3),
                         "unit": "binary RV/s", "timing": "CUDA events, mean of 2 launches after 1 warm-up, max over ranks",
                         # the roofline block uses the f_bin 0.7 variants: the executed count of the ncu capture is
                         # per binary RV AT f_bin 0.7 (at lower fractions the single-star slots weigh more per binary)
                         "_units_per_s_rank": sum(v["binary_rvs"] for v in c3 if v["fbin"] == 0.7) /
                                              (1e-3 * sum(v["ms"] for v in c3 if v["fbin"] == 0.7))}

        # ---- C4: the ten data_RT20 clusters, weighted dispersion, 10^6 realizations each --------------------
        n_real4 = int(args.c4_realizations)
        c4, c4_rv = [], 0
        bufs = []
        for name, sig_obs, n, m0, m1 in RT20:
            err = M17 if name == "M17" else np.ones(n) * 1.2                     # findBestDistributions.py:291-294
            bufs.append((name, sig_obs, n, _abi.PopParams.defaults(min_mass=float(m0), max_mass=float(m1)),
                         dv.to_device(err, "float32", dev), dv.empty((n_real4,), "float32", dev),
                         dv.zeros((4,), "int64", dev)))

        def run_c4():
            for ci, (name, sig_obs, n, pop, errd, out, c4cnt) in enumerate(bufs):
                _abi.check(lib.rvmc_sigma1d_fused(C.byref(pop), 100 + ci, dv.ptr(errd), n, 1, 0, rank * n_real4, n_real4,
                                                  dv.ptr(out), dv.ptr(c4cnt), st))
        ms4 = event_ms(run_c4, 3, 1)
        launches += 4 * len(bufs)
        for name, sig_obs, n, pop, errd, out, c4cnt in bufs:
            n_rv = int(dv.to_host(c4cnt)[0]) // 4
            c4_rv += n_rv
            c4.append({"cluster": name, "n_stars": n, "mass_range": [pop.min_mass, pop.max_mass], "binary_rvs": n_rv,
                       "median_sigma_1d": float(torch.median(out).item()), "observed_sigma_1d": sig_obs})
        configs["C4"] = {"workload": "BASELINE configs[3]: the ten clusters of data_RT20/Nstars_massRange.txt (N = 5..332 "
                                     "stars, their mass ranges, f_bin 0.7, Sana+2012 periods), weighted dispersion "
                                     "(RVMC.py:262-268), errors 1.2 km/s (M17: its own vector), %.0e realizations each per "
                                     "GPU, fused infinite-pool kernel; observed_sigma_1d is the table's `sigma` column"
                                     % n_real4,
                         "clusters": c4, "ms": ms4, "binary_rvs_per_s": world * c4_rv / (ms4 * 1e-3), "unit": "binary RV/s",
                         "star_draws_per_s": world * sum(b[2] for b in bufs) * n_real4 / (ms4 * 1e-3),
                         "timing": "CUDA events around the ten launches, mean of 3 after 1 warm-up, max over ranks",
                         "_units_per_s_rank": c4_rv / (ms4 * 1e-3)}
        del bufs, run_c4

        # ---- C1: RVMC_example.py shape -- 10^5-star pool, 100 stars x 10^4 realizations ------------------------
        pop1 = _abi.PopParams.defaults()
        d_e100 = dv.to_device(np.full(100, 1.2), "float32", dev)
        out1 = dv.empty((10 ** 4,), "float32", dev)
        c1cnt = dv.zeros((4,), "int64", dev)

        def run_c1():
            _abi.check(lib.rvmc_sigma1d_fused(C.byref(pop1), 1234, dv.ptr(d_e100), 100, 0, 1, rank * 10 ** 4, 10 ** 4,
                                              dv.ptr(out1), dv.ptr(c1cnt), st))
        ms1 = event_ms(run_c1, 20, 3)
        launches += 23
        n_rv1 = int(dv.to_host(c1cnt)[0]) // 23

        def c1_class(k):
            c = RVMC(number_of_stars=10 ** 5, seed=77 + k + 1000 * rank, device=dev)
            c.calculate_binary_velocities(materialize=False)
            return c.subsample(sample_size=100, measured_errors=1.2, number_of_samples=10 ** 4)
        for k in range(2):
            s1 = c1_class(-1 - k)
        barrier()
        t0 = time.perf_counter()
        for k in range(5):
            s1 = c1_class(k)
        barrier()
        c1_s = max_over_ranks((time.perf_counter() - t0) / 5)
        configs["C1"] = {"workload": "BASELINE configs[0] (RVMC_example.py:6-37): Sana+2012 defaults, f_bin 0.7, clusters of "
                                     "100 stars, 10^4 realizations, measured_errors 1.2",
                         "fused": {"what": "infinite-pool fused kernel (rvmc_sigma1d_fused), one launch", "ms": ms1,
                                   "binary_rvs": n_rv1, "binary_rvs_per_s": world * n_rv1 / (ms1 * 1e-3),
                                   "median_sigma_1d": float(torch.median(out1).item()),
                                   "note": "250 blocks of work for 444 resident: launch latency, not throughput"},
                         "class_finite_pool": {"what": "RVMC(number_of_stars=1e5); calculate_binary_velocities(); "
                                                       "subsample(100, 1.2, 1e4) -> sigma_1d on the host (10^5 pool orbits in "
                                                       "fp64 + mixing + bootstrap), wall clock per pass",
                                               "seconds": c1_s, "pool_binary_rvs_per_s": world * 1e5 / c1_s,
                                               "median_sigma_1d": float(np.median(s1))}}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        launch_ms = float(np.mean(per_launch_ms))
        rv_per_s_kernel = units_per_step / (launch_ms * 1e-3)
        pipes = {}
        if not args.no_pipe_peaks:
            for kind, name in ((0, "ffma"), (1, "mufu"), (2, "imad_wide"), (3, "dfma"), (4, "issue_mix")):
                ops, ms = C.c_double(), C.c_double()
                _abi.check(lib.rvmc_pipe_peak(kind, 2000, C.byref(ops), C.byref(ms)))
                pipes[name] = ops.value
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at the default size from the committed
        # ncu --set full capture (the results minus what still sits in the 126 MB L2 when the kernel ends)
        traffic, traffic_src = args.traffic_bytes, "--traffic-bytes"
        if traffic is None and (nstars, ns) == (100, 10 ** 6):
            traffic = EXEC_MULTI.get("traffic_bytes_packed_flags" if det is None else "traffic_bytes",
                                     EXEC_MULTI.get("traffic_bytes"))
            traffic_src = EXEC_MULTI["source"]
        sms = C.c_int()
        _abi.check(lib.rvmc_device_info(local, C.byref(sms), None, None, None))
        clk_hz = 1e6 * (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        roof = issue_roofline(rv_per_s_kernel, ALG_MULTI, EXEC_MULTI, sms.value, clk_hz,
                              "rvmc::biascorr_fused_kernel<6, FAST, QUIRK, !RVOUT, NOISY>", out_bytes / units_per_step,
                              peaks, peak_src, pipes, launch_ms, traffic, traffic_src)
        if pipes:
            roof["pipe_peaks_lane_ops_per_s"] = pipes
            roof["pipe_peak_source"] = ("rvmc_pipe_peak microbenchmarks (FFMA, MUFU.EX2, IMAD.WIDE, DFMA, mixed issue) run "
                                        "on this device in this process; MEASURED_PEAKS.json has no FP32/SFU figure")
        for key in ("C3", "C4"):
            if key in configs:
                ups = configs[key].pop("_units_per_s_rank")
                configs[key]["roofline"] = issue_roofline(
                    ups, ALG_SINGLE, EXEC_SINGLE, sms.value, clk_hz, "rvmc::sigma1d_fused_kernel<SINGLE, KEYED, 1, !GUARD>",
                    4.0 / (12 * 0.7), peaks, peak_src, pipes)
                configs[key]["roofline"]["note"] = (
                    "a unit is one binary RV; the kernel also draws the coin + noise of every SINGLE star slot, so the "
                    "executed count per binary RV grows as f_bin falls (the ncu figure is for S = 12, f_bin 0.7; the C3 "
                    "block is computed from the four f_bin 0.7 variants)")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 period/phase reduction)",
                "data": "synthetic", "config": workload_config(world, nstars, ns, out_bytes), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "api": "BiasCorr(...).calculate_radial_velocities(); get_detected_binaries()",
                        "flag_transfer": {"bytes": "bytes (one uint8 per binary by DMA)",
                                          "bits": "bits (bit-packed by the kernel, widened to NumPy bool on the host by "
                                                  "rvmc_unpack_bits_host, %d threads)" % dv.host_threads(),
                                          "hybrid": "hybrid (%.0f %% of the flags as bytes by DMA while %d host threads "
                                                    "widen the bit-packed rest)" % (100 * share, dv.host_threads())
                                          }[args.flag_transfer],
                        "detected_fraction": frac_det,
                        "host_phases_ms_per_step_rank0": e2e_phases,
                        "with_delta_rv": {"value": e2e_drv_value, "unit": UNIT, "steps": e2e_drv_steps,
                                          "d2h_bytes_per_step": d2h + nstars * ns * 4,
                                          "api": "... plus reading bc.delta_rv (float32, 4 B per binary)"}},
                "gpu_launches": launches, "roofline": roof, "grid": grid_info, "grid_independent": grid_indep,
                "configs": configs,
                "counters": {"binary_epoch_rvs": int(counters[0]), "detected": int(counters[1]),
                             "fp64_guard_lanes": int(counters[2]), "nonconverged": int(counters[3])}}
        if world == 1 and not args.no_cpu_baseline:
            import multiprocessing as mp
            use_ref, kind, how = cpu_kind()
            cores = host_cores()
            with mp.get_context("fork").Pool(cores) as pool:
                cpu_pass(pool, cores, 1, 20000, 5, use_ref)                  # warm-up (imports, page faults)
                n, dt = 0, 0.0
                for k in range(4):
                    a, b = cpu_pass(pool, cores, 2, 50000, 11 + k, use_ref)
                    n += a
                    dt += b
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": "4 passes of %d cores x 2 stars x 50000 samples x 6 epochs; %s" % (cores, how)}
            c1_ref = cpu_c1_seconds()
            if c1_ref is not None and "C1" in configs:
                configs["C1"]["reference_cpu_1core_seconds"] = c1_ref
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--stars", type=int, default=100)
    ap.add_argument("--samples", type=int, default=10 ** 6)
    ap.add_argument("--flag-transfer", default=os.environ.get("RVMC_FLAG_TRANSFER", "auto"),
                    choices=["auto", "bits", "bytes", "hybrid"],
                    help="how the detection flags reach the host: auto (default: bytes for one rank per host, bit-packed + host "
                         "unpack when ranks share the host), bits, bytes, or both routes at once")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipe-peaks", action="store_true")
    ap.add_argument("--no-grid", action="store_true", help="skip the grid-points/s legs (C5)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1 / C3 / C4 legs")
    ap.add_argument("--c3-orbits", type=float, default=1e9, help="binary orbits per C3 variant and GPU")
    ap.add_argument("--c4-realizations", type=float, default=1e6, help="realizations per C4 cluster and GPU")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
