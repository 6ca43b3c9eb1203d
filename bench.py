#!/usr/bin/env python
"""Headline benchmark: binary-epoch RVs per second on the multi-epoch detection workload
(BASELINE.json configs[1]: 10^8 binaries x 6 epochs, Delta-RV > 20 km/s at 4 sigma).

    python bench.py --gpus 1 --steps K --warmup W                 # this framework (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference algorithm on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU, weak scaling

A step = one pass of the hot path (rvmc_biascorr_fused: sample + Kepler + RV at 6 epochs + noise +
Delta-RV + significance test) over 10^8 binaries per GPU, writing Delta-RV (fp32) and the detection
flag (u8) for every binary to HBM.  `value` is timed with CUDA events on the launching stream with
everything resident on the device; `e2e` is the same metric through the drop-in Python class
(`BiasCorr(...)` -> `calculate_radial_velocities()` -> `get_detected_binaries()`) with host inputs and
the detection array copied back to host memory inside the timed region.

Only this file's `cpu_baseline` leg and `--impl reference` import `oracle/` (the CPU restatement
of the reference algorithm); the product path never does.
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

# algorithmic work per binary-epoch RV on the multi-epoch path
ALG_FP32_FLOP = 100.0
ALG_SFU_OPS = 17.0
ALG_INT_OPS = 55.0
ALG_FP64_FLOP = 6.0
ALG_HBM_BYTES = 5.0 / 6.0      # 4 B Delta-RV + 1 B flag per binary of 6 epochs


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ---------------------------------------------------------------------------------------------
def _cpu_chunk(args):
    """One bounded sample of the workload on one core: `nstars` stars x `ns` samples x 6 epochs of
    RVMC_bias_corr.py:175-243,406-422,451-476 (NumPy + SciPy-style vectorised secant)."""
    seed, nstars, ns = args
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


def cpu_pass(pool, cores, nstars, ns, seed0):
    t0 = time.perf_counter()
    res = pool.map(_cpu_chunk, [(seed0 + c, nstars, ns) for c in range(cores)])
    dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU algorithm on all host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import multiprocessing as mp
    cores = host_cores()
    nstars, ns = 2, 50000        # per core and step: 6e5 binary-epoch RVs (~0.6 s at ~1e6 RV/s/core)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for w in range(args.warmup):
            cpu_pass(pool, cores, nstars, ns, 1000 * w)
        total, elapsed = 0, 0.0
        for k in range(args.steps):
            n, dt = cpu_pass(pool, cores, nstars, ns, 77 + 1000 * k)
            total += n
            elapsed += dt
    value = total / elapsed
    sample = "%d cores x %d stars x %d samples x 6 epochs per step (oracle port: NumPy + SciPy-style secant)" % (
        cores, nstars, ns)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def workload_config(n_gpus, nstars=100, ns=10 ** 6):
    return {"workload": "BiasCorr multi-epoch detection (BASELINE configs[1]): %d stars x %d samples = %.0e binaries "
                        "per GPU x 6 epochs [0,1,7,30,180,365] d, IMF masses 6-20, Sana+2012 defaults, rv_errors 2.0 "
                        "km/s, Delta-RV > 20 km/s at 4 sigma" % (nstars, ns, nstars * ns),
            "binaries_per_gpu": nstars * ns, "epochs": 6, "parallelism": "dp%d (samples sharded, no collective)" % n_gpus,
            "l2": "inputs are < 8 KB of tables; outputs (%.0f MB/step) exceed the 126 MB L2" % (nstars * ns * 5 / 1e6)}


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


def run_gpu_arm(args):
    import ctypes as C

    import torch
    import torch.distributed as dist

    from rvmc_b200 import _abi
    from rvmc_b200 import _device as dv
    from rvmc_b200.bias_corr import BiasCorr

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

    d_nep = dv.to_device(np.full(nstars, 6, np.int32), "int32", dev)
    d_tobs = dv.to_device(np.tile(DAYS6, (nstars, 1)), "float64", dev)
    d_err = dv.to_device(np.full((nstars, 6), 2.0, np.float32), "float32", dev)
    drv = dv.empty((nstars, ns), "float32", dev)
    # the outputs of the product path (BiasCorr.calculate_radial_velocities): Delta-RV per binary and the
    # detection flags, bit-packed (or one byte each with --flag-transfer bytes)
    nw = (ns + 31) // 32
    det = dv.empty((nstars, ns), "uint8", dev) if args.flag_transfer == "bytes" else None
    bits = dv.empty((nstars, nw), "int32", dev) if args.flag_transfer == "bits" else None
    cnt = dv.zeros((4,), "int64", dev)
    BiasCorr.flag_transfer = args.flag_transfer

    def step():
        _abi.check(lib.rvmc_biascorr_fused_packed(C.byref(p), seed, noise_seed, 0, nstars, 6, dv.ptr(d_nep),
                                                  dv.ptr(d_tobs), dv.ptr(d_err), 0, None, None, 1, 20.0, 4.0,
                                                  first_sample, ns, ns, dv.ptr(drv), dv.ptr(det), dv.ptr(bits), nw,
                                                  None, dv.ptr(cnt), None, st))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

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
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * units_per_step * args.steps / (total_ms_max * 1e-3)

    # ---- e2e: the drop-in class with host inputs and the detection array back on the host ----------
    t_obs_host = [DAYS6.copy() for _ in range(nstars)]
    h2d = nstars * 6 * 8 + nstars * 6 * 4 + nstars * 4
    d2h = nstars * ns if args.flag_transfer == "bytes" else nstars * nw * 4
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step(k):
        bc = BiasCorr(t_obs_host, number_of_samples=ns, rv_errors=2.0, seed=1000 + k + 7919 * rank, device=dev)
        bc.calculate_radial_velocities(delta_RV=20, n_sigma_RV=4)
        d = bc.get_detected_binaries(delta_RV=20, n_sigma_RV=4)
        return d

    for k in range(2):
        d = e2e_step(-1 - k)
    frac = float(d.mean())
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        d = e2e_step(k)
        assert d.shape == (nstars, ns) and d.dtype == np.bool_
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * units_per_step * e2e_steps / float(t.item())

    # ---- second headline of BASELINE.json: grid points per second (configs[4]) ----------------------
    grid_info = None
    if not args.no_grid:
        from rvmc_b200 import grid as rgrid
        axes = dict(period_pi=np.linspace(-0.95, 0.5, 64), mass_ratio_kappa=np.linspace(-0.95, 1.0, 64),
                    fbin=np.linspace(0.0625, 1, 16))
        rgrid.grid_search({k: v[:4] for k, v in axes.items()}, number_of_samples=1000, seed=1, device=dev)
        g_times = []
        for rep in range(3):
            barrier()
            t0 = time.perf_counter()
            gres = rgrid.grid_search(axes, number_of_samples=10 ** 5, seed=2024, observed_sigma=25.0, device=dev)
            barrier()
            g_times.append(time.perf_counter() - t0)
        tg = torch.tensor([min(g_times)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        npts = int(np.prod(gres["shape"]))
        grid_info = {"workload": "findBestDistributions grid (BASELINE configs[4]): 64 x 64 x 16 (pi, kappa, fbin) x 10^5 "
                                 "realizations x 12 stars (M17 errors), statistics + best fit; points dealt round-robin "
                                 "to the ranks, 80-byte statistics all-gathered; f_bin axis evaluated in one pass",
                     "points": npts, "seconds": float(tg.item()), "points_per_s": npts / float(tg.item()),
                     "scaling": "strong", "timing": "host wall clock, barrier + synchronize on both sides, best of 3, "
                                                    "max over ranks; includes building the grid and the host copy of "
                                                    "the statistics",
                     "best_index": list(gres["best_index"])}
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
        # Roofline of the dominant kernel.  The path has no dense contraction and writes < 1 B per unit,
        # so neither the tensor cores nor HBM bound it: the SM instruction pipes do.
        # achieved = ALGORITHMIC work per binary-epoch RV (100 FP32 flop + 17 SFU ops +
        # 55 INT ops + 6 FP64 flop = 178 lane-instructions) x units/s of the CUDA-event timed launches;
        # the binding ceiling is the issue rate (1 warp-instruction / clk / SM sub-partition), below the
        # SFU (17 ops) and FP32 (100 flop) ceilings.  peak = SMs x 4 x 32 x SM clock sampled under load.
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at the default size (ncu --set full,
        # profiles/r01_biascorr_fused_fullsize_ncu.md): 0.13 MB + 442.49 MB; the algorithmic 500 MB of
        # results minus what still sits in the 126 MB L2 when the kernel ends
        traffic, traffic_src = args.traffic_bytes, "--traffic-bytes"
        if traffic is None and (nstars, ns) == (100, 10 ** 6):
            traffic, traffic_src = 442.62e6, "profiles/r01_biascorr_fused_fullsize_ncu.md"
        sms = C.c_int()
        _abi.check(lib.rvmc_device_info(local, C.byref(sms), None, None, None))
        clk_hz = 1e6 * (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0)
        alg_instr = ALG_FP32_FLOP + ALG_SFU_OPS + ALG_INT_OPS + ALG_FP64_FLOP
        issue_peak = sms.value * 4 * 32 * clk_hz
        roof = {"bound": "issue (instruction pipes; not hbm, not tensor)",
                "achieved": rv_per_s_kernel * alg_instr / 1e12, "peak": issue_peak / 1e12,
                "unit": "T lane-instr/s", "frac": rv_per_s_kernel * alg_instr / issue_peak,
                "peak_source": "%d SMs x 4 schedulers x 32 lanes x %.0f MHz (NVML, sampled during the timed region)"
                               % (sms.value, clk_hz / 1e6),
                "traffic": traffic, "traffic_source": traffic_src, "kernel": "rvmc::biascorr_fused_kernel<6, FAST, QUIRK, !RVOUT, NOISY>", "launch_ms": launch_ms,
                "algorithmic_per_unit": {"fp32_flop": ALG_FP32_FLOP, "sfu_ops": ALG_SFU_OPS, "int_ops": ALG_INT_OPS,
                                         "fp64_flop": ALG_FP64_FLOP, "lane_instr": alg_instr, "hbm_bytes": ALG_HBM_BYTES},
                "ncu": "profiles/r01_biascorr_fused_fullsize_ncu.md (same launch, ncu --set full): 77 % of issue slots "
                       "busy; 142 EXECUTED thread-instructions per unit -- fewer than the 178 algorithmic ones of SURVEY "
                       "8d, because sin/cos of the eccentric anomaly, the cubic starter and the noise normals are "
                       "evaluated more cheaply than the survey's operation count assumed, so frac (algorithmic work / "
                       "issue ceiling) exceeds the executed issue utilisation; FMA pipe 43 %, ALU 36 %, XU 54 %, "
                       "FP64 7 %, tensor 0 %, DRAM throughput < 2 %",
                "issue_slots_busy_ncu": 0.776, "executed_lane_instr_per_unit_ncu": 142.4,
                "hbm": {"achieved": rv_per_s_kernel * ALG_HBM_BYTES / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                        "peak_source": "MEASURED_PEAKS.json" if peak_src == "measured" else "fallback 6650 GB/s",
                        "frac": rv_per_s_kernel * ALG_HBM_BYTES / 1e9 / peaks.get("hbm_gbs", 6650.0)}}
        if pipes:
            # the individual pipes, against rates measured live on this device by rvmc_pipe_peak
            roof["fp32"] = {"achieved": rv_per_s_kernel * ALG_FP32_FLOP / 1e12, "peak": 2.0 * pipes["ffma"] / 1e12,
                            "unit": "TFLOP/s", "frac": rv_per_s_kernel * ALG_FP32_FLOP / (2.0 * pipes["ffma"])}
            roof["sfu"] = {"achieved": rv_per_s_kernel * ALG_SFU_OPS / 1e12, "peak": pipes["mufu"] / 1e12,
                           "unit": "T MUFU-op/s", "frac": rv_per_s_kernel * ALG_SFU_OPS / pipes["mufu"]}
            roof["fp64"] = {"achieved": rv_per_s_kernel * ALG_FP64_FLOP / 1e12, "peak": 2.0 * pipes["dfma"] / 1e12,
                            "unit": "TFLOP/s", "frac": rv_per_s_kernel * ALG_FP64_FLOP / (2.0 * pipes["dfma"])}
            roof["pipe_peaks_lane_ops_per_s"] = pipes
            roof["pipe_peak_source"] = ("rvmc_pipe_peak microbenchmarks (FFMA, MUFU.EX2, IMAD.WIDE, DFMA, mixed issue) run "
                                        "on this device in this process; MEASURED_PEAKS.json has no FP32/SFU figure")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 (f64 period/phase reduction)",
                "data": "synthetic", "config": workload_config(world, nstars, ns), "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "api": "BiasCorr(...).calculate_radial_velocities(); get_detected_binaries()",
                        "detected_fraction": frac},
                "gpu_launches": args.steps, "roofline": roof, "grid": grid_info,
                "counters": {"binary_epoch_rvs": int(counters[0]), "detected": int(counters[1]),
                             "fp64_guard_lanes": int(counters[2]), "nonconverged": int(counters[3])}}
        if world == 1 and not args.no_cpu_baseline:
            import multiprocessing as mp
            cores = host_cores()
            with mp.get_context("fork").Pool(cores) as pool:
                cpu_pass(pool, cores, 1, 20000, 5)                  # warm-up (imports, page faults)
                n, dt = 0, 0.0
                for k in range(4):
                    a, b = cpu_pass(pool, cores, 2, 50000, 11 + k)
                    n += a
                    dt += b
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "4 passes of %d cores x 2 stars x 50000 samples x 6 epochs (oracle port of "
                                              "RVMC_bias_corr.py: NumPy + SciPy-style vectorised secant)" % cores}
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
    ap.add_argument("--flag-transfer", default=os.environ.get("RVMC_FLAG_TRANSFER", "bits"), choices=["bits", "bytes"],
                    help="how the detection flags reach the host: bit-packed + host unpack (default) or uint8")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipe-peaks", action="store_true")
    ap.add_argument("--no-grid", action="store_true", help="skip the secondary grid-points/s measurement")
    ap.add_argument("--traffic-bytes", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/); "
                         "default: the full-size capture profiles/r01_biascorr_fused_fullsize_ncu.md")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
