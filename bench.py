#!/usr/bin/env python
"""bench.py -- batched sim-steps/s of the hydrodynamic force path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload = "rm3_irregular_ensemble"): SURVEY.md section 8(d) -- RM3-shaped two-body design
(D = 12, L = 1001 radiation lags over 60 s, dt = 0.01 s, excitation IRF +-30 s -> 6000 lags), JONSWAP
Hs 2.5 m / Tp 8 s / gamma 3.3, 1000 spectrum components, one random-phase realisation per instance
(seed 1 + global instance index), 16384 instances PER GPU (weak scaling: instances are independent, tables are
replicated, no data-path collective).  A "step" is one lock-step force evaluation of every instance:
hydrostatics + radiation convolution + excitation convolution + total.  The velocity-history window is
pre-filled (>= 6000 steps) before anything is timed, so the timed steps are steady state.

value : instance-steps/s with the step's pose/velocity already resident in HBM (hc_step_device), timed with CUDA
        events on the ensemble's stream, max over ranks.
e2e   : the same through the host-buffer C-ABI call (hc_step): pinned host pose/velocity -> H2D -> kernels -> D2H
        force, every step, wall clock around K synchronous calls, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# headline workload (BASELINE.json metric) -- module-level names are rebound by select_workload()
DT = 0.01
DOFS = 12
RIRF_STEPS = 1001
EXC_STEPS = 6000
SEA = dict(Hs=2.5, Tp=8.0, gamma=3.3, fmin=0.001, fmax=1.0, nfreq=1000, ramp=20.0)
SNAP = 1e-8
GVEC = (0.0, 0.0, -9.81)
WORKLOAD = "rm3_irregular_ensemble"
PREFILL = 6010


def workload_tables():
    from hydrochrono_b200 import synth
    if WORKLOAD == "sphere_irregular_ensemble":
        z = np.load(os.path.join(ROOT, "tests", "golden", "sphere_tables.npz"))
        body = {k: z[k] for k in ("cg", "cb", "lin_matrix", "inf_added_mass", "rirf_K", "rirf_t", "exc_mag", "exc_phase",
                                  "exc_irf_f", "exc_irf_t")}
        body["disp_vol"] = float(z["disp_vol"])
        return {"rho": float(z["rho"]), "g": float(z["g"]), "water_depth": float(z["water_depth"]), "w": z["w"],
                "bodies": [body]}
    return synth.rm3_like()


def select_workload(name):
    """Secondary workload on the reference's real tables: sphere x 16384, N = 1, D = 6, L = 1001 lags
    with rirf spacing == dt = 0.015, excitation IRF resampled to 8334 lags, Hs 2 / Tp 12 / gamma 1."""
    global DT, DOFS, RIRF_STEPS, EXC_STEPS, SEA, WORKLOAD, PREFILL
    WORKLOAD = name
    if name == "sphere_irregular_ensemble":
        DT, DOFS, RIRF_STEPS, EXC_STEPS, PREFILL = 0.015, 6, 1001, 8334, 1010
        SEA = dict(Hs=2.0, Tp=12.0, gamma=1.0, fmin=0.001, fmax=1.0, nfreq=1000, ramp=60.0)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="instances per GPU")
    ap.add_argument("--prefill", type=int, default=-1, help="untimed steps to fill the history window (-1 = auto)")
    ap.add_argument("--faithful", action="store_true", help="bracket_snap = 0 (bit-faithful bracketing)")
    ap.add_argument("--rad-chunk", type=int, default=0)
    ap.add_argument("--exc-chunk", type=int, default=0)
    ap.add_argument("--cpu-instances", type=int, default=32)
    ap.add_argument("--cpu-steps", type=int, default=600)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-lookahead", action="store_true", help="per-step excitation kernel only")
    ap.add_argument("--lookahead-mode", type=int, default=0, help="0 auto (= 5), 2 in-stream, 3 background, 4 in-stream DMMA, 5 background DMMA")
    ap.add_argument("--rad-lookahead", type=int, default=0, help="radiation look-ahead block (12 DoF): 0 auto (on, background), 1 off, 2 on (background), 3 on (in-stream)")
    ap.add_argument("--rad-kernel", type=int, default=0, help="0 auto (= 1), 1 FP64 FMA pipe, 2 FP64 tensor cores (DMMA, 12 DoF only), 3 tensor cores (rows 0-7) + FMA pipe (rows 8-11)")
    ap.add_argument("--workload", default="rm3_irregular_ensemble",
                    choices=["rm3_irregular_ensemble", "sphere_irregular_ensemble"])
    return ap.parse_args()


def motion(amp, om, t, nb, offset=0):
    ph = 0.01 * (offset + np.arange(nb))[:, None]
    return amp * np.sin(om * t + ph), amp * om * np.cos(om * t + ph)


class ClockSampler:
    """SM clock and throttle reasons of ONE GPU, sampled during the timed region by a thread of the rank that drives
    it (in-process NVML, ~0.1 ms per sample every 20 ms).  A single `nvidia-smi -lms` process polling all the GPUs of
    an 8-rank job stalls kernel launches of every rank while it walks the devices (measured: 60 M instead of 70 M
    instance-steps/s per GPU at N = 8), so nvidia-smi is only the fallback when the NVML binding is missing."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th, self.nvml, self.proc = index, [], False, None, None, None
        self.period = float(os.environ.get("HC_BENCH_CLOCK_PERIOD", "0.02"))

    def start(self):
        if os.environ.get("HC_BENCH_NO_CLOCKS"):       # experiment switch: no sampling at all
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            try:      # the CUDA ordinal need not be the NVML index: match by UUID when torch exposes it
                import torch
                self.h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(self.index).uuid))
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.nvml = None
            try:
                q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                     "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                              "--format=csv,noheader,nounits", "-lms", "100"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.th = threading.Thread(target=self._read, daemon=True)
                self.th.start()
            except Exception:
                self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [n.nvmlClocksThrottleReasonHwSlowdown, n.nvmlClocksThrottleReasonHwThermalSlowdown,
                n.nvmlClocksThrottleReasonSwThermalSlowdown, n.nvmlClocksThrottleReasonSwPowerCap]
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((sm, self.mx, [1.0 if (r & b) else 0.0 for b in bits]))
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            try:
                self.samples.append((float(p[0]), float(p[1]), [1.0 if v.lower().startswith("active") else 0.0 for v in p[2:6]]))
            except (ValueError, IndexError):
                continue

    def stop(self):
        """[median sm, min sm, max clock, 4 reason flags, samples] of this GPU (None entries when unavailable)."""
        self.stop_flag = True
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
        if self.th:
            self.th.join(timeout=2)
        if not self.samples:
            return None
        sm = [x[0] for x in self.samples]
        flags = [max(x[2][i] for x in self.samples) for i in range(4)]
        return [float(np.median(sm)), float(min(sm)), float(max(x[1] for x in self.samples))] + flags + [float(len(sm))]


def hbm_peak():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the radiation kernel from the committed ncu --set full capture, if any."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return d
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (line-faithful port of the reference's CPU path) on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_arm(n_inst, n_steps, prefill, quiet=False):
    from hydrochrono_b200 import synth
    from oracle import hc_oracle as orc
    raw = workload_tables()
    T = orc.Tables(raw)
    cores = orc.num_threads()
    amp, om = synth.prescribed_motion(DOFS)
    duration = (prefill + 2 * n_steps + 64) * DT
    insts = []
    for i in range(n_inst):
        inst = orc.Instance(T, omp_mode=0)
        inst.set_irregular(dt=DT, duration=duration, ramp=SEA["ramp"], Hs=SEA["Hs"], Tp=SEA["Tp"], fmin=SEA["fmin"],
                           fmax=SEA["fmax"], nfreq=SEA["nfreq"], gamma=SEA["gamma"], seed=1 + i,
                           share_irf_from=insts[0] if insts else None)
        insts.append(inst)
    # untimed: fill the velocity-history window (instances across cores)
    orc.bench_steps(insts, prefill, 0.0, DT, 1, amp, om, GVEC)
    t0 = prefill * DT
    sec_best, _ = orc.bench_steps(insts, n_steps, t0, DT, 1, amp, om, GVEC)
    best = n_inst * n_steps / sec_best
    # reference-style threading: instances one after another, OpenMP across radiation lags (hydro_forces.cpp:593-647)
    n_ref = max(1, min(n_inst, 4))
    ref_steps = max(10, n_steps // 4)
    sec_ref, _ = orc.bench_steps(insts[:n_ref], ref_steps, t0 + n_steps * DT, DT, 0, amp, om, GVEC)
    ref_style = n_ref * ref_steps / sec_ref
    return {"value": best, "unit": "instance-steps/s", "cores": cores, "kind": "port",
            "sample": "%d instances x %d steady-state steps after a %d-step history prefill, one instance per thread "
                      "(best effort); reference-style threading (OpenMP across lags, instances serial): %.1f "
                      "instance-steps/s on %d instances x %d steps" % (n_inst, n_steps, prefill, ref_style, n_ref,
                                                                       ref_steps),
            "reference_style_value": ref_style}


def run_reference(args, rank, world):
    if rank != 0:
        return
    prefill = args.prefill if args.prefill >= 0 else PREFILL
    t0 = time.time()
    res = None
    vals = []
    for _ in range(max(1, min(args.steps, 3)) if args.steps < 10 else 1):
        res = cpu_arm(args.cpu_instances, args.cpu_steps, prefill)
        vals.append(res["value"])
    v = float(np.median(vals))
    line = {
        "metric": "batched sim-steps/sec (RM3 irregular ensemble)", "value": v, "unit": "instance-steps/s",
        "impl": "reference", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * args.batch / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "instances_per_gpu": args.batch, "dofs": DOFS,
                   "rirf_steps": RIRF_STEPS, "exc_irf_steps": EXC_STEPS, "dt": DT, "sea_state": SEA,
                   "note": "reference CPU path (oracle port; the reference itself needs Chrono/Eigen/HDF5 and cannot "
                           "be built here) on a bounded sample of the same workload; ms_per_step is the time the CPU "
                           "would need for one lock-step of all instances_per_gpu instances"},
        "cpu_baseline": res,
        "e2e": {"value": v, "unit": "instance-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.time() - t0,
    }
    line["cpu_baseline"]["value"] = v
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def main():
    args = parse()
    select_workload(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import hydrochrono_b200 as hc
    from hydrochrono_b200 import synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hydro force path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    K, W = args.steps, max(args.warmup, 3)
    prefill = args.prefill if args.prefill >= 0 else PREFILL
    snap = 0.0 if args.faithful else SNAP
    total_steps = prefill + 2 * (W + K) + 128
    raw = workload_tables()
    T = hc.Tables.from_raw(raw)
    # the ensemble launches on this stream and the events are recorded on it; high priority so that the per-step
    # kernels outrank the look-ahead passes the library runs on its low-priority side streams
    stream = torch.cuda.Stream(device=dev, priority=-1)
    ens = hc.Ensemble(T, batch=B, device=local_rank, dt_hint=DT, bracket_snap=snap, rad_chunk=args.rad_chunk,
                      exc_chunk=args.exc_chunk, use_graph=not args.no_graph, stream=stream.cuda_stream,
                      exc_lookahead=1 if args.no_lookahead else args.lookahead_mode, rad_kernel=args.rad_kernel, rad_lookahead=args.rad_lookahead)
    from hydrochrono_b200 import shard
    # weak scaling: every rank owns a contiguous block of B instances of the (world * B)-instance ensemble
    lo, hi = shard.shard_range(world * B, world, rank)
    seeds = shard.instance_seeds(lo, hi)
    t_setup = time.time()
    ens.set_waves_irregular(dt=DT, duration=total_steps * DT, ramp=SEA["ramp"], Hs=SEA["Hs"], Tp=SEA["Tp"],
                            fmin=SEA["fmin"], fmax=SEA["fmax"], nfreq=SEA["nfreq"], gamma=SEA["gamma"], seeds=seeds)
    eta_s = ens.profile()["eta_synthesis_seconds"]
    nf, n_eta, le = ens.irregular_sizes()
    assert le == [EXC_STEPS] * (DOFS // 6), le

    amp, om = synth.prescribed_motion(DOFS)
    NBUF = 8
    h_pose = [torch.empty((B, DOFS), dtype=torch.float64).pin_memory() for _ in range(NBUF)]
    h_vel = [torch.empty((B, DOFS), dtype=torch.float64).pin_memory() for _ in range(NBUF)]
    h_force = torch.empty((B, DOFS), dtype=torch.float64).pin_memory()
    for i in range(NBUF):
        p, v = motion(amp, om, i * DT, B, rank * B)
        h_pose[i].copy_(torch.from_numpy(p))
        h_vel[i].copy_(torch.from_numpy(v))
    d_pose = [x.to(dev) for x in h_pose]
    d_vel = [x.to(dev) for x in h_vel]
    d_force = torch.empty((B, DOFS), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()

    step_no = [0]
    t_now = [0.0]     # advanced as Chrono advances ChTime: t += dt

    def dev_step():
        n = step_no[0]
        ens.step_device(t_now[0], d_pose[n % NBUF], d_vel[n % NBUF], d_force, GVEC)
        step_no[0] = n + 1
        t_now[0] += DT

    def host_step():
        n = step_no[0]
        ens.step(t_now[0], h_pose[n % NBUF].numpy(), h_vel[n % NBUF].numpy(), GVEC, out=h_force.numpy())
        step_no[0] = n + 1
        t_now[0] += DT

    # ---- untimed: fill the radiation history window -----------------------------------------------
    for _ in range(prefill):
        dev_step()
    ens.sync()
    assert ens.history_len() >= min(prefill, RIRF_STEPS - 1), ens.history_len()
    hist_rows = ens.history_len()
    setup_s = time.time() - t_setup

    # ---- device-resident leg (value) -----------------------------------------------------------------
    for _ in range(W):
        dev_step()
    ens.sync()
    sampler = ClockSampler(local_rank)              # every rank samples the GPU it drives
    launches0 = ens.profile()["kernel_launches"]
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(stream)                              # CUDA events on the launching stream
    for _ in range(K):
        dev_step()
    ev1.record(stream)
    t_enq = time.perf_counter() - t0               # host time to enqueue the K steps (the GPU runs behind)
    ens.sync()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    t_dev = ev0.elapsed_time(ev1) * 1e-3
    barrier()
    mine = sampler.stop()
    # over the ranks: the slowest GPU's median and minimum SM clock, any throttle reason seen anywhere
    vec = torch.tensor(mine if mine else [0.0] * 8, dtype=torch.float64, device=dev)
    lo, hi = vec.clone(), vec.clone()
    if world > 1:
        if not mine:
            lo[:2] = float("inf")
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        tot = vec[7:8].clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        nsamp = float(tot.item())
    else:
        nsamp = float(vec[7].item())
    lo, hi = lo.tolist(), hi.tolist()
    if nsamp > 0 and np.isfinite(lo[0]):
        clocks = {"sm_mhz": lo[0], "sm_min_mhz": lo[1], "sm_max_mhz": hi[2],
                  "reasons": [nm for nm, f in zip(ClockSampler.NAMES, hi[3:7]) if f > 0], "samples": int(nsamp), "gpus": world,
                  "source": "NVML, one sampling thread per rank (20 ms period) during the timed region; sm_mhz = the "
                            "slowest GPU's median"}
    else:
        clocks = {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"], "samples": 0, "gpus": world}
    launches = ens.profile()["kernel_launches"] - launches0
    t_dev = max_over_ranks(t_dev)

    # ---- per-kernel device times (CUDA events on the ensemble's stream; graph off while profiling) ----
    ens.set_profiling(True)
    rb_T = ens.rad_lookahead_steps()                 # steps per radiation look-ahead block (0: per-step kernel)
    nprof = min(max(K // 4, 20), 100) if not rb_T else rb_T * max(2, round(100 / rb_T))   # whole blocks
    for _ in range(3):
        dev_step()
    ens.sync()
    ens.kernel_ms(reset=True)
    ens.rad_block_stats(reset=True)
    for _ in range(nprof):
        dev_step()
    ens.sync()
    kms = ens.kernel_ms(reset=True)
    rb_stats = ens.rad_block_stats(reset=True)
    rb_on = bool(rb_T) and rb_stats["steps_served"] == nprof
    ens.set_profiling(False)

    # ---- end-to-end leg (host buffers through hc_step) -----------------------------------------------
    for _ in range(W):
        host_step()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        host_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier()
    t_e2e = max_over_ranks(t_e2e)
    checksum = float(h_force.numpy()[:, 2].sum())

    # ---- the same device-resident leg with bit-faithful bracketing (bracket_snap = 0), short ----------------
    faithful = None
    if snap > 0:
        ens.set_bracket_snap(0.0)
        kf = max(50, K // 5)
        for _ in range(W):
            dev_step()
        ens.sync()
        barrier()
        ev0.record(stream)
        for _ in range(kf):
            dev_step()
        ev1.record(stream)
        ens.sync()
        torch.cuda.synchronize()
        t_f = max_over_ranks(ev0.elapsed_time(ev1) * 1e-3)
        faithful = {"value": world * B * kf / t_f, "unit": "instance-steps/s", "steps": kf, "ms_per_step": 1e3 * t_f / kf,
                    "note": "bracket_snap = 0: every convolution query is bracketed with the reference's == / lerp "
                            "logic, a ~1e-9-weight second history row is read where the query misses a sample by "
                            "floating-point noise"}
        ens.set_bracket_snap(snap)

    fp64_peak = hc.measure_fp64_peak(local_rank) if rank == 0 else None
    mma_peak = hc.measure_fp64_mma_peak(local_rank) if rank == 0 else None
    if rank == 0:
        peak, peak_src = hbm_peak()
        # distinct history rows touched per step: one per lag with exact hits / snapping, up to two otherwise
        rows = RIRF_STEPS if snap > 0 else min(2 * RIRF_STEPS, hist_rows)
        rad_bytes = B * (8 * DOFS * rows + 8 * DOFS * 3) + 8 * DOFS * DOFS * RIRF_STEPS
        exc_bytes = B * 8 * (EXC_STEPS + 1) + 8 * (DOFS + 2) * EXC_STEPS
        ach = rad_bytes / (kms["radiation"] * 1e-3) / 1e9
        ach_exc = exc_bytes / (kms["excitation"] * 1e-3) / 1e9 if kms["excitation"] > 0 else None
        exc_flops = 2.0 * DOFS * EXC_STEPS * B                      # per step: 2 * D * Le per instance
        rad_flops = 2.0 * DOFS * DOFS * RIRF_STEPS * B
        exc_tf = exc_flops / (kms["excitation"] * 1e-3) / 1e12 if kms["excitation"] > 0 else None
        rad_tf = rad_flops / (kms["radiation"] * 1e-3) / 1e12
        traffic = ncu_traffic() if WORKLOAD == "rm3_irregular_ensemble" else None
        value = world * B * K / t_dev
        step_bytes = rad_bytes + (exc_bytes if args.no_lookahead else (B * 8 * (EXC_STEPS + 8) // 8))
        step_flops = rad_flops + exc_flops
        exc_mma = (not args.no_lookahead) and args.lookahead_mode in (0, 4, 5)
        exc_peak = mma_peak if exc_mma else fp64_peak
        rad_roof = {"bound": "hbm", "kernel": ("k_radiation_mma12 (DMMA m8n8k4)" if (DOFS == 12 and args.rad_kernel == 2)
                                               else "k_radiation_hybrid12 (DMMA + DFMA)" if (DOFS == 12 and args.rad_kernel == 3)
                                               else "k_radiation<%d>" % DOFS), "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": (traffic or {}).get("radiation_dram_bytes_per_launch"),
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": rad_bytes,
                    "kernel_ms": kms["radiation"],
                    "fp64_tflops": rad_tf, "fp64_peak_tflops": fp64_peak,
                    "fp64_frac": rad_tf / fp64_peak if fp64_peak else None}
        if rb_on:
            # k_rad_block<12>: one launch = the resident rows' share of rb_T steps; lags served by the block at block
            # step j: (L - 1) - floor(j / m)  (the younger lags belong to k_rad_step)
            m = rb_T // 8
            lag_steps = sum((RIRF_STEPS - 1) - (j // m) for j in range(rb_T))
            blk_flops = 2.0 * DOFS * DOFS * lag_steps * B
            blk_ms = rb_stats["avg_ms"]
            blk_tf = blk_flops / (blk_ms * 1e-3) / 1e12
            hist_bytes = 8.0 * DOFS * B * m * (RIRF_STEPS - 1)         # every resident row once per block
            rad_roof = {"bound": "tensor", "kernel": "k_rad_block<%d> (radiation look-ahead: %d steps per pass over the "
                                                     "history, FP64 tensor cores, DMMA m8n8k4)" % (DOFS, rb_T),
                        "achieved": blk_tf, "peak": mma_peak, "unit": "TFLOP/s", "frac": blk_tf / mma_peak,
                        "traffic": (traffic or {}).get("rad_block_dram_bytes_per_launch"),
                        "peak_source": "FP64 tensor-core peak measured in this run (hc_measure_fp64_mma_peak, DMMA m8n8k4 "
                                       "loop); MEASURED_PEAKS.json holds no FP64 figure (B200 nominal: 40 TFLOP/s)",
                        "algorithmic_flops_per_launch": blk_flops, "launch_ms": blk_ms, "steps_per_launch": rb_T,
                        "history_bytes_per_launch": hist_bytes,
                        "kernel_ms": kms["radiation"],
                        "kernel_ms_note": "launch_ms = one whole pass, extrapolated from the per-step slices timed in "
                                          "the profiling pass (a pass launched whole, --rad-lookahead 3, measures "
                                          "6.4 ms = 0.95 of the peak: profiles/r01d_radblock.summary.txt); radiation "
                                          "per step = launch_ms / %d + k_step<12>" % rb_T,
                        "hbm_view": {"algorithmic_bytes_per_step": rad_bytes, "gbs": ach, "hbm_peak": peak, "frac": ach / peak,
                                     "note": "SURVEY 8(d) bytes of the per-step formulation over the measured radiation "
                                             "time per step: the block pass reads each history row once per %d steps, "
                                             "so the per-step HBM roofline no longer binds (frac > 1); "
                                             "--rad-lookahead 1 measures the per-step kernel k_radiation<12> against "
                                             "that roofline" % rb_T}}
            step_bytes = hist_bytes / rb_T + (B * 8 * (EXC_STEPS + 8) // 8)
        step_s = t_dev / K
        e2e = world * B * K / t_e2e
        line = {
            "metric": "batched sim-steps/sec (RM3 irregular ensemble)", "value": value, "unit": "instance-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * t_dev / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "instances_per_gpu": B, "instances_total": world * B,
                       "dofs": DOFS, "rirf_steps": RIRF_STEPS, "exc_irf_steps": EXC_STEPS, "dt": DT,
                       "spectrum_components": nf, "eta_samples": n_eta, "sea_state": SEA,
                       "bracket_snap": snap, "history_prefill_steps": prefill, "cuda_graph": not args.no_graph,
                       "excitation_lookahead_steps": 1 if args.no_lookahead else 8,
                       "radiation_lookahead_steps": rb_T if rb_on else 1,
                       "excitation_lookahead_mode": ("off" if args.no_lookahead else
                                                     {0: "background stream, DMMA", 2: "in-stream", 3: "background stream", 4: "in-stream, DMMA",
                                                      5: "background stream, DMMA"}
                                                     .get(args.lookahead_mode, str(args.lookahead_mode))),
                       "l2": "inputs larger than L2 (126 MB), no flush needed: %.1f GB history window + %.1f GB eta per "
                             "GPU, streamed once per %d / 8 steps (~%.2f GB of them touched per step)"
                             % (8e-9 * DOFS * B * ens.history_len(), 8e-9 * B * n_eta, rb_T if rb_on else 1,
                                1e-9 * step_bytes),
                       "timing": "value: CUDA events on the ensemble's stream around K hc_step_device calls (wall %.4f s, "
                                 "of which %.4f s to enqueue them); "
                                 "e2e: wall clock around K synchronous hc_step calls; per-kernel ms: CUDA events inside "
                                 "the library on the same stream" % (t_wall, t_enq)},
            "e2e": {"value": e2e, "unit": "instance-steps/s", "h2d_bytes_per_step": 2 * B * DOFS * 8 + 64,
                    "d2h_bytes_per_step": B * DOFS * 8, "ms_per_step": 1e3 * t_e2e / K},
            "gpu_launches": int(launches) * world,
            "launches_before_timed_region": int(launches0),      # ncu --launch-skip for a launch list of the timed steps
            "clocks": clocks,
            "roofline": {**rad_roof,
                         "excitation": ({"kernel": ("k_exc_block_mma<%d> (look-ahead, 8 steps per eta pass, DMMA m8n8k4)" % DOFS
                                                    if args.lookahead_mode in (0, 4, 5) else
                                                    "k_exc_block<%d> (look-ahead, 8 steps per eta pass)" % DOFS), "bound": "fp64",
                                         "achieved": exc_tf, "peak": exc_peak, "unit": "TFLOP/s",
                                         "frac": exc_tf / exc_peak if (exc_tf and exc_peak) else None,
                                         "peak_source": "measured in this run (%s)" % ("hc_measure_fp64_mma_peak, DMMA loop" if exc_mma
                                                                                        else "hc_measure_fp64_peak, DFMA loop"),
                                         "flops_per_step": exc_flops, "kernel_ms_per_step": kms["excitation"],
                                         "traffic": (traffic or {}).get("exc_block_dram_bytes_per_launch"),
                                         "traffic_note": "dram bytes per k_exc_block launch (one launch per 8 steps)"}
                                        if not args.no_lookahead else
                                        {"kernel": "k_excitation<12>", "bound": "hbm", "achieved": ach_exc, "peak": peak,
                                         "unit": "GB/s", "frac": (ach_exc / peak) if ach_exc else None,
                                         "algorithmic_bytes_per_launch": exc_bytes, "kernel_ms": kms["excitation"],
                                         "traffic": (traffic or {}).get("excitation_dram_bytes_per_launch")})},
            "step_roofline": {"hbm_gbs": step_bytes / step_s / 1e9, "hbm_frac": step_bytes / step_s / 1e9 / peak,
                              "fp64_tflops": step_flops / step_s / 1e12,
                              "fp64_frac": (step_flops / step_s / 1e12 / max(fp64_peak, mma_peak)) if fp64_peak else None,
                              "fp64_peaks_tflops": {"fma_pipe": fp64_peak, "tensor_dmma": mma_peak},
                              "note": "whole step (all kernels, overlapped): algorithmic bytes and flops per step over "
                                      "the measured step time; the step needs both resources at once"},
            "kernel_ms": kms,
            "kernel_ms_note": "isolated kernel durations (the profiling pass runs every kernel back-to-back in one "
                              "stream).  excitation = look-ahead block time / 8.  With the radiation look-ahead on: "
                              "radiation = this step's slice of the next block's k_rad_block<12> pass + k_step<12> (append, "
                              "block partials, rows appended since the snapshot AND finalize, fused), finalize ~ 0.  "
                              "In the timed region the excitation block of the next 8 steps and the slices of the next "
                              "radiation block run on side streams underneath the per-step kernels and the host <-> "
                              "device copies, so ms_per_step < sum(kernel_ms)",
            "setup": {"eta_synthesis_s": eta_s, "setup_and_prefill_s": setup_s},
            "checksum": checksum,
            "faithful_bracketing": faithful,
        }
        if not args.no_cpu:
            # the CPU leg runs in a fresh interpreter (no torch / CUDA threads competing for the cores): the same
            # code path as `bench.py --impl reference`
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", WORKLOAD,
                                      "--cpu-instances",
                                      str(args.cpu_instances), "--cpu-steps", str(args.cpu_steps), "--prefill", str(prefill),
                                      "--batch", str(B)], capture_output=True, text=True, timeout=900,
                                     env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
                line["cpu_baseline"] = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
            except Exception as ex:   # never lose the GPU line because the CPU leg failed
                line["cpu_baseline"] = {"value": None, "unit": "instance-steps/s", "cores": None, "kind": "port",
                                        "sample": "cpu leg failed: %r" % (ex,)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
