#!/usr/bin/env python
"""bench.py -- accepted RK45 ray-steps per second (fp64) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA engine behind the C ABI)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the reference algorithm on host cores

One "step" = one pass of the ray path (WR.ray_run: t=0 state + RK45 integration of every ray) over
one batch of synthetic sources.  Workload = BASELINE.json configs[1]: global 1-degree synthetic
climatological background ("Clim", seed 20250327), sources at every grid point with |lat| <= 88 deg
(63 720 sources), k = 1..6, stationary, 30 days, tstep 2 h  ->  1 146 960 ray slots, nt = 361,
7 x 361 x 1 146 960 x 8 B = 23.2 GB of trajectories per step.

JSON keys: see the contract in the task statement; `roofline.bound` is "fp64" because the hot kernel
is bound by the FP64 (DFMA) pipe, not by HBM or the tensor cores; an `hbm` view of
the same kernel is given in `roofline.hbm` for completeness.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_STEP = 1342.0          # 6 RHS x 167 + 340 step algebra (algorithmic fp64 flop)
BYTES_PER_OUTPUT = 56.0         # 7 doubles written per (ray, output time)
NOMINAL_FP64_TFLOPS = 37.2      # 148 SM x 64 DFMA/clk x 2 x 1.965 GHz
# context for every CPU figure in the line (the UNMODIFIED Python reference, measured in the build container)
PY_REFERENCE_NOTE = ("For scale, BASELINE.md: the unmodified Python reference runs 1.6e4-1.9e4 accepted steps/s on one core "
                     "(one batch) and 1.31e5 steps/s on 8 cores (8 processes x 200 sources) -- the C port timed here is "
                     "about 60x faster per core, so ratios against it are conservative.")


def workload(name, rank=0):
    """Returns dict(u, v, lon, lat, src_lon_deg, src_lat_deg, zwn, days)."""
    import rwrt_b200  # noqa: F401
    from rwrt_b200 import synthetic
    if name == "c2":
        nlon, nlat = 360, 181
        u, v = synthetic.clim(nlon, nlat)
        slon, slat = synthetic.all_gridpoint_sources(nlon, nlat, 88.0)
        # weak scaling: every rank integrates BASELINE configs[1] itself (identical replicas).  RWRT_BENCH_SHIFT_SOURCES=1
        # shifts rank r's sources by 0.046 r degrees instead (round 1's definition): the pass is then bounded by the
        # longest single ray of a rank's set -- rank 1's set holds one that takes 7x the mean number of steps, a serial
        # chain of 45 ms, whatever the scheduling
        if os.environ.get("RWRT_BENCH_SHIFT_SOURCES") == "1":
            slon = (slon + 0.37 * rank / 8.0) % 360.0
        zwn, days = np.arange(1.0, 7.0), 30.0
    elif name in ("c2-small", "c2-q"):      # same physics, 1/16 or 1/4 of the sources: quick checks / profiling
        stride = 16 if name == "c2-small" else 4
        nlon, nlat = 360, 181
        u, v = synthetic.clim(nlon, nlat)
        slon, slat = synthetic.all_gridpoint_sources(nlon, nlat, 88.0)
        slon, slat = slon[rank % stride::stride], slat[rank % stride::stride]
        zwn, days = np.arange(1.0, 7.0), 30.0
    elif name in ("c5", "c5-small", "c5-full"):
        # BASELINE configs[4]: 0.25 deg Clim background, 1440 x 386 sources (|lat| <= 48.125 deg) x k=1..6 x 3 roots
        # = 10.0 M ray slots, 60 days.  c5-full: all of them (the `strong` block shards this list over the ranks);
        # c5: one eighth (every 8th source, offset by rank): 1.25 M slots, 50 GB of trajectories -- the per-GPU share
        # of an 8-GPU run as a weak-scaling workload (--workload c5); c5-small: 1/128 (quick checks).
        nlon, nlat = 1440, 721
        u, v = synthetic.clim(nlon, nlat)
        slon, slat = synthetic.all_gridpoint_sources(nlon, nlat, 48.125)
        stride = {"c5": 8, "c5-small": 128, "c5-full": 1}[name]
        slon, slat = slon[rank % stride::stride], slat[rank % stride::stride]
        zwn, days = np.arange(1.0, 7.0), 60.0
    elif name in ("c4", "c4-full"):
        # BASELINE configs[3]: 1000-member ensemble of perturbed 1 deg backgrounds, 56 sources x k=1..6 x 3 roots
        # (1008 ray slots) per member.  c4 = one GPU's share of the 8-GPU run (125 members, offset by rank);
        # c4-full = all 1000 members on one GPU (6.3 GB of packed fields, 1.0 M slots).
        nlon, nlat = 360, 181
        nmem = 125 if name == "c4" else 1000
        first = (rank % 8) * 125 if name == "c4" else 0
        mem = [synthetic.ensemble_member(first + e, nlon, nlat) for e in range(nmem)]
        u, v = np.stack([m[0] for m in mem]), np.stack([m[1] for m in mem])
        rng = np.random.default_rng(4)
        slon = np.tile(rng.uniform(0.0, 360.0, 56), nmem)
        slat = np.tile(rng.uniform(-70.0, 70.0, 56), nmem)
        member_of = np.repeat(np.arange(nmem, dtype=np.int32), 56)
        zwn, days = np.arange(1.0, 7.0), 30.0
        lon, lat = synthetic.grid_coords(nlon, nlat)
        return dict(u=u, v=v, lon=lon, lat=lat, src_lon_deg=slon, src_lat_deg=slat, zwn=zwn, days=days,
                    nlon=nlon, nlat=nlat, member_of=member_of, nmember=nmem)
    else:
        raise SystemExit(f"unknown workload {name}")
    lon, lat = synthetic.grid_coords(nlon, nlat)
    return dict(u=u, v=v, lon=lon, lat=lat, src_lon_deg=slon, src_lat_deg=slat, zwn=zwn, days=days,
                nlon=nlon, nlat=nlat)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def mark(self):
        """Start of the timed region: only samples that arrive after this count (nvidia-smi needs ~1 s to start,
        so the sampler itself is started before the warm-up)."""
        self.t_mark = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t_mark = getattr(self, "t_mark", 0.0)
        rows = [r[1:] for r in self.rows if r[0] >= t_mark]
        window = "timed region"
        if not rows:   # timed region shorter than one sampling period: the last samples of the (identical) warm-up load
            rows, window = [r[1:] for r in self.rows[-5:]], "warm-up (timed region shorter than the sampling period)"
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on the host cores (oracle C restatement)
# ---------------------------------------------------------------------------------------------
def _cpu_worker(args):
    wl_name, idx, days = args
    from oracle import oracle
    wl = workload(wl_name)
    F = oracle.bs_ready(wl["u"], wl["v"], wl["lat"], True)
    slon, slat = np.radians(wl["src_lon_deg"][idx]), np.radians(wl["src_lat_deg"][idx])
    nt = int(days * 86400.0 / 7200.0) + 1
    t0 = time.perf_counter()
    out = oracle.ray_init(F, wl["lon"], wl["lat"], slon, slat, wl["zwn"], 0.0, nt)
    acc, _ = oracle.ray_run_rk45(F, wl["lon"], wl["lat"], out, 7200.0, days * 86400.0)
    return int(acc.sum()), time.perf_counter() - t0


def cpu_sample(wl_name, nsrc, cores, seed=0):
    """Time the oracle on a seeded random sample of `nsrc` sources of the workload, on `cores` processes."""
    import multiprocessing as mp
    wl = workload(wl_name)
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(len(wl["src_lon_deg"]), size=min(nsrc, len(wl["src_lon_deg"])), replace=False))
    parts = [p for p in np.array_split(idx, cores) if len(p)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker((wl_name, parts[0], wl["days"]))]
    else:
        with mp.get_context("fork").Pool(len(parts)) as pool:
            res = pool.map(_cpu_worker, [(wl_name, p, wl["days"]) for p in parts])
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return steps, slowest, wall, len(idx)


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import build_oracle
    build_oracle.build()
    cores = os.cpu_count() or 1
    cores = min(cores, 64)
    nsrc = a.ref_sources or max(960, 500 * cores)
    vals = []
    for i in range(a.warmup + a.steps):
        steps, slowest, wall, n = cpu_sample(a.workload, nsrc, cores, seed=i)
        if i >= a.warmup:
            vals.append((steps, slowest))
    steps = sum(v[0] for v in vals)
    secs = sum(v[1] for v in vals)
    value = steps / secs
    sample = (f"seeded random sample of {n} of the {a.workload} sources per step, split over {cores} processes "
              f"(oracle C restatement of the reference algorithm, gcc -O2 -ffp-contract=off; the Python reference "
              f"itself cannot travel to this box: see DESIGN.md).  " + PY_REFERENCE_NOTE)
    line = {"impl": "reference", "metric": "accepted RK45 ray-steps/s (fp64)", "value": value, "unit": "steps/s",
            "n_gpus": 0, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * secs / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": a.workload, "sample_sources_per_step": n, "cores": cores},
            "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import rwrt_b200  # noqa: F401
    from rwrt_b200 import build as rbuild
    from rwrt_b200 import lib

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        rbuild.build()
    if world > 1:
        dist.barrier()
    lib.require_gpu()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = lib.load()
    wl = workload(a.workload, int(os.environ.get("RWRT_BENCH_RANK_SHIFT", rank)))   # (env: time another rank's source set on one GPU)
    nlon, nlat = wl["nlon"], wl["nlat"]
    nsource, nzwn = len(wl["src_lon_deg"]), len(wl["zwn"])
    tstep, ttotal = 7200.0, wl["days"] * 86400.0
    nt = int(ttotal / tstep) + 1
    N = 3 * nsource * nzwn
    slon = np.radians(wl["src_lon_deg"]); slat = np.radians(wl["src_lat_deg"])
    flags = lib.MODE_STRICT if a.strict else lib.MODE_FAST

    h = lib.vp()
    nmember = int(wl.get("nmember", 1))
    lib.check(L.rwrt_create(local, nlon, nlat, 1, nmember, C.byref(h)), "rwrt_create")
    stream = torch.cuda.current_stream()
    sp = lib.vp(stream.cuda_stream)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(u=wl["u"], v=wl["v"], lon=wl["lon"], lat=wl["lat"], cl=np.cos(wl["lat"]), sl=np.sin(wl["lat"]),
              slon=slon, slat=slat, zwn=wl["zwn"]).items()}
    mem_ptr = None
    if "member_of" in wl:
        d["mem"] = torch.from_numpy(np.ascontiguousarray(wl["member_of"])).to(dev)
        mem_ptr = lib.ptr(d["mem"])
        a.no_e2e = a.no_cpu = True                          # the host-pointer entry and the CPU sample are single-member
    lib.check(L.rwrt_bs_ready(h, lib.ptr(d["u"]), lib.ptr(d["v"]), lib.ptr(d["lon"]), lib.ptr(d["lat"]),
                              lib.ptr(d["cl"]), lib.ptr(d["sl"]), None, None, sp), "rwrt_bs_ready")
    out = torch.empty((7, nt, N), dtype=torch.float64, device=dev)
    acc = torch.zeros(N, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def one_step():
        lib.check(L.rwrt_ray_init(h, nsource, lib.ptr(d["slon"]), lib.ptr(d["slat"]), mem_ptr, nzwn, lib.ptr(d["zwn"]),
                                  0.0, nt, lib.ptr(out), sp), "rwrt_ray_init")
        lib.check(L.rwrt_ray_run_rk45(h, nsource, mem_ptr, nzwn, nt, tstep, ttotal, 1e-6, 1e-6, 1e-3, 0.1, flags,
                                      lib.ptr(out), lib.ptr(acc), None, None, lib.ptr(total), sp), "rwrt_ray_run_rk45")

    stats = (C.c_double * 3)()
    times = (C.c_double * 4)()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(a.warmup):
        one_step()
        flush.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.mark()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    kern_ms, integ_ms, steps_done, launches = [], [], 0, 0
    torch.cuda.synchronize()
    for i in range(a.steps):
        ev[i][0].record(stream)
        one_step()
        ev[i][1].record(stream)
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
        L.rwrt_last_run_stats(h, stats)     # syncs on the run's end event
        L.rwrt_last_run_times(h, times)     # CUDA events recorded by the library on the launching stream, around the kernels
        launches += int(stats[0]) + 1       # + ray_init
        kern_ms.append(stats[2])
        integ_ms.append(times[1])           # the integrate kernel alone (the dominant kernel)
        steps_done += int(total.item())
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    if os.environ.get("RWRT_BENCH_VERBOSE") and rank == 0:
        print("per-step kernel ms:", [round(k, 2) for k in kern_ms], file=sys.stderr)
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    s = torch.tensor([float(steps_done)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    ms_max, steps_all = float(t.item()), float(s.item())
    value = steps_all / (ms_max * 1e-3)

    # ---- end-to-end through the host-buffer C-ABI call (H2D of inputs, D2H of the 7 result arrays) ----
    e2e = None
    if not a.no_e2e:
        e2e = run_e2e(a, L, h, wl, slon, slat, nsource, nzwn, nt, N, tstep, ttotal, flags, world, dev)

    # ---- the same pass with the ray records as the native result (no transpose to the reference planes) ----
    records = None
    if not a.no_records and "member_of" not in wl:
        records = run_records(a, L, h, d, nsource, nzwn, nt, N, tstep, ttotal, flags, sp, stream, flush, dev)
    # ---- grid-only delivery end to end (host inputs -> sweep -> all-reduce -> grids on the host) ----
    e2e_grid = None
    if not a.no_e2e and "member_of" not in wl:
        e2e_grid = run_e2e_grid(a, wl, nsource, world, rank, dev)
    del out
    torch.cuda.empty_cache()
    L.rwrt_trim(h)
    # ---- north_star's multi-GPU path: the 10 M-slot 0.25-degree sweep sharded over the ranks, gridded, all-reduced ----
    strong = None
    if not a.no_strong and (a.workload == "c2" or a.strong):   # part of the default (driver) line; opt-in for other workloads
        strong = run_strong(a, L, rank, world, local, dev, flags)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel: rk45_integrate_coop alone, timed by CUDA events on its stream ----
    peak_tf = C.c_double(0)
    L.rwrt_measure_fp64_peak(h, C.byref(peak_tf))
    if a.save_peak:
        with open(a.save_peak, "w") as f:
            f.write(f"rwrt_measure_fp64_peak (register-resident DFMA chains, 8 per thread, 8 x 256 threads per SM, best of 5): "
                    f"{peak_tf.value:.3f} TFLOP/s; nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = {NOMINAL_FP64_TFLOPS} TFLOP/s; "
                    f"clocks during the timed region of this run: {json.dumps(clocks)}\n")
    step_ms = float(np.mean(kern_ms))
    kms = float(np.mean(integ_ms)) if integ_ms and min(integ_ms) > 0 else step_ms
    steps_per_launch = steps_done / a.steps
    achieved_tf = steps_per_launch * FLOP_PER_STEP / (kms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_alg = steps_per_launch / 1.07 * BYTES_PER_OUTPUT   # ~1.07 accepted steps per (ray, output time)
    traffic = traffic_step = traffic_src = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the COMMITTED ncu --set full captures (not measured in this run)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("workload") == a.workload:
            traffic = tj.get("dram_bytes_per_launch")
            traffic_step = tj.get("dram_bytes_per_step")
            traffic_src = tj.get("source")
    except Exception:
        pass
    peak = peak_tf.value or NOMINAL_FP64_TFLOPS
    roof = {"bound": "fp64", "achieved": achieved_tf, "peak": peak,
            "unit": "TFLOP/s", "frac": achieved_tf / peak,
            "peak_source": "measured DFMA-chain microbenchmark in this run (rwrt_measure_fp64_peak); nominal 37.2; "
                           "MEASURED_PEAKS.json has no FP64 figure",
            "flop_per_step": FLOP_PER_STEP, "kernel_ms": kms, "traffic": traffic,
            "traffic_source": traffic_src and f"committed capture, not measured in this run: {traffic_src}",
            "traffic_step": traffic_step,
            "kernel": "rk45_integrate_coop<OUT_ROWS> alone (CUDA events on the launching stream around its launch, "
                      "rwrt_last_run_times[1])",
            "step": {"ms": step_ms, "frac": steps_per_launch * FLOP_PER_STEP / (step_ms * 1e-3) / 1e12 / peak,
                     "what": "the whole pass in one event pair: compaction result wait excluded, rk45_prepare + integrator + "
                             "expand_records transpose to the reference planes"},
            "hbm": {"achieved": hbm_alg / (kms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": hbm_alg / (kms * 1e-3) / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
    # ---- CPU baseline: the oracle port on one core, bounded sample ----
    cpu = None
    if not a.no_cpu:
        from oracle import build_oracle
        build_oracle.build()
        st, slowest, wall, n = cpu_sample(a.workload, a.cpu_sources, 1, seed=0)
        cpu = {"value": st / slowest, "unit": "steps/s", "cores": 1, "kind": "port",
               "sample": f"seeded random sample of {n} sources of the workload ({st} accepted steps, {slowest:.1f} s), "
                         "oracle C restatement, single thread.  " + PY_REFERENCE_NOTE}
    line = {"metric": "accepted RK45 ray-steps/s (fp64)", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{a.workload}: Clim {360.0 / nlon:g}deg background, {nsource} sources x k=1..{nzwn} x 3 roots = {N} "
                                   f"ray slots per GPU, {wl['days']:.0f} d, tstep 2 h, nt={nt}, rtol=atol=1e-6" + (f", {nmember} ensemble members" if nmember > 1 else ""),
                       "mode": "strict" if a.strict else "fast", "accepted_steps_per_step_per_gpu": steps_per_launch,
                       "l2": f"flushed between timed iterations (256 MB write); outputs ({7 * nt * N * 8 / 1e9:.1f} GB) exceed L2 anyway",
                       "timed_region": "rwrt_ray_init + rwrt_ray_run_rk45 (fields resident in HBM)",
                       "ranks": "every rank runs this same configuration (identical replicas, no data-path collective); the sharded 10 M-slot "
                                "sweep with its all-reduce is the `strong` block"},
            "clocks": clocks, "e2e": e2e, "e2e_grid": e2e_grid, "gpu_launches": launches, "roofline": roof,
            "records": records, "strong": strong, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    L.rwrt_destroy(h)
    if world > 1:
        dist.destroy_process_group()


def run_records(a, L, h, d, nsource, nzwn, nt, N, tstep, ttotal, flags, sp, stream, flush, dev):
    """The timed pass of `value` with the ray records as the native result (rwrt_ray_run_rk45_records): the integrator's
    64-byte records stay on the device and the transpose to the reference's seven planes happens only on demand."""
    import torch
    from rwrt_b200 import lib
    row0 = torch.empty((7, N), dtype=torch.float64, device=dev)
    acc = torch.zeros(N, dtype=torch.int64, device=dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    times = (C.c_double * 4)()

    def one():
        lib.check(L.rwrt_ray_init(h, nsource, lib.ptr(d["slon"]), lib.ptr(d["slat"]), None, nzwn, lib.ptr(d["zwn"]), 0.0, 1,
                                  lib.ptr(row0), sp), "rwrt_ray_init")
        lib.check(L.rwrt_ray_run_rk45_records(h, nsource, None, nzwn, nt, tstep, ttotal, 1e-6, 1e-6, 1e-3, 0.1, flags,
                                              lib.ptr(row0), lib.ptr(acc), None, None, lib.ptr(total), sp),
                  "rwrt_ray_run_rk45_records")
    try:
        one()
        flush.zero_()
        torch.cuda.synchronize()
        ms, integ, steps = [], [], 0
        for _ in range(a.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            one()
            e1.record(stream)
            flush.zero_()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
            L.rwrt_last_run_times(h, times)
            integ.append(times[1])
            steps += int(total.item())
        return {"value": steps / (sum(ms) * 1e-3), "unit": "steps/s", "ms_per_step": float(np.mean(ms)),
                "integrate_ms": float(np.mean(integ)),
                "what": "rwrt_ray_init + rwrt_ray_run_rk45_records: same rays, results left as 64-byte ray records on the device "
                        "(13 GB of DRAM writes per pass instead of 50.9 GB with the transpose); rwrt_records_expand / "
                        "rwrt_grid_accumulate_records consume them on demand"}
    except lib.RwrtError as e:
        return {"value": None, "note": f"skipped: {e}"}


def run_strong(a, L, rank, world, local, dev, flags):
    """BASELINE configs[4] as ONE job: 0.25-degree Clim background, 1440 x 386 sources x k=1..6 x 3 roots = 10 005 120
    ray slots, 60 days, the sources sharded over the ranks (dist.shard_indices); every rank runs rwrt_ray_sweep_grid on
    its share (no trajectory arrays: the reference's layout would need 404 GB, wr.py:157-167) and the (3, 4320, 720)
    int64 ray-density / flux accumulators are merged by ONE NCCL all-reduce -- both inside the event pair."""
    import torch
    import torch.distributed as dist
    from rwrt_b200 import dist as rdist
    from rwrt_b200 import gridding, lib
    wl = workload("c5-full")
    nlon, nlat = wl["nlon"], wl["nlat"]
    nsrc_all, nzwn = len(wl["src_lon_deg"]), len(wl["zwn"])
    idx = rdist.shard_indices(nsrc_all, rank, world, a.shard)
    slon, slat = np.radians(wl["src_lon_deg"][idx]), np.radians(wl["src_lat_deg"][idx])
    nsource = len(idx)
    tstep, ttotal = 7200.0, wl["days"] * 86400.0
    nt = int(ttotal / tstep) + 1
    cell = (0.25, 0.25)
    ncx, ncy = gridding.grid_shape(*cell)
    h = lib.vp()
    lib.check(L.rwrt_create(local, nlon, nlat, 1, 1, C.byref(h)), "rwrt_create")
    stream = torch.cuda.current_stream()
    sp = lib.vp(stream.cuda_stream)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(u=wl["u"], v=wl["v"], lon=wl["lon"], lat=wl["lat"], cl=np.cos(wl["lat"]), sl=np.sin(wl["lat"]),
              slon=slon, slat=slat, zwn=wl["zwn"]).items()}
    lib.check(L.rwrt_bs_ready(h, lib.ptr(d["u"]), lib.ptr(d["v"]), lib.ptr(d["lon"]), lib.ptr(d["lat"]),
                              lib.ptr(d["cl"]), lib.ptr(d["sl"]), None, None, sp), "rwrt_bs_ready")
    grids = gridding.new_grids(*cell, dev)
    total = torch.zeros(1, dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    ar = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

    def one_pass():
        grids.zero_()
        lib.check(L.rwrt_ray_sweep_grid(h, nsource, lib.ptr(d["slon"]), lib.ptr(d["slat"]), None, nzwn, lib.ptr(d["zwn"]), 0.0,
                                        nt, tstep, ttotal, 1e-6, 1e-6, 1e-3, 0.1, flags, cell[0], cell[1], ncx, ncy,
                                        lib.ptr(grids[0]), lib.ptr(grids[1]), lib.ptr(grids[2]), None, None, lib.ptr(total), sp),
                  "rwrt_ray_sweep_grid")
        ar[0].record(stream)
        if world > 1:
            dist.all_reduce(grids, op=dist.ReduceOp.SUM)      # the path's one collective (NCCL over NVLink)
        ar[1].record(stream)

    nsteps = max(1, min(a.steps, a.strong_steps))
    one_pass()                                                # warm-up (also sizes the handle's work lists)
    flush.zero_()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms, ms_ar, steps = [], [], 0
    for _ in range(nsteps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        one_pass()
        e1.record(stream)
        flush.zero_()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        ms_ar.append(ar[0].elapsed_time(ar[1]))
        steps += int(total.item())
    if world > 1:
        dist.barrier()
    t = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
    s = torch.tensor([float(steps)], dtype=torch.float64, device=dev)
    per_rank = torch.zeros(world, dtype=torch.float64, device=dev)
    per_rank[rank] = sum(ms) / nsteps
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    csum = gridding.checksum(grids)                           # of the all-reduced grids: must not depend on N
    count_total = int(grids[0].sum().item())
    L.rwrt_destroy(h)
    return {"metric": "accepted RK45 ray-steps/s (fp64), whole 10 M-slot sweep incl. gridding + all-reduce",
            "value": float(s.item()) / (float(t.item()) * 1e-3), "unit": "steps/s", "scaling": "strong", "n_gpus": world,
            "steps": nsteps, "warmup": 1, "ms_per_step": float(t.item()) / nsteps,
            "ms_allreduce": float(np.mean(ms_ar)), "allreduce_bytes": int(grids.numel() * 8),
            "collective": "one torch.distributed all_reduce(SUM) of the (3, 4320, 720) int64 accumulators over NCCL" if world > 1
                          else "none (1 rank)",
            "per_rank_ms": [round(float(x), 2) for x in per_rank.tolist()], "shard": a.shard,
            "accepted_steps_per_pass": float(s.item()) / nsteps, "grid_checksum": f"0x{csum:016x}",
            "grid_samples": count_total,
            "workload": f"c5-full: Clim 0.25deg background, {nsrc_all} sources x k=1..{nzwn} x 3 roots = {3 * nsrc_all * nzwn} ray slots "
                        f"sharded over {world} GPU(s) ({3 * nsource * nzwn} on rank 0), 60 d, tstep 2 h, nt={nt}, gridded on 0.25deg cells "
                        f"(three longitude circles); no (7, nt, N) arrays exist (they would be {7 * nt * 3 * nsrc_all * nzwn * 8 / 1e9:.0f} GB)",
            "api": "rwrt_ray_sweep_grid (C ABI, device pointers) + torch.distributed.all_reduce"}


def run_e2e_grid(a, wl, nsource, world, rank, dev):
    """Grid-only delivery end to end through the package's public API (WR facade + gridding.sweep + dist.allreduce_grids):
    every call uploads the winds and the sources from host memory, recomputes the basic state, sweeps, all-reduces the
    grids and brings them to the host -- what a user who wants ray density / flux (not 23 GB of trajectories) runs."""
    import torch
    import torch.distributed as dist
    from rwrt_b200 import dist as rdist
    from rwrt_b200 import gridding
    from rwrt_b200.wr import WR
    nlon, nlat = wl["nlon"], wl["nlat"]
    cell = (360.0 / nlon, 180.0 / (nlat - 1))
    w = WR(len(wl["zwn"]), nsource, 7200.0, wl["days"] * 86400.0, 0.0, nx=nlon, ny=nlat, device=dev.index)
    w.bs.lon, w.bs.lat = wl["lon"], wl["lat"]
    w.set_zwn(wl["zwn"])
    host = torch.empty((3,) + gridding.grid_shape(*cell), dtype=torch.int64, pin_memory=True)

    def call():
        w.bs.u, w.bs.v = wl["u"], wl["v"]                                  # host arrays: H2D inside bs.ready
        w.bs.ready(xcyclic=True, keep_fields18=False, keep_diag=False)
        w.set_source_array(wl["src_lon_deg"], wl["src_lat_deg"])           # degrees on the host, like the reference
        g = gridding.sweep(w, *cell, want_accepted=False)
        rdist.allreduce_grids(g)
        host.copy_(g, non_blocking=True)
        torch.cuda.synchronize()
        return w.total_accepted

    call()
    if world > 1:
        dist.barrier()
    reps = max(1, min(a.steps, 8))
    t0 = time.perf_counter()
    steps, per = 0, []
    for _ in range(reps):
        t1 = time.perf_counter()
        steps += call()
        per.append(round(1e3 * (time.perf_counter() - t1), 1))
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    s = torch.tensor([float(steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    h2d = wl["u"].nbytes + wl["v"].nbytes + 4 * wl["lat"].nbytes + 2 * wl["lon"].nbytes + 16 * nsource + 8 * len(wl["zwn"])
    w.clean()
    return {"value": float(s.item()) / float(t.item()), "unit": "steps/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(host.numel() * 8 + 8), "ms_per_step": 1e3 * float(t.item()) / reps, "ms_per_call": per,
            "api": "WR facade: bs.ready + gridding.sweep (rwrt_ray_sweep_grid) + dist.allreduce_grids + grids to pinned host memory",
            "what": f"same sources and rays as `e2e`, delivered as (3, {host.shape[1]}, {host.shape[2]}) int64 ray-density / flux "
                    "grids instead of seven trajectory arrays"}


def host_timeline(L, h):
    """Summary of rwrt_last_host_windows: where the last host-buffer call spent its time."""
    from rwrt_b200 import lib
    n = C.c_int(0)
    buf = (C.c_double * (8 * 256))()
    if L.rwrt_last_host_windows(h, buf, 256, C.byref(n)) != 0 or n.value == 0:
        return None
    rows = np.array(buf[: 8 * min(n.value, 256)]).reshape(-1, 8)
    link_ms = rows[:, 5] - rows[:, 4]
    hxw = rows[:, 0] == 0
    out = {"windows": int(len(rows)), "direct_windows": int((~hxw).sum()),
           "link_busy_ms": round(float(link_ms.sum()), 1), "link_gbs": round(float(rows[:, 1].sum() / max(link_ms.sum(), 1e-9) / 1e6), 1),
           "slowest_window_link_gbs": round(float((rows[:, 1] / np.maximum(link_ms, 1e-9) / 1e6).min()), 1),
           "last_copy_done_ms": round(float(rows[:, 5].max()), 1),
           "gpu_ready_last_ms": round(float(rows[:, 2].max()), 1)}
    ph = (C.c_double * 8)()
    if L.rwrt_last_host_phases(h, ph) == 0:
        out["phases_ms"] = {"arena": round(ph[0], 1), "inputs_queued": round(ph[1], 1), "pipeline_origin": round(ph[2], 1),
                            "windows_queued": round(ph[3], 1), "device_done": round(ph[4], 1), "threads_joined": round(ph[5], 1),
                            "released": round(ph[6], 1), "thread_startup": round(ph[7], 1)}
    if hxw.any():
        out["host_expand_ms_per_window_median"] = round(float(np.median(rows[hxw, 7] - rows[hxw, 6])), 2)
        out["host_expand_ms_per_window_max"] = round(float((rows[hxw, 7] - rows[hxw, 6]).max()), 2)
        out["wait_for_host_threads_ms_max"] = round(float((rows[hxw, 6] - rows[hxw, 5]).max()), 2)
        out["last_window_written_ms"] = round(float(rows[hxw, 7].max()), 1)
    return out


def run_e2e(a, L, h, wl, slon, slat, nsource, nzwn, nt, N, tstep, ttotal, flags, world, dev):
    """Same metric through rwrt_real2d_host: host inputs, pinned host outputs, copies inside the timed region."""
    import torch
    import torch.distributed as dist
    from rwrt_b200 import lib
    nbytes = 7 * nt * N * 8
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 0
    if avail and avail < 2.5 * nbytes * max(1, world):
        return {"value": None, "unit": "steps/s", "note": f"skipped: {avail/1e9:.0f} GB host RAM available, "
                f"{nbytes/1e9:.1f} GB pinned result arrays needed per GPU"}
    host = torch.empty(7 * nt * N, dtype=torch.float64, pin_memory=True)
    acc_h = torch.empty(N, dtype=torch.int64, pin_memory=True)
    tot_h = torch.zeros(1, dtype=torch.int64, pin_memory=True)
    pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in
           dict(u=wl["u"], v=wl["v"], lon=wl["lon"], lat=wl["lat"], cl=np.cos(wl["lat"]), sl=np.sin(wl["lat"]),
                slon=slon, slat=slat, zwn=wl["zwn"]).items()}
    h2d = sum(t.numel() * t.element_size() for t in pin.values())
    d2h = nbytes + N * 8 + 8

    def call():
        lib.check(L.rwrt_real2d_host(h, lib.ptr(pin["u"]), lib.ptr(pin["v"]), lib.ptr(pin["lon"]), lib.ptr(pin["lat"]),
                                     lib.ptr(pin["cl"]), lib.ptr(pin["sl"]), nsource, lib.ptr(pin["slon"]),
                                     lib.ptr(pin["slat"]), None, nzwn, lib.ptr(pin["zwn"]), 0.0, nt, tstep, ttotal,
                                     1e-6, 1e-6, 1e-3, 0.1, flags, lib.ptr(host), lib.ptr(acc_h), None,
                                     lib.ptr(tot_h), None), "rwrt_real2d_host")

    call()                                   # warm-up (first touch of the pinned pages)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # diagnostics only: PCIe generation / performance state / SM clock while the host-buffer calls run (the GPU is ~15 %
    # busy in them; per-call times drift on some boxes and this shows whether the link or the clocks moved with them)
    link_rows, link_proc = [], None
    try:
        link_proc = subprocess.Popen(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pstate,clocks.sm",
                                      "--format=csv,noheader,nounits", "-lms", "1000", "-i", str(dev.index or 0)],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=lambda: [link_rows.append(ln.strip()) for ln in link_proc.stdout], daemon=True).start()
    except Exception:
        link_proc = None
    reps = max(1, min(a.steps, 8))
    t0 = time.perf_counter()
    steps = 0
    per_rep, timelines = [], []
    for _ in range(reps):
        t_rep = time.perf_counter()
        call()                               # synchronises before returning
        per_rep.append(round(1e3 * (time.perf_counter() - t_rep), 1))
        timelines.append(host_timeline(L, h))
        steps += int(tot_h.item())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    s = torch.tensor([float(steps)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    link = None
    try:
        if link_proc is not None:
            link_proc.terminate()
            link_proc.wait(timeout=2)
            cols = [[c.strip() for c in r.split(",")] for r in link_rows if r.count(",") == 3]
            if cols:
                link = {"pcie_gen": sorted({c[0] for c in cols}), "pcie_width": sorted({c[1] for c in cols}),
                        "pstate": sorted({c[2] for c in cols}), "sm_mhz": [min(float(c[3]) for c in cols), max(float(c[3]) for c in cols)],
                        "samples": len(cols)}
    except Exception:
        link = None
    import ctypes as C
    hs = (C.c_double * 5)()
    lib.check(L.rwrt_last_host_stats(h, hs), "rwrt_last_host_stats")   # bytes the library actually moved in the last call
    if hs[3] > 0:
        how = (f"{int(hs[2])} time windows: {int(hs[2] - hs[4])} as packed rows of the live rays that {int(hs[3])} host threads expand into "
               f"the {nbytes/1e9:.1f} GB of result planes, {int(hs[4])} as device-expanded rows copied straight into them (last call)")
    else:
        how = f"device-expanded rows cross the link in {int(hs[2])} time windows"
    return {"value": float(s.item()) / float(t.item()), "unit": "steps/s", "h2d_bytes_per_step": int(hs[1]),
            "d2h_bytes_per_step": int(hs[0]), "result_bytes_per_step": d2h, "ms_per_step": 1e3 * float(t.item()) / reps,
            "ms_per_call": per_rep,          # this rank's calls (the value is total steps / total time, max over ranks)
            # where the fastest and the slowest of those calls spent their time (rwrt_last_host_windows: link rate per time
            # window, host threads' time per window, waits)
            "timeline_fastest_call": timelines[int(np.argmin(per_rep))], "timeline_slowest_call": timelines[int(np.argmax(per_rep))],
            "link_during_calls": link,
            "api": "rwrt_real2d_host (C ABI, host pointers; pinned inputs and result arrays)", "transfer": how}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=5)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="c2")
    p.add_argument("--strict", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-strong", action="store_true")
    p.add_argument("--strong", action="store_true")
    p.add_argument("--no-records", action="store_true")
    p.add_argument("--strong-steps", type=int, default=3)
    p.add_argument("--shard", default="stride", choices=["stride", "block"])
    p.add_argument("--save-peak", default="")
    p.add_argument("--cpu-sources", type=int, default=4000)   # ~14 s of single-thread CPU work on the GPU box
    p.add_argument("--ref-sources", type=int, default=0)
    a = p.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
