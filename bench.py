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


def workload(name, rank=0):
    """Returns dict(u, v, lon, lat, src_lon_deg, src_lat_deg, zwn, days)."""
    import rwrt_b200  # noqa: F401
    from rwrt_b200 import synthetic
    if name == "c2":
        nlon, nlat = 360, 181
        u, v = synthetic.clim(nlon, nlat)
        slon, slat = synthetic.all_gridpoint_sources(nlon, nlat, 88.0)
        # weak scaling: every rank integrates a full C2-sized source set, shifted by a fraction of a cell
        slon = (slon + 0.37 * rank / 8.0) % 360.0
        zwn, days = np.arange(1.0, 7.0), 30.0
    elif name in ("c2-small", "c2-q"):      # same physics, 1/16 or 1/4 of the sources: quick checks / profiling
        stride = 16 if name == "c2-small" else 4
        nlon, nlat = 360, 181
        u, v = synthetic.clim(nlon, nlat)
        slon, slat = synthetic.all_gridpoint_sources(nlon, nlat, 88.0)
        slon, slat = slon[rank % stride::stride], slat[rank % stride::stride]
        zwn, days = np.arange(1.0, 7.0), 30.0
    elif name in ("c5", "c5-small"):
        # BASELINE configs[4]: 0.25 deg Clim background, 1440 x 386 sources (|lat| <= 48.125 deg) x k=1..6 x 3 roots
        # = 10.0 M ray slots, 60 days.  One GPU integrates one eighth of the sources (every 8th, offset by rank):
        # 1.25 M slots, 50 GB of trajectories -- the per-GPU share of the 8-GPU run.
        nlon, nlat = 1440, 721
        u, v = synthetic.clim(nlon, nlat)
        slon, slat = synthetic.all_gridpoint_sources(nlon, nlat, 48.125)
        stride = 8 if name == "c5" else 128
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
              f"itself cannot travel to this box: see DESIGN.md)")
    line = {"impl": "reference", "metric": "accepted RK45 ray-steps/s (fp64)", "value": value, "unit": "steps/s",
            "n_gpus": 0, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * secs / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": a.workload, "sample_sources_per_step": n},
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
    wl = workload(a.workload, rank)
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
    kern_ms, steps_done, launches = [], 0, 0
    torch.cuda.synchronize()
    for i in range(a.steps):
        ev[i][0].record(stream)
        one_step()
        ev[i][1].record(stream)
        flush.zero_()                       # L2 flush between timed iterations (outside the event pair)
        L.rwrt_last_run_stats(h, stats)     # syncs on the integrate kernels' end event
        launches += int(stats[0]) + 1       # + ray_init
        kern_ms.append(stats[2])
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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (rk45_integrate; prepare + fill are inside the same event pair) ----
    peak_tf = C.c_double(0)
    L.rwrt_measure_fp64_peak(h, C.byref(peak_tf))
    kms = float(np.mean(kern_ms))
    steps_per_launch = steps_done / a.steps
    achieved_tf = steps_per_launch * FLOP_PER_STEP / (kms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_alg = steps_per_launch / 1.07 * BYTES_PER_OUTPUT   # ~1.07 accepted steps per (ray, output time)
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of the integrate kernel from the committed ncu capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("workload") == a.workload:
            traffic = tj.get("dram_bytes_per_launch")
    except Exception:
        pass
    roof = {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf.value or NOMINAL_FP64_TFLOPS,
            "unit": "TFLOP/s", "frac": achieved_tf / (peak_tf.value or NOMINAL_FP64_TFLOPS),
            "peak_source": "measured DFMA-chain microbenchmark in this run (rwrt_measure_fp64_peak); nominal 37.2",
            "flop_per_step": FLOP_PER_STEP, "kernel_ms": kms, "traffic": traffic,
            "kernel": "rk45_integrate_coop (+ rk45_prepare and the expand_records transpose inside the same event pair)",
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
                         "oracle C restatement, single thread"}
    line = {"metric": "accepted RK45 ray-steps/s (fp64)", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_max / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{a.workload}: Clim {360.0 / nlon:g}deg background, {nsource} sources x k=1..{nzwn} x 3 roots = {N} "
                                   f"ray slots per GPU, {wl['days']:.0f} d, tstep 2 h, nt={nt}, rtol=atol=1e-6" + (f", {nmember} ensemble members" if nmember > 1 else ""),
                       "mode": "strict" if a.strict else "fast", "accepted_steps_per_step_per_gpu": steps_per_launch,
                       "l2": f"flushed between timed iterations (256 MB write); outputs ({7 * nt * N * 8 / 1e9:.1f} GB) exceed L2 anyway",
                       "timed_region": "rwrt_ray_init + rwrt_ray_run_rk45 (fields resident in HBM)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line), flush=True)
    L.rwrt_destroy(h)
    if world > 1:
        dist.destroy_process_group()


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
                                      "--format=csv,noheader,nounits", "-lms", "100", "-i", str(dev.index or 0)],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=lambda: [link_rows.append(ln.strip()) for ln in link_proc.stdout], daemon=True).start()
    except Exception:
        link_proc = None
    reps = max(1, min(a.steps, 8))
    t0 = time.perf_counter()
    steps = 0
    per_rep = []
    for _ in range(reps):
        t_rep = time.perf_counter()
        call()                               # synchronises before returning
        per_rep.append(round(1e3 * (time.perf_counter() - t_rep), 1))
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
    p.add_argument("--cpu-sources", type=int, default=4000)   # ~14 s of single-thread CPU work on the GPU box
    p.add_argument("--ref-sources", type=int, default=0)
    a = p.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
