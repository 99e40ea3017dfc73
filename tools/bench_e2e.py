"""Sweep of the host-buffer entry (rwrt_real2d_host) on BASELINE configs[1]: one process, one set of pinned buffers,
several settings of the environment knobs the library reads at call time.

    python tools/bench_e2e.py "RWRT_HOST_WINDOWS=12" "RWRT_HOST_THREADS=14,RWRT_HX_ISA=avx2" ...
"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    import rwrt_b200  # noqa: F401
    from rwrt_b200 import build as rbuild, lib
    rbuild.build()
    L = lib.load()
    lib.require_gpu()
    wl = bench.workload(os.environ.get("RWRT_BENCH_WORKLOAD", "c2"))
    slon, slat = np.radians(wl["src_lon_deg"]), np.radians(wl["src_lat_deg"])
    nsource, nzwn = len(slon), len(wl["zwn"])
    tstep, ttotal = 7200.0, wl["days"] * 86400.0
    nt = int(ttotal / tstep) + 1
    N = 3 * nsource * nzwn
    nlat, nlon = len(wl["lat"]), len(wl["lon"])
    h = lib.vp()
    lib.check(L.rwrt_create(0, nlon, nlat, 1, 1, C.byref(h)))
    host = torch.empty(7 * nt * N, dtype=torch.float64, pin_memory=True)
    acc_h = torch.empty(N, dtype=torch.int64, pin_memory=True)
    tot_h = torch.zeros(1, dtype=torch.int64, pin_memory=True)
    pin = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in
           dict(u=wl["u"], v=wl["v"], lon=wl["lon"], lat=wl["lat"], cl=np.cos(wl["lat"]), sl=np.sin(wl["lat"]),
                slon=slon, slat=slat, zwn=wl["zwn"]).items()}

    def call():
        lib.check(L.rwrt_real2d_host(h, lib.ptr(pin["u"]), lib.ptr(pin["v"]), lib.ptr(pin["lon"]), lib.ptr(pin["lat"]),
                                     lib.ptr(pin["cl"]), lib.ptr(pin["sl"]), nsource, lib.ptr(pin["slon"]),
                                     lib.ptr(pin["slat"]), None, nzwn, lib.ptr(pin["zwn"]), 0.0, nt, tstep, ttotal,
                                     1e-6, 1e-6, 1e-3, 0.1, lib.MODE_FAST, lib.ptr(host), lib.ptr(acc_h), None,
                                     lib.ptr(tot_h), None), "rwrt_real2d_host")

    for setting in [""] + sys.argv[1:]:
        env = dict(kv.split("=") for kv in setting.split(",") if kv)
        for k, v in env.items():
            os.environ[k] = v
        call()      # warm-up: staging buffers of this setting
        ts, tls = [], []
        for _ in range(int(os.environ.get("RWRT_BENCH_CALLS", "4"))):
            t0 = time.perf_counter()
            call()
            ts.append(1e3 * (time.perf_counter() - t0))
            tls.append(bench.host_timeline(L, h))
        hs = (C.c_double * 5)()
        L.rwrt_last_host_stats(h, hs)
        print(f"{setting or 'default':50s} ms/pass min {min(ts):7.1f} median {sorted(ts)[len(ts)//2]:7.1f} all {[round(t) for t in ts]}   "
              f"d2h {hs[0]/1e9:.2f} GB, {int(hs[2])} windows ({int(hs[4])} direct), {int(hs[3])} threads, steps {int(tot_h.item())}", flush=True)
        print("     fastest call:", tls[int(np.argmin(ts))], flush=True)
        print("     slowest call:", tls[int(np.argmax(ts))], flush=True)
        for k in env:
            del os.environ[k]
    L.rwrt_destroy(h)


if __name__ == "__main__":
    main()
