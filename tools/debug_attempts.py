"""C2: distribution of attempts (accepted + rejected RK steps) per ray -- how long is the longest serial chain?"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    import rwrt_b200  # noqa: F401
    from rwrt_b200 import lib
    L = lib.load()
    wl = bench.workload("c2", int(os.environ.get("SHIFT", "0")))
    if os.environ.get("SHIFT", "0") != "0":
        os.environ["RWRT_BENCH_SHIFT_SOURCES"] = "1"
        wl = bench.workload("c2", int(os.environ["SHIFT"]))
    dev = torch.device("cuda", 0)
    h = lib.vp()
    lib.check(L.rwrt_create(0, wl["nlon"], wl["nlat"], 1, 1, C.byref(h)))
    sp = lib.stream_ptr()
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(u=wl["u"], v=wl["v"], lon=wl["lon"], lat=wl["lat"], cl=np.cos(wl["lat"]), sl=np.sin(wl["lat"]), zwn=wl["zwn"],
              slon=np.radians(wl["src_lon_deg"]), slat=np.radians(wl["src_lat_deg"])).items()}
    lib.check(L.rwrt_bs_ready(h, lib.ptr(d["u"]), lib.ptr(d["v"]), lib.ptr(d["lon"]), lib.ptr(d["lat"]), lib.ptr(d["cl"]), lib.ptr(d["sl"]),
                              None, None, sp))
    ns, nz, nt = len(wl["src_lon_deg"]), len(wl["zwn"]), 361
    N = 3 * ns * nz
    row0 = torch.empty((7, N), dtype=torch.float64, device=dev)
    acc = torch.zeros(N, dtype=torch.int64, device=dev)
    att = torch.zeros(N, dtype=torch.int64, device=dev)
    lib.check(L.rwrt_ray_init(h, ns, lib.ptr(d["slon"]), lib.ptr(d["slat"]), None, nz, lib.ptr(d["zwn"]), 0.0, 1, lib.ptr(row0), sp))
    lib.check(L.rwrt_ray_run_rk45_records(h, ns, None, nz, nt, 7200.0, 30 * 86400.0, 1e-6, 1e-6, 1e-3, 0.1, 0, lib.ptr(row0), lib.ptr(acc),
                                          lib.ptr(att), None, None, sp))
    torch.cuda.synchronize()
    times = (C.c_double * 4)()
    L.rwrt_last_run_times(h, times)
    a = att.cpu().numpy()
    a = a[a > 0].astype(np.float64)
    print(f"integrate {times[1]:.2f} ms; live rays {a.size}; attempts per ray: mean {a.mean():.0f}, p99 {np.percentile(a, 99):.0f}, p99.99 {np.percentile(a, 99.99):.0f}, "
          f"max {a.max():.0f}; total {a.sum():.4g}; lanes {148 * 3 * 128}: throughput bound = total / lanes = {a.sum() / (148 * 3 * 128):.0f} attempts per lane; "
          f"longest ray = {a.max() / (a.sum() / (148 * 3 * 128)):.2f} of that")
    top = np.sort(a)[-8:]
    print("largest:", top)
    L.rwrt_destroy(h)


if __name__ == "__main__":
    main()
