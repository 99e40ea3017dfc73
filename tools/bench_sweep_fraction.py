"""One GPU: grid-only sweep of 1/1, 1/2, 1/4, 1/8 of the C5 sources (strided shards): where does strong scaling lose time?"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    import rwrt_b200  # noqa: F401
    from rwrt_b200 import gridding, lib
    L = lib.load()
    wl = bench.workload("c5-full")
    dev = torch.device("cuda", 0)
    h = lib.vp()
    lib.check(L.rwrt_create(0, wl["nlon"], wl["nlat"], 1, 1, C.byref(h)))
    stream = torch.cuda.current_stream()
    sp = lib.vp(stream.cuda_stream)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(u=wl["u"], v=wl["v"], lon=wl["lon"], lat=wl["lat"], cl=np.cos(wl["lat"]), sl=np.sin(wl["lat"]), zwn=wl["zwn"]).items()}
    lib.check(L.rwrt_bs_ready(h, lib.ptr(d["u"]), lib.ptr(d["v"]), lib.ptr(d["lon"]), lib.ptr(d["lat"]), lib.ptr(d["cl"]), lib.ptr(d["sl"]),
                              None, None, sp))
    slon_all, slat_all = np.radians(wl["src_lon_deg"]), np.radians(wl["src_lat_deg"])
    nz, nt = len(wl["zwn"]), int(wl["days"] * 12) + 1
    cell = (0.25, 0.25)
    ncx, ncy = gridding.grid_shape(*cell)
    g = gridding.new_grids(*cell, dev)
    tot = torch.zeros(1, dtype=torch.int64, device=dev)
    times = (C.c_double * 4)()
    stats = (C.c_double * 3)()
    for world in [int(x) for x in os.environ.get("RWRT_SWEEP_FRACTIONS", "1,2,4,8,16").split(",")]:   # 1/world of the sources
        idx = np.arange(0, len(slon_all), world)
        sl = torch.from_numpy(slon_all[idx].copy()).to(dev)
        sa = torch.from_numpy(slat_all[idx].copy()).to(dev)

        def one():
            g.zero_()
            lib.check(L.rwrt_ray_sweep_grid(h, len(idx), lib.ptr(sl), lib.ptr(sa), None, nz, lib.ptr(d["zwn"]), 0.0, nt, 7200.0,
                                            wl["days"] * 86400.0, 1e-6, 1e-6, 1e-3, 0.1, 0, cell[0], cell[1], ncx, ncy, lib.ptr(g[0]),
                                            lib.ptr(g[1]), lib.ptr(g[2]), None, None, lib.ptr(tot), sp))
        one()
        torch.cuda.synchronize()
        ms = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            one()
            e1.record(stream)
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        L.rwrt_last_run_times(h, times)
        L.rwrt_last_run_stats(h, stats)
        steps = int(tot.item())
        print(f"1/{world:<2d} of the sources: {3 * len(idx) * nz:9d} slots, {int(stats[1]):8d} live rays ({stats[1] / (148 * 3 * 128):5.1f} per lane)  "
              f"pass {np.mean(ms):8.2f} ms  x{world} = {world * np.mean(ms):8.2f}   run {times[0]:8.2f} integrate {times[1]:8.2f} setup {times[2]:6.2f} "
              f"rest {times[3]:6.2f}   {steps / np.mean(ms) / 1e6:7.1f} M steps/s  ({steps / 1e6:.1f} M steps)", flush=True)
    L.rwrt_destroy(h)


if __name__ == "__main__":
    main()
