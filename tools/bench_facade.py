"""One-shot timing of the facade on BASELINE configs[1] (what a user of main_wr.real2d_hnf sees, minus NetCDF):
WR(...) -> bs.ready -> ray_run(mode='numpy', inte_method='rk45') with the seven result arrays in host memory.

    python tools/bench_facade.py [pin]        # pin: result arrays in pinned memory (WR.pin_results = True)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    import rwrt_b200  # noqa: F401
    from rwrt_b200.wr import WR
    pin = "pin" in sys.argv[1:]
    wl = bench.workload("c2")
    nlon, nlat = len(wl["lon"]), len(wl["lat"])
    torch.zeros(1, device="cuda")           # CUDA context outside the timed region
    for rep in range(2):
        t0 = time.perf_counter()
        w = WR(len(wl["zwn"]), len(wl["src_lon_deg"]), 7200.0, wl["days"] * 86400.0, 0.0, nx=nlon, ny=nlat, device=0)
        w.pin_results = pin
        w.bs.lon, w.bs.lat, w.bs.u, w.bs.v = wl["lon"], wl["lat"], wl["u"], wl["v"]
        w.bs.ready(xcyclic=True)
        w.set_zwn(wl["zwn"])
        w.set_source_array(wl["src_lon_deg"], wl["src_lat_deg"])
        t1 = time.perf_counter()
        w.ray_run(mode="numpy", root_method="numpy", inte_method="rk45")
        t2 = time.perf_counter()
        alive = int(np.isfinite(w.rlat[-1]).sum())
        t3 = time.perf_counter()
        print(f"run {rep} ({'pinned' if pin else 'pageable'} result arrays): setup {t1 - t0:.2f} s, ray_run {t2 - t1:.2f} s "
              f"({w.total_accepted / (t2 - t1):.3g} accepted steps/s with the results on the host), first use of rlat "
              f"{t3 - t2:.2f} s; {alive} rays alive at the end", flush=True)
        w.clean()
        del w


if __name__ == "__main__":
    main()
