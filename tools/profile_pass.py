"""Profiling driver (run under ncu): one C2 pass with device-resident ray records, the on-demand transpose, gridding from the
records, and a fused grid-only sweep -- every round-2 kernel once after a warm-up pass.

    ncu --set full --clock-control none --import-source on -k regex:'rk45_integrate_coop|expand_records|grid_records' -s 1 -c 4 \
        -o gpurun_out/r2_prof python tools/profile_pass.py
"""
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
    wl = bench.workload(os.environ.get("WL", "c2"))
    dev = torch.device("cuda", 0)
    h = lib.vp()
    lib.check(L.rwrt_create(0, wl["nlon"], wl["nlat"], 1, 1, C.byref(h)))
    sp = lib.stream_ptr()
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
         dict(u=wl["u"], v=wl["v"], lon=wl["lon"], lat=wl["lat"], cl=np.cos(wl["lat"]), sl=np.sin(wl["lat"]), zwn=wl["zwn"],
              slon=np.radians(wl["src_lon_deg"]), slat=np.radians(wl["src_lat_deg"])).items()}
    lib.check(L.rwrt_bs_ready(h, lib.ptr(d["u"]), lib.ptr(d["v"]), lib.ptr(d["lon"]), lib.ptr(d["lat"]), lib.ptr(d["cl"]), lib.ptr(d["sl"]),
                              None, None, sp))
    ns, nz = len(wl["src_lon_deg"]), len(wl["zwn"])
    nt = int(wl["days"] * 12) + 1
    N = 3 * ns * nz
    row0 = torch.empty((7, N), dtype=torch.float64, device=dev)
    tot = torch.zeros(1, dtype=torch.int64, device=dev)
    tstep, ttotal = 7200.0, wl["days"] * 86400.0

    def records_pass():
        lib.check(L.rwrt_ray_init(h, ns, lib.ptr(d["slon"]), lib.ptr(d["slat"]), None, nz, lib.ptr(d["zwn"]), 0.0, 1, lib.ptr(row0), sp))
        lib.check(L.rwrt_ray_run_rk45_records(h, ns, None, nz, nt, tstep, ttotal, 1e-6, 1e-6, 1e-3, 0.1, 0, lib.ptr(row0), None, None, None,
                                              lib.ptr(tot), sp))
    records_pass()                                   # warm-up: integrate<OUT_ROWS>  (skipped by -s 1)
    records_pass()                                   # 1: integrate<OUT_ROWS>
    rows = min(nt - 1, int(os.environ.get("EXPAND_ROWS", str(nt - 1))))
    dst = torch.empty((7, rows, N), dtype=torch.float64, device=dev)
    lib.check(L.rwrt_records_expand(h, 1, 1 + rows, lib.ptr(dst), sp))          # 2: expand_records
    cell = (1.0, 1.0)
    ncx, ncy = gridding.grid_shape(*cell)
    g = gridding.new_grids(*cell, dev)
    lib.check(L.rwrt_grid_accumulate_records(h, cell[0], cell[1], ncx, ncy, lib.ptr(g[0]), lib.ptr(g[1]), lib.ptr(g[2]), sp))   # 3: grid_records
    del dst
    g2 = gridding.new_grids(*cell, dev)
    lib.check(L.rwrt_ray_sweep_grid(h, ns, lib.ptr(d["slon"]), lib.ptr(d["slat"]), None, nz, lib.ptr(d["zwn"]), 0.0, nt, tstep, ttotal, 1e-6,
                                    1e-6, 1e-3, 0.1, 0, cell[0], cell[1], ncx, ncy, lib.ptr(g2[0]), lib.ptr(g2[1]), lib.ptr(g2[2]), None, None,
                                    lib.ptr(tot), sp))                           # 4: integrate<OUT_GRID>
    torch.cuda.synchronize()
    assert torch.equal(g, g2), "fused sweep and gridding from the records disagree"
    print("steps", int(tot.item()), "grid samples", int(g[0].sum().item()))
    L.rwrt_destroy(h)


if __name__ == "__main__":
    main()
