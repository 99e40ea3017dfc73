"""Timing of the fixed-step RK4 path (rwrt_ray_run_rk4; the reference's default inte_method='') on C2."""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwrt_b200  # noqa
from rwrt_b200 import lib, synthetic
L = lib.load(); lib.require_gpu()
nlon, nlat = 360, 181
u, v = synthetic.clim(nlon, nlat); lon, lat = synthetic.grid_coords(nlon, nlat)
sl, sa = synthetic.all_gridpoint_sources(nlon, nlat, 88.0)
zwn = np.arange(1.0, 7.0); ns, nz = len(sl), len(zwn); nt = 361; N = 3 * ns * nz
dev = torch.device("cuda", 0)
h = lib.vp(); lib.check(L.rwrt_create(0, nlon, nlat, 1, 1, C.byref(h)))
d = {k: torch.from_numpy(np.ascontiguousarray(x)).to(dev) for k, x in dict(u=u, v=v, lon=lon, lat=lat, cl=np.cos(lat), sl=np.sin(lat), slon=np.radians(sl), slat=np.radians(sa), zwn=zwn).items()}
sp = lib.stream_ptr()
lib.check(L.rwrt_bs_ready(h, lib.ptr(d["u"]), lib.ptr(d["v"]), lib.ptr(d["lon"]), lib.ptr(d["lat"]), lib.ptr(d["cl"]), lib.ptr(d["sl"]), None, None, sp))
out = torch.empty((7, nt, N), dtype=torch.float64, device=dev)
for flags, name in ((lib.MODE_FAST, "fast"), (lib.MODE_STRICT, "strict")):
    ts = []
    for i in range(3):
        lib.check(L.rwrt_ray_init(h, ns, lib.ptr(d["slon"]), lib.ptr(d["slat"]), None, nz, lib.ptr(d["zwn"]), 0.0, nt, lib.ptr(out), sp))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.check(L.rwrt_ray_run_rk4(h, ns, None, nz, nt, 7200.0, 0.1, flags, lib.ptr(out), None, sp))
        e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    alive = int((~torch.isnan(out[3][1:])).sum().item())
    print(f"rk4 {name}: {min(ts):.1f} ms per pass, {alive} alive ray-steps -> {alive / min(ts) * 1e3:.3g} RK4 steps/s ({N * (nt - 1) / min(ts) * 1e3:.3g} slot-steps/s)")
