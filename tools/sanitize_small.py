"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rwrt_b200  # noqa
from rwrt_b200 import synthetic
from rwrt_b200.wr import WR
from rwrt_b200.SHSF import SHSF

u, v = synthetic.clim(144, 73)
lon, lat = synthetic.grid_coords(144, 73)
sl, sa = synthetic.random_sources(int(os.environ.get("NSRC", "400")), 11)
w = WR(3, len(sl), 7200.0, 4 * 86400.0, 0.0, nx=144, ny=73, device=0)
w.bs.lon, w.bs.lat, w.bs.u, w.bs.v = lon, lat, u, v
w.bs.ready(xcyclic=True)
w.set_zwn(np.array([1.0, 2.0, 4.0]))
w.set_source_array(sl, sa)
for windows in ("1", "3"):
    os.environ["RWRT_WINDOWS"] = windows
    w.ray_run(mode="numpy", root_method="numpy", inte_method="rk45")
    print("windows", windows, "accepted", w.total_accepted)
w.ray_run(mode="numpy", root_method="numpy", inte_method="")
out = SHSF(np.random.default_rng(0).normal(size=(36, 72)), 9)
print("rk4 + shsf ok", out.shape)
