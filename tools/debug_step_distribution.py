"""Distribution of accepted steps per ray (C5, 1/8 of the sources): how long are the longest rays compared with the mean?"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import rwrt_b200  # noqa: F401
    from rwrt_b200 import gridding, synthetic
    from rwrt_b200.wr import WR
    wl = bench.workload(os.environ.get("WL", "c5"))
    w = WR(len(wl["zwn"]), len(wl["src_lon_deg"]), 7200.0, wl["days"] * 86400.0, 0.0, nx=wl["nlon"], ny=wl["nlat"])
    w.bs.lon, w.bs.lat, w.bs.u, w.bs.v = wl["lon"], wl["lat"], wl["u"], wl["v"]
    w.bs.ready(xcyclic=True, keep_fields18=False, keep_diag=False)
    w.set_zwn(wl["zwn"])
    w.set_source_array(wl["src_lon_deg"], wl["src_lat_deg"])
    gridding.sweep(w, 0.25, 0.25)
    acc = w.accepted[w.status != 0].astype(np.float64)
    print("live rays", acc.size, "mean steps", acc.mean(), "median", np.median(acc), "max", acc.max(), "p99", np.percentile(acc, 99), "p99.9",
          np.percentile(acc, 99.9), "p99.99", np.percentile(acc, 99.99))
    for thr in (1.5, 2, 3, 5, 10):
        print(f"  rays with > {thr} x mean steps: {(acc > thr * acc.mean()).sum()}")
    st = w.status[w.status != 0]
    print("status counts", {int(c): int((st == c).sum()) for c in np.unique(st)})


if __name__ == "__main__":
    main()
