"""Quanta x windows bisection on a small case (each configuration in its own process with a timeout)."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(tag):
    import rwrt_b200  # noqa: F401
    from rwrt_b200 import synthetic
    from rwrt_b200.wr import WR
    u, v = synthetic.clim(360, 181)
    sl, sa = synthetic.random_sources(int(os.environ.get("NS", "3000")), 7)
    zwn = np.arange(1.0, 7.0)
    w = WR(len(zwn), len(sl), 7200.0, float(os.environ.get("DAYS", "12")) * 86400.0, 0.0, nx=360, ny=181)
    w.bs.lon, w.bs.lat = synthetic.grid_coords(360, 181)
    w.bs.u, w.bs.v = u, v
    w.bs.ready(xcyclic=True, keep_fields18=False, keep_diag=False)
    w.set_zwn(zwn)
    w.set_source_array(sl, sa)
    w.ray_run(mode="numpy", inte_method="rk45", to_host=False)   # rwrt_ray_run_rk45: honours RWRT_WINDOWS
    out = np.stack([getattr(w, n) for n in ("rlon", "rlat", "rzwn", "rmwn", "ramp", "rug", "rvg")])
    np.savez(f"/tmp/dq_{tag}.npz", out=out, acc=w.accepted, status=w.status)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        return child(sys.argv[2])
    cfgs = {"ref": dict(RWRT_QUANTUM_SHIFT="0"), "q": dict(), "q_w3": dict(RWRT_WINDOWS="3"), "w3": dict(RWRT_QUANTUM_SHIFT="0", RWRT_WINDOWS="3"),
            "qmin2_w3": dict(RWRT_QUANTUM_MIN="2", RWRT_WINDOWS="3"), "q_w4_b2": dict(RWRT_BLOCKS_PER_SM="2", RWRT_WINDOWS="4")}
    for tag, env in cfgs.items():
        try:
            r = subprocess.run([sys.executable, __file__, "child", tag], env=dict(os.environ, **env), timeout=90, capture_output=True, text=True)
            if r.returncode != 0:
                print(tag, "FAILED", r.stderr[-300:])
        except subprocess.TimeoutExpired:
            print(tag, "TIMEOUT (hang)")
    ref = np.load("/tmp/dq_ref.npz")
    for tag in cfgs:
        if tag == "ref" or not os.path.exists(f"/tmp/dq_{tag}.npz"):
            continue
        g = np.load(f"/tmp/dq_{tag}.npz")
        same = (g["out"] == ref["out"]) | (np.isnan(g["out"]) & np.isnan(ref["out"]))
        bad = ~same.all(axis=(0, 1))
        print(tag, "slots differing:", int(bad.sum()), "of", bad.size, " acc differ:", int((g["acc"] != ref["acc"]).sum()))
        if bad.any():
            idx = np.argwhere(bad)[:5]
            for ir, s, z in idx:
                rows = np.where(~same[:, :, ir, s, z].all(axis=0))[0]
                print("   slot", (ir, s, z), "first differing row", rows[:3], "acc", g["acc"][ir, s, z], ref["acc"][ir, s, z], "status", g["status"][ir, s, z], ref["status"][ir, s, z])


if __name__ == "__main__":
    main()
