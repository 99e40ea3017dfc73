import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import rwrt_b200
from rwrt_b200 import lib, synthetic
import test_gpu_fullsize as T
nlon, nlat = 360, 181
slon, slat = synthetic.all_gridpoint_sources(nlon, nlat, 88.0)
stride = int(os.environ.get("STRIDE", "1"))
slon, slat = slon[::stride], slat[::stride]
zwn = np.arange(1.0, 7.0)
L, h, d, dev = T._setup(nlon, nlat, slon, slat, zwn)
ns, nz = len(slon), len(zwn)
out0, acc0, st0, tot0, nt, N = T._run(L, h, d, dev, ns, nz)
print("base", T._checksum(out0), tot0)
for w in (2, 5, 9, 3, 7, 11):
    os.environ["RWRT_WINDOWS"] = str(w)
    out, acc, st, tot, _, _ = T._run(L, h, d, dev, ns, nz)
    same = (out == out0) | (torch.isnan(out) & torch.isnan(out0))
    bad = (~same).any(dim=0).any(dim=0)
    nb = int(bad.sum().item())
    print("windows", w, T._checksum(out), tot, "bad slots", nb)
    if nb:
        idx = torch.nonzero(bad)[:5, 0].tolist()
        for s in idx:
            rows = torch.nonzero((~same[:, :, s]).any(dim=0))[:, 0]
            print("  slot", s, "first bad row", int(rows[0]), "n bad rows", len(rows), "acc", int(acc0[s]), int(acc[s]),
                  "vals", out0[0, int(rows[0]), s].item(), out[0, int(rows[0]), s].item())
            r0 = int(rows[0])
            print("     prev row lon/lat deg", np.degrees(out0[0, r0 - 1, s].item()) % 360, np.degrees(out0[1, r0 - 1, s].item()),
                  " this row base", np.degrees(out0[0, r0, s].item()) % 360, np.degrees(out0[1, r0, s].item()), " diff lon/lat", out[0, r0, s].item() - out0[0, r0, s].item(), out[1, r0, s].item() - out0[1, r0, s].item())
    del out, same
del os.environ["RWRT_WINDOWS"]
