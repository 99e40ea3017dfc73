"""Timing of the SHSF path for the ensemble case (BASELINE configs[3]: 1000 members on the 1-degree grid):
    python tools/bench_shsf.py            (GPU box)
Prints ms per batch, achieved FP64 GFLOP/s of the four products, and the oracle's time for ONE member on the host."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rwrt_b200  # noqa
from rwrt_b200.SHSF import SHSFPlan
from oracle import shsf as O

for n, L, nm in ((180, 44, 1000), (180, 89, 1000), (720, 359, 16)):
    t0 = time.perf_counter()
    plan = SHSFPlan(n, L, 2, 1, 0)
    t_plan = time.perf_counter() - t0
    g = torch.randn((nm, n, 2 * n), dtype=torch.float64, device="cuda")
    out = plan.apply_device(g)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        plan.apply_device(g, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    Lc = min(L, n // 2 - 1); nl = Lc + 1; cw = 2 * nl; no = 2 * L + 2
    flop = 2.0 * nm * (n * 2 * n * cw + cw * nl * n + cw * no * nl + no * no * cw)
    line = f"n={n} L={L} members={nm}: plan {t_plan*1e3:.0f} ms, apply {ms:.2f} ms, {flop/ms/1e6:.0f} GFLOP/s fp64"
    if n <= 180:
        x = g[0].cpu().numpy()
        t0 = time.perf_counter(); O.shsf(x, L); t1 = time.perf_counter() - t0
        line += f"; oracle (numpy, 1 member) {t1*1e3:.0f} ms -> {t1*nm:.1f} s for the batch"
    print(line)
    plan.close()
