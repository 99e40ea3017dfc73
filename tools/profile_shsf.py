import os, sys
sys.path.insert(0, "/root/repo")
import torch
import rwrt_b200  # noqa
from rwrt_b200.SHSF import SHSFPlan
plan = SHSFPlan(180, 89, 2, 1, 0)
g = torch.randn((1000, 180, 360), dtype=torch.float64, device="cuda")
out = plan.apply_device(g)
plan.apply_device(g, out)
torch.cuda.synchronize()
plan.close()
