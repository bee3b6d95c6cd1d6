#!/usr/bin/env python
"""A few launches of so101_koopman_mpc_step for ncu: python tools/prof_mpc.py [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
km = KoopmanModel.from_npz(os.path.join(ROOT, "tests", "golden", "koopman_dkuc.npz"))
g = torch.Generator(device="cuda").manual_seed(0)
obs = (torch.rand((8, n), generator=g, device="cuda", dtype=torch.float32) - 0.5)
uff = torch.zeros((n, 4, 5), dtype=torch.float64, device="cuda")
u_prev = torch.zeros((5, n), dtype=torch.float64, device="cuda")
ctrl = torch.empty((5, n), dtype=torch.float64, device="cuda")
for _ in range(3):
    km.mpc_step(obs, True, uff, 1, u_prev, ctrl, None, 10, "delta_mpc", 0.5)
torch.cuda.synchronize()
print("ok", float(ctrl.abs().max()))
