#!/usr/bin/env python
"""One batched IK launch for ncu / timing: python tools/ik_case.py <n_tracks> [pose]
Tracks = the reference's four curves (300 way-points) with random centre shifts; `pose` adds an orientation target
(the site orientation at the first way-point's position-only solution, so most way-points stay reachable)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables
from lerobot_mujoco_sim2real_b200.TrajectoryGenerator import CartesianTrajectoryGenerator, reference_curve

n = int(sys.argv[1])
pose = len(sys.argv) > 2
gen = CartesianTrajectoryGenerator(tables=builtin_tables())
base = np.stack([reference_curve(nm, ix)[0] for nm in ("Fig8", "Circle") for ix in (0, 1)])
xyz = torch.as_tensor(base[np.arange(n) % 4] + np.random.default_rng(0).uniform(-0.03, 0.03, (n, 1, 3))).cuda()
quat = torch.tensor([0.9997, 0.0243, 0.0, 0.0], dtype=torch.float64).cuda() if pose else None
import time
t0 = time.time()
while time.time() - t0 < 0.5:      # an idle B200 sits at low clocks: keep it busy before timing a latency-bound kernel
    gen.solve_tracks(xyz, quat)
    torch.cuda.synchronize()
ms = 1e30
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); q, st = gen.solve_tracks(xyz, quat); b.record(); torch.cuda.synchronize()
    ms = min(ms, a.elapsed_time(b))
it = (st >> 8).double().mean().item()
print(f"ok n={n} pose={pose} {ms:.3f} ms, {n * 300 / ms * 1e-3:.2f} M way-points/s, success {(st & 1).double().mean().item():.4f}, "
      f"iterations per way-point {it:.2f}, {n * 300 * (it + 1) / ms * 1e-3:.1f} M FK+solve iterations/s")
