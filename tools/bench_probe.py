#!/usr/bin/env python
"""The bench workload (4096 envs f64, 100 control steps, rows written, L2 flushed between launches) with the contact path on
and off, and with the kernel families forced - what the headline number is made of."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
t = builtin_tables()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for name, hulls, rows in (("contact off, no rows", None, False), ("contact off, rows", None, True), ("contact on, no rows", "auto", False), ("contact on, rows", "auto", True)):
    env = SOARM101VecEnv(tables=t, num_envs=n, hulls=hulls)
    run = (lambda: env.rollout(100, "random", seed=42)) if rows else (lambda: env.rollout_discard(100, "random", seed=42))
    for _ in range(3): run()
    ts = []
    for rep in range(10):
        flush.fill_(rep)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print(f"{name:22s}: median {ts[len(ts)//2]:.3f} ms  min {ts[0]:.3f}  max {ts[-1]:.3f}", flush=True)
