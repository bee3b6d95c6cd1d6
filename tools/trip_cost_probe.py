#!/usr/bin/env python
"""What the lookout warp's tripwire costs the team (and the one-warp kernel): same rollout with ntrip = 0."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
for dtype in ("float64", "float32"):
    for n in (4096, 131072):
        for trip in (True, False):
            t = builtin_tables()
            if not trip: t.ntrip = 0
            env = SOARM101VecEnv(tables=t, num_envs=n, dtype=dtype)
            T = 100 if n <= 9472 else 20
            env.rollout_discard(2, "random"); torch.cuda.synchronize()
            best = 1e30
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); env.rollout_discard(T, "random", seed=7); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            print(f"{dtype} n={n:7d} tripwire={'on ' if trip else 'off'}: {best:8.3f} ms  {n*T*10/best/1e6:8.2f} G physics-steps/s", flush=True)
