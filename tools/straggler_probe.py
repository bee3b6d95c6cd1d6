#!/usr/bin/env python
"""Per-warp time spread of a small batch: each 32-env group of a 4096-env rollout timed alone."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T_
from lerobot_mujoco_sim2real_b200 import _lib
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv

tables = builtin_tables()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
for dtype in ("float64",):
    for split in ("0", "1"):
        env = SOARM101VecEnv(tables=tables, num_envs=32, dtype=dtype)
        env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_TEAM if split == "1" else T_.FAMILY_ONEWARP)
        ts, lim, newt = [], [], []
        for off in range(0, N, 32):
            spec = env.make_spec("random", 7, off)
            def run():
                _lib.check(_lib.lib().so101_batch_rollout(env._h, C.byref(spec), 100, env.frame_skip, None, 0, env._stream()))
            run(); torch.cuda.synchronize()
            env.stats()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            st = env.stats()
            ts.append(e0.elapsed_time(e1)); lim.append(st["limit_steps"] if "limit_steps" in st else -1)
            newt.append(st["newton_iters"] / st["physics_steps"])
        ts = np.array(ts); lim = np.array(lim)
        print(f"{dtype} split={split}: per-warp ms min {ts.min():.3f} median {np.median(ts):.3f} p90 {np.quantile(ts,0.9):.3f} "
              f"max {ts.max():.3f}; limit-steps/warp median {np.median(lim):.0f} max {lim.max()}  corr(t,lim)={np.corrcoef(ts,lim)[0,1]:.2f}")
        order = np.argsort(ts)[-5:]
        print("   slowest:", [(int(o), round(float(ts[o]), 3), int(lim[o]), round(newt[o], 3)) for o in order])
