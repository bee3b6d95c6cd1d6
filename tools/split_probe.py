#!/usr/bin/env python
"""SPLIT (warp-pair) vs one-warp kernels: timing and bitwise agreement on the same rollout."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T_
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv

tables = builtin_tables()
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "1024,4096,9472,18944".split(","))]
for dtype in ("float64", "float32"):
    for n in sizes:
        res = {}
        for split in ("0", "1"):
            env = SOARM101VecEnv(tables=tables, num_envs=n, dtype=dtype)
            env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_TEAM if split == "1" else T_.FAMILY_ONEWARP)
            env.rollout_discard(2, "random")
            torch.cuda.synchronize()
            best = 1e30
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                env.rollout_discard(100, "random", seed=7)
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            q, qd, w = env.get_state()
            res[split] = (best, q.clone(), qd.clone(), w.clone(), env.flags().clone())
            del env
        same = all(torch.equal(res["0"][k], res["1"][k]) for k in range(1, 5))
        dq = (res["0"][1] - res["1"][1]).abs().max().item()
        print(f"{dtype} n={n:6d}: one-warp {res['0'][0]:8.3f} ms  split {res['1'][0]:8.3f} ms  "
              f"speedup {res['0'][0]/res['1'][0]:.3f}  bitwise_equal={same} max|dq|={dq:.3e}", flush=True)
