#!/usr/bin/env python
"""Block-size sweep for the one-warp kernels (SO101_OPT_BLOCK option): wave quantisation vs warps per SM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T_
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
tables = builtin_tables()
cases = [("float64", 131072, [256, 224, 192]), ("float64", 65536, [256, 224, 192]), ("float64", 32768, [256, 224, 192, 128]),
         ("float64", 16384, [128, 96, 64]),
         ("float32", 131072, [512, 480, 448, 416]), ("float32", 65536, [512, 480, 448, 256]), ("float32", 32768, [256, 224, 448])]
if len(sys.argv) > 1:
    cases = eval(sys.argv[1])
for dtype, n, blks in cases:
    for blk in blks:
        env = SOARM101VecEnv(tables=tables, num_envs=n, dtype=dtype)
        env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_ONEWARP)
        env.set_option(T_.OPT_BLOCK, blk)
        T = 40
        env.rollout_discard(2, "random"); torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); env.rollout_discard(T, "random"); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        warps = (n + 31) // 32; blocks = (warps * 32 + blk - 1) // blk
        print(f"{dtype} n={n:7d} blk={blk:4d} blocks={blocks:5d} ({blocks/148:.2f}/SM): {best:8.3f} ms  {n*T*10/best/1e6:8.1f} G/1000 physics-steps/s", flush=True)
        del env
