#!/usr/bin/env python
"""One wave of the one-warp kernels at different block sizes: how much faster does a block step with fewer warps on its SM?
argv: dtype kind T; prints ms per launch for 148 blocks of each block size (contact on)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv

dtype, kind, Tn = sys.argv[1], sys.argv[2], int(sys.argv[3])
t = builtin_tables()
for hulls in (None, "auto"):
    for blk in ([32, 64, 128, 256] if dtype == "float64" else [32, 64, 128, 256, 512]):
        n = 148 * blk
        env = SOARM101VecEnv(tables=t, num_envs=n, dtype=dtype, hulls=hulls)
        env.set_option(T.OPT_KERNEL_FAMILY, T.FAMILY_ONEWARP)
        env.set_option(T.OPT_BLOCK, blk)
        env.rollout_discard(2, kind)
        torch.cuda.synchronize()
        best = 1e30
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); env.rollout_discard(Tn, kind, seed=42 + rep); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        fl = env.flags()
        print(f"{dtype} {kind} T={Tn} contact={'on' if hulls else 'off'} blk={blk} n={n}: {best:.3f} ms, "
              f"in contact {float((fl & T.FLAG_CONTACT).ne(0).double().mean()):.4f}", flush=True)
