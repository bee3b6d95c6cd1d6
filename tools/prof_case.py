#!/usr/bin/env python
"""One short fused rollout for ncu: python tools/prof_case.py <dtype> <n_envs> <T> [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv

dtype, n, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
rows = len(sys.argv) > 4
env = SOARM101VecEnv(tables=builtin_tables(), num_envs=n, dtype=dtype)
for _ in range(2):
    if rows:
        env.rollout(T, "random")
    else:
        env.rollout_discard(T, "random")
torch.cuda.synchronize()
print("ok", env.stats())
