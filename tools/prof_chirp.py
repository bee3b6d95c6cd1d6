#!/usr/bin/env python
"""One long chirp rollout (many envs reach the table) for ncu: python tools/prof_chirp.py <dtype> <n_envs> <T>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
dtype, n, Tn = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
env = SOARM101VecEnv(tables=builtin_tables(), num_envs=n, dtype=dtype)
env.set_option(T.OPT_KERNEL_FAMILY, T.FAMILY_ONEWARP)
env.set_option(T.OPT_SLICED, 2)
for _ in range(2):
    env.rollout_discard(Tn, "chirp", seed=42)
torch.cuda.synchronize()
fl = env.flags()
print("ok", env.stats(), "in contact", float((fl & T.FLAG_CONTACT).ne(0).double().mean()))
