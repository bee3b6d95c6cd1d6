#!/usr/bin/env python
"""One 32-env group of the benchmark rollout alone (team kernel), for ncu: python tools/prof_group.py <group> [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T_, _lib
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
g = int(sys.argv[1]); seed = int(sys.argv[2]) if len(sys.argv) > 2 else 42
env = SOARM101VecEnv(tables=builtin_tables(), num_envs=32, dtype="float64")
env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_TEAM)
spec = env.make_spec("random", seed, g * 32)
for _ in range(2):
    _lib.check(_lib.lib().so101_batch_rollout(env._h, C.byref(spec), 100, env.frame_skip, None, 0, env._stream()))
torch.cuda.synchronize()
print("ok", env.stats())
