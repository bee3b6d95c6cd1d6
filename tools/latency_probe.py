#!/usr/bin/env python
"""Single-env latency of the drop-in SOARM101Env.step (host numpy in/out) and small-batch step()."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lerobot_mujoco_sim2real_b200 import builtin_tables
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
import ctypes as C
from lerobot_mujoco_sim2real_b200 import _lib, tables as T
for n in (1, 32, 256, 1024, 4096):
    env = SOARM101VecEnv(tables=builtin_tables(), num_envs=n, dtype="float64")
    u = np.zeros((6, n)); obs = np.zeros((8, n), dtype=np.float32)
    env.reset()
    L = _lib.lib()
    for _ in range(20):
        L.so101_batch_step_host(env._h, u.ctypes.data, 6, 10, obs.ctypes.data, None, env._stream())
    t0 = time.perf_counter()
    K = 200
    for _ in range(K):
        L.so101_batch_step_host(env._h, u.ctypes.data, 6, 10, obs.ctypes.data, None, env._stream())
    dt = (time.perf_counter() - t0) / K
    print(f"n={n:5d}: step_host (10 sub-steps, H2D+kernel+D2H+sync) {dt*1e6:8.1f} us  -> {n/dt/1e3:9.1f} k env-steps/s")
