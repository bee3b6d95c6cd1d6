#!/usr/bin/env python
"""The bench workload (4096 envs f64, 100 control steps) per control seed: launch time and the contact statistics of the batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
env = SOARM101VecEnv(tables=builtin_tables(), num_envs=4096)
env.rollout_discard(100, "random", seed=1)
for seed in range(42, 62):
    best = 1e9
    for rep in range(2):
        env.clear_flags(); env.stats()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.rollout_discard(100, "random", seed=seed); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    st = env.stats(); fl = env.flags()
    print(f"seed {seed}: {best:.3f} ms  in contact at some time {int((fl & T.FLAG_CONTACT).ne(0).sum())}  newton/step {st['newton_iters'] / st['physics_steps']:.4f}  "
          f"ls_evals {st['ls_evals']}  table-flag {int((fl & T.FLAG_TRIP_TABLE).ne(0).sum())}", flush=True)
