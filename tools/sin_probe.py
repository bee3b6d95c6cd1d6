#!/usr/bin/env python
"""Where do long sin rollouts of small batches lose time?  argv: n T kind; both kernel families, contact off/on, solver statistics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
n, Tn, kind = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
t = builtin_tables()
for fam in (T.FAMILY_ONEWARP, T.FAMILY_TEAM):
    for hulls in (None, "auto"):
        env = SOARM101VecEnv(tables=t, num_envs=n, dtype="float64", hulls=hulls)
        env.set_option(T.OPT_KERNEL_FAMILY, fam)
        env.rollout_discard(2, kind)
        torch.cuda.synchronize()
        env.stats()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); env.rollout_discard(Tn, kind, seed=42); b.record(); torch.cuda.synchronize()
        st = env.stats()
        q, v, w = env.get_state()
        print(f"family {fam} hulls {hulls}: {a.elapsed_time(b):.3f} ms  newton/step {st['newton_iters'] / st['physics_steps']:.4f} "
              f"ls/step {st['ls_evals'] / st['physics_steps']:.5f} limit steps {st['limit_steps']}  max|qvel| {float(v.abs().max()):.3f}", flush=True)
