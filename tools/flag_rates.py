#!/usr/bin/env python
"""How many envs end a rollout with each flag: argv = list of "n:T:kind:dtype" (tools/contact_perf.py's case format)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv

t = builtin_tables()
for case in sys.argv[1:] or ["65536:20:random:float64", "65536:200:random:float64", "65536:200:sin:float64", "65536:200:chirp:float64"]:
    n, Tn, kind, dtype = case.split(":")
    env = SOARM101VecEnv(tables=t, num_envs=int(n), dtype=dtype)
    env.rollout_discard(int(Tn), kind, seed=42)
    fl = env.flags()
    q, v, w = env.get_state()
    out = {name: float((fl & getattr(T, "FLAG_" + name)).ne(0).double().mean()) for name in
           ("BADSTATE", "TRIP_TABLE", "TRIP_SELF", "LIMIT", "MAXITER", "CONTACT")}
    print(case, {k: round(x, 5) for k, x in out.items()}, "max|q|", [round(float(x), 3) for x in q.abs().amax(0)], flush=True)
