"""Bisect helper: one rollout in a fresh process.  argv: family n T kind hulls(0/1) dtype"""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, numpy as np
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
t = builtin_tables()
fam, n, Tn, kind, hl = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
dtype = sys.argv[6] if len(sys.argv) > 6 else "float64"
env = SOARM101VecEnv(tables=t, num_envs=n, hulls="auto" if hl else None, dtype=dtype)
env.set_option(T.OPT_KERNEL_FAMILY, fam)
env.rollout(Tn, kind, seed=42)
torch.cuda.synchronize()
fl = env.flags().cpu().numpy()
print("ok", sys.argv[1:], "contact", ((fl & T.FLAG_CONTACT) != 0).mean(), "table", ((fl & T.FLAG_TRIP_TABLE) != 0).mean(), flush=True)
