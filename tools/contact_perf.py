#!/usr/bin/env python
"""What table-plane contact costs: the same rollouts with the contact path on (hull data loaded) and off (envs that reach
the table are only flagged).  argv: list of "n:T:kind:dtype" cases."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv

t = builtin_tables()
cases = sys.argv[1:] or ["4096:100:random:float64", "131072:20:random:float64", "131072:200:chirp:float64", "65536:200:random:float32"]
for case in cases:
    n, Tn, kind, dtype, *opt = case.split(":")          # optional 5th field: SO101_OPT_REGROUP value for the "on" run
    n, Tn = int(n), int(Tn)
    res = {}
    for name, hulls in (("off", None), ("on", "auto")):
        env = SOARM101VecEnv(tables=t, num_envs=n, dtype=dtype, hulls=hulls)
        if opt and name == "on":
            env.set_option(T.OPT_REGROUP, int(opt[0]))
        if len(opt) > 1:                                    # optional 6th field: block size of the one-warp kernels (both runs)
            env.set_option(T.OPT_KERNEL_FAMILY, T.FAMILY_ONEWARP)
            env.set_option(T.OPT_BLOCK, int(opt[1]))
        env.rollout_discard(2, kind)
        torch.cuda.synchronize()
        best = 1e30
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); env.rollout_discard(Tn, kind, seed=42 + rep); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        fl = env.flags().cpu().numpy()
        res[name] = (best, float(((fl & T.FLAG_CONTACT) != 0).mean()), float(((fl & T.FLAG_TRIP_TABLE) != 0).mean()))
        del env
    print(f"{case}: off {res['off'][0]:.3f} ms (flagged {res['off'][2]:.4f})  on {res['on'][0]:.3f} ms (in contact {res['on'][1]:.4f}, "
          f"unsimulated {res['on'][2]:.5f})  ratio {res['on'][0] / res['off'][0]:.3f}", flush=True)
