#!/usr/bin/env python
"""Which team decides the launch time of a small batch with table contact: every 32-env group of the benchmark rollout
(4096 envs, 100 control steps, random controls) timed alone, with the contact path on and off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T_
from lerobot_mujoco_sim2real_b200 import _lib
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv

tables = builtin_tables()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 42
res = {}
for name, hulls in (("off", None), ("on", "auto")):
    env = SOARM101VecEnv(tables=tables, num_envs=32, dtype="float64", hulls=hulls)
    env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_TEAM)
    ts, ncon, newt = [], [], []
    for off in range(0, N, 32):
        spec = env.make_spec("random", seed, off)
        def run():
            _lib.check(_lib.lib().so101_batch_rollout(env._h, C.byref(spec), 100, env.frame_skip, None, 0, env._stream()))
        run(); torch.cuda.synchronize(); env.stats(); env.clear_flags()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        st = env.stats()
        fl = env.flags().cpu().numpy()
        ts.append(e0.elapsed_time(e1)); newt.append(st["newton_iters"] / st["physics_steps"])
        ncon.append(int(((fl & (T_.FLAG_CONTACT | T_.FLAG_TRIP_TABLE)) != 0).sum()))
    res[name] = (np.array(ts), np.array(ncon), np.array(newt))
on, off = res["on"], res["off"]
print(f"alone, contact off: median {np.median(off[0]):.3f} max {off[0].max():.3f} ms; on: median {np.median(on[0]):.3f} max {on[0].max():.3f} ms")
order = np.argsort(on[0])[::-1][:10]
for o in order:
    print(f"  group {int(o):4d}: on {on[0][o]:.3f} ms  off {off[0][o]:.3f} ms  envs in contact {on[1][o]} (flagged when off: {off[1][o]})  newton/step {on[2][o]:.4f}")
print("groups without contact: on", np.median(on[0][on[1] == 0]), "off", np.median(off[0][on[1] == 0]))
