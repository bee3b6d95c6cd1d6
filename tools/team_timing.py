#!/usr/bin/env python
"""Debug build only (python lerobot-mujoco-sim2real_b200/build.py -DSO101_TIMING): cycles of the dynamics warp of one team
by phase and kind of step.  argv: group [seed] [hulls 0/1]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T_, _lib
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
g = int(sys.argv[1]); seed = int(sys.argv[2]) if len(sys.argv) > 2 else 42
hulls = "auto" if (len(sys.argv) <= 3 or sys.argv[3] == "1") else None
nenv = int(sys.argv[4]) if len(sys.argv) > 4 else 32        # 4096 with group 0: the whole bench batch (128 teams at once)
env = SOARM101VecEnv(tables=builtin_tables(), num_envs=nenv, dtype="float64", hulls=hulls)
env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_TEAM)
spec = env.make_spec("random", seed, g * 32)
L = _lib.lib()
out = (C.c_ulonglong * 28)()
for rep in range(2):
    L.so101_debug_timing_double(out, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(L.so101_batch_rollout(env._h, C.byref(spec), 100, env.frame_skip, None, 0, env._stream()))
    e1.record(); torch.cuda.synchronize()
L.so101_debug_timing_double(out, 0)
t = np.array(list(out)[:15], dtype=np.float64).reshape(3, 5)
print(f"group {g}: {e0.elapsed_time(e1):.3f} ms")
for k, name in enumerate(("plain", "box tripped", "contact solved")):
    n = max(t[k, 0], 1)
    print(f"  {name:15s} steps {int(t[k,0]):5d}  cycles/step: before (A) {t[k,1]/n:7.0f}  wait (A) {t[k,2]/n:7.0f}  solve {t[k,3]/n:7.0f}  (E)..end {t[k,4]/n:7.0f}  total {t[k,1:].sum()/n:7.0f}")
h = np.array(list(out)[16:24], dtype=np.float64)
print(f"  geometry warp: {h[1] / max(h[0], 1):.0f} cycles from step start to (A), {h[2] / max(h[0], 1):.0f} cycles of work after (A); "
      f"waits {h[3] / max(h[0], 1):.0f} at (A); lookout warp: {h[5] / max(h[4], 1):.0f} cycles to (A), waits {h[6] / max(h[4], 1):.0f}")
n7 = max(float(out[23]), 1.0)
print(f"  contact steps ({int(out[23])}): M load + lagged qacc_smooth {out[15] / n7:.0f} cycles, direct solve of the other lanes {out[24] / n7:.0f}, contact block {out[25] / n7:.0f}")
