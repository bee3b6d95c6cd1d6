#!/usr/bin/env python
"""Where the end-to-end (host buffers) time of the headline workload goes: device-only launch vs rollout_host variants."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
N, TC = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 100)
env = SOARM101VecEnv(tables=builtin_tables(), num_envs=N)
g = torch.Generator().manual_seed(0)
U = (torch.rand((TC + 1, 5, N), generator=g, dtype=torch.float64) - 0.5).pin_memory()
q0 = torch.zeros((6, N), dtype=torch.float64); q0[:5] = (torch.rand((5, N), generator=g, dtype=torch.float64) - 0.5) * 0.6
q0 = q0.pin_memory()
rows = torch.empty((N, TC + 1, 13), dtype=torch.float64).pin_memory()
Ud = U.cuda(); rows_d = torch.empty((N, TC + 1, 13), dtype=torch.float64, device="cuda")
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t.append((time.perf_counter() - t0) * 1e3)
    return min(t), sorted(t)[len(t) // 2]
print("device rollout(tensor U resident, rows to device):", timed(lambda: env.rollout(TC, "tensor", u=Ud, out=rows_d, flags=T.ROLL_NO_RESET)))
print("H2D U only:", timed(lambda: Ud.copy_(U, non_blocking=True)))
print("D2H rows only (contiguous):", timed(lambda: rows.copy_(rows_d, non_blocking=True)))
for direct in (2, 1):
    env.set_option(T.OPT_HOST_DIRECT, direct)
    for ch in ("0", "1", "2", "3"):
        env.set_option(T.OPT_HOST_CHUNKS, int(ch))
        print(f"rollout_host direct={direct} (1: rows stored into the pinned host buffer by the kernel, 2: staged) chunks={ch}:",
              timed(lambda: env.rollout_host(TC, "tensor", u_host=U, qpos0_host=q0, out_host=rows)))
env.set_option(T.OPT_HOST_DIRECT, 2)
for ch in (("1", "2", "3", "4", "6", "8", "12") if len(sys.argv) <= 2 else ()):
    env.set_option(T.OPT_HOST_CHUNKS, int(ch))
    for even in (False, True):
        env.set_option(T.OPT_HOST_EVEN, int(even))
        print(f"rollout_host chunks={ch} even={even}:", timed(lambda: env.rollout_host(TC, "tensor", u_host=U, qpos0_host=q0, out_host=rows)))
