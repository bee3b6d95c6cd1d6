#!/usr/bin/env python
"""What is left between the end-to-end call and its own kernel: the bench's e2e workload (fixed host controls and initial angles),
L2 flushed before every timed call as bench.py does.  (a) the device-resident launch from the same start, (b) rollout_host."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
N, TC, SEED = 4096, 100, 42
env = SOARM101VecEnv(tables=builtin_tables(), num_envs=N, seed=SEED)
g = torch.Generator().manual_seed(SEED)
U = (torch.rand((TC + 1, 5, N), generator=g, dtype=torch.float64) - 0.5).pin_memory()
q0 = torch.zeros((6, N), dtype=torch.float64); q0[:5] = (torch.rand((5, N), generator=g, dtype=torch.float64) - 0.5) * 0.6
q0 = q0.pin_memory()
rows = torch.empty((N, TC + 1, 13), dtype=torch.float64).pin_memory()
Ud, q0d = U.cuda(), q0.cuda()
rows_d = torch.empty((N, TC + 1, 13), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
z = torch.zeros((N, 6), dtype=torch.float64, device="cuda")
def dev():
    env.set_state(q0d.t().contiguous(), z, z)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.fill_(1); a.record()
    env.rollout(TC, "tensor", u=Ud, out=rows_d, flags=T.ROLL_NO_RESET)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)
def host():
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.fill_(1); torch.cuda.synchronize(); t0 = time.perf_counter(); a.record()
    env.rollout_host(TC, "tensor", u_host=U, qpos0_host=q0, out_host=rows)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b), (time.perf_counter() - t0) * 1e3
for _ in range(3): dev(); host()
d = sorted(dev() for _ in range(10)); h = sorted(host() for _ in range(10))
print(f"kernel alone (same start, same controls): median {d[5]:.3f} ms   rollout_host: median {h[5][0]:.3f} ms by events, {h[5][1]:.3f} ms host clock")
for direct in (2, 1):
    env.set_option(T.OPT_HOST_DIRECT, direct)
    for ch in (0, 1):
        env.set_option(T.OPT_HOST_CHUNKS, ch)
        h = sorted(host() for _ in range(10))
        print(f"  direct={direct} chunks={ch}: {h[5][0]:.3f} ms")
