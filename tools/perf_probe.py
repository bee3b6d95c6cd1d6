#!/usr/bin/env python
"""Quick throughput probe (not the bench): physics-steps/s of the fused rollout vs batch size."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lerobot_mujoco_sim2real_b200 import builtin_tables
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv, fma_peak_tflops

tables = builtin_tables()
print("fma peak fp64 %.2f TF  fp32 %.2f TF" % (fma_peak_tflops("float64"), fma_peak_tflops("float32")))
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "4096,16384,65536,131072,262144,1048576".split(","))]
for dtype in ("float64", "float32"):
    for n in sizes:
        env = SOARM101VecEnv(tables=tables, num_envs=n, dtype=dtype)
        T = 20 if n >= 262144 else 100
        env.rollout_discard(2, "random")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        env.stats()
        e0.record()
        env.rollout_discard(T, "random")
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        st = env.stats()
        steps = n * T * 10
        print(f"{dtype} n={n:8d} T={T:4d}: {ms:9.3f} ms  {steps/ms*1e3/1e6:9.1f} M physics-steps/s  "
              f"newton/step {st['newton_iters']/st['physics_steps']:.3f} ls/step {st['ls_evals']/st['physics_steps']:.3f}", flush=True)
        del env
