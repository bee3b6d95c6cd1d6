#!/usr/bin/env python
"""Per-frame cost of the Koopman_MPC loop body on one GPU: so101_koopman_mpc_step (encoder + gains + clip) and the stepper's
env step with gravity compensation, for n envs.  argv: n [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lerobot_mujoco_sim2real_b200 import builtin_tables
from lerobot_mujoco_sim2real_b200.Koopman_MPC import BatchedKoopmanMPC
from lerobot_mujoco_sim2real_b200.TrajectoryGenerator import CartesianTrajectoryGenerator
from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 50
tables = builtin_tables()
km = KoopmanModel.from_npz(os.path.join(ROOT, "tests", "golden", "koopman_dkuc.npz"))
gen = CartesianTrajectoryGenerator(tables=tables)
names = np.arange(n) % 2 == 0
centers = np.stack([np.full(n, 0.4), 0.01 * np.sin(np.arange(n)), 0.2 + 0.01 * np.cos(np.arange(n))], axis=1)
xyz, q, st = gen.generate_batch(names, idx=np.ones(n, dtype=np.int64), traj_scale=np.full(n, 0.5), centers=centers)
env = SOARM101VecEnv(tables=tables, num_envs=n, gravity_compensation=True)
ev = lambda: torch.cuda.Event(enable_timing=True)
for mpc_type in ("delta_mpc", "mpc"):
    a, b = ev(), ev()
    loop = BatchedKoopmanMPC(env, km, xyz, q, H=10, MPC_type=mpc_type)
    a.record(); loop._fold_reference(loop.total_frames); b.record(); torch.cuda.synchronize()
    ff_ms = a.elapsed_time(b)
    loop.runBefore()
    for _ in range(3): loop.runMPC()
    torch.cuda.synchronize()
    a, b = ev(), ev()
    a.record()
    for _ in range(frames): loop.runMPC()
    b.record(); torch.cuda.synchronize()
    tot = a.elapsed_time(b) / frames
    # the two launches alone
    e = [ev() for _ in range(3)]
    tm = ts = 0.0
    aout = torch.empty((n, 5), dtype=torch.float64, device="cuda")
    for _ in range(frames):
        e[0].record()
        km.mpc_step(env._obs, True, loop.uff, 20, loop.u_prev, loop._ctrl, aout, 10, mpc_type, 0.5)
        e[1].record()
        env.step_soa(loop._ctrl)
        e[2].record(); torch.cuda.synchronize()
        tm += e[0].elapsed_time(e[1]); ts += e[1].elapsed_time(e[2])
    lifts = 14.3e3 * 2 * n
    print(f"{mpc_type}: n={n}: reference fold (once, {xyz.shape[1]} rows per curve) {ff_ms:.1f} ms; loop {tot:.3f} ms per frame "
          f"(mpc_step {tm / frames:.3f} ms = {lifts / (tm / frames * 1e-3) / 1e12:.1f} TFLOP/s fp64, env step {ts / frames:.3f} ms)", flush=True)
