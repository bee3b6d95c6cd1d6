#!/usr/bin/env python
"""TrajectoryGenerator-driven Koopman dataset generation on G GPUs (BASELINE.json config 4's wording, rows N3 + N4 on
the stepper): every rank builds its shard of Cartesian reference curves on the device, solves the inverse kinematics
of all of them in one launch, runs the reference's Koopman_MPC.py loop along them (lift -> closed-form MPC -> env.step
with gravity compensation) and returns the closed-loop run in the dataset row layout; NCCL gathers the rows on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29555 \
        tools/mpc_dataset_multi.py [--curves 262144] [--frames 100] [--out gpurun_out/mpc_dataset_Ggpu.json]

Curve parameters are a pure function of the GLOBAL curve index, so the dataset does not depend on G (rank 0 checks a
slice of another rank's shard by re-running it alone)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from lerobot_mujoco_sim2real_b200 import builtin_tables, sharding
from lerobot_mujoco_sim2real_b200.Koopman_MPC import BatchedKoopmanMPC
from lerobot_mujoco_sim2real_b200.TrajectoryGenerator import CartesianTrajectoryGenerator
from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--curves", type=int, default=1 << 18)
ap.add_argument("--frames", type=int, default=100)
ap.add_argument("--out", default=None)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
tables = builtin_tables("scene_with_table_v.xml")
km = KoopmanModel.from_npz(os.path.join(ROOT, "tests", "golden", "koopman_dkuc.npz"), device=dev)   # the reference's shipped model


def curve_params(lo, hi):
    """Per-curve parameters keyed by the global curve index (Philox-free: a hash of the index is enough here)."""
    i = np.arange(lo, hi, dtype=np.uint64)
    h = (i * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(11)
    u = [((h >> np.uint64(8 * k)) & np.uint64(0xFF)).astype(np.float64) / 255.0 for k in range(4)]
    names = u[0] < 0.5                                                  # True = Fig8, False = Circle
    centers = np.stack([0.4 + 0 * u[1], 0.03 * (u[1] - 0.5), 0.2 + 0.03 * (u[2] - 0.5)], axis=1)    # y-z plane (idx = 1)
    return names, centers, 0.4 + 0.2 * u[3]


def run_shard(lo, hi, frames):
    n = hi - lo
    gen = CartesianTrajectoryGenerator(tables=tables, device=local)
    names, centers, scale = curve_params(lo, hi)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    xyz, q, st = gen.generate_batch(names, idx=np.ones(n, dtype=np.int64), traj_scale=scale, centers=centers)
    ev[1].record()
    env = SOARM101VecEnv(tables=tables, num_envs=n, dtype="float64", device=local, gravity_compensation=True)
    loop = BatchedKoopmanMPC(env, km, xyz, q, H=10)
    actual = loop.run(frames)
    rows = loop.dataset_rows()
    ev[2].record()
    torch.cuda.synchronize()
    ee_err = (actual[:, 10:, :3] - xyz[:, 11:frames + 1]).norm(dim=2).mean().item() if frames > 11 else float("nan")
    return rows, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), float((st & 1).double().mean().item()), ee_err


def tmax(x):
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


lo, hi = sharding.shard_range(args.curves, rank, world)
run_shard(lo, min(hi, lo + 256), 3)                                  # warm-up (kernels, allocator, NCCL below)
if world > 1:
    sharding.gather_rows(torch.zeros((hi - lo, 1, 13), dtype=torch.float64, device=dev), args.curves, dst=0)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
rows, ik_ms, loop_ms, ok, ee_err = run_shard(lo, hi, args.frames)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
full = sharding.gather_rows(rows, args.curves, dst=0)
b.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
res = {"n_gpus": world, "curves_total": args.curves, "curves_per_gpu": hi - lo, "frames": args.frames,
       "ik_ms": tmax(ik_ms), "mpc_loop_ms": tmax(loop_ms), "gather_ms": tmax(a.elapsed_time(b)), "wall_ms": tmax(wall * 1e3),
       "env_steps_per_s_wall": args.curves * args.frames / tmax(wall), "dataset_bytes": args.curves * args.frames * 13 * 8,
       "ik_success_frac_rank0": ok, "ee_tracking_error_mm_rank0": ee_err * 1e3}
if rank == 0:
    assert full.shape == (args.curves, args.frames, 13)
    r = world - 1
    off = sharding.shard_range(args.curves, r, world)[0] + 777
    ref = run_shard(off, off + 512, args.frames)[0]
    res["gathered_rows_bitwise_equal_to_a_slice_rerun_alone"] = bool(torch.equal(ref, full[off:off + 512]))
    line = json.dumps(res)
    print(line)
    if args.out:
        open(args.out, "w").write(line + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
