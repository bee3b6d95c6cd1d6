#!/usr/bin/env python
"""BASELINE.json config 4 at full size on G GPUs of one box: 2^20 envs sharded by rank, one fused rollout per rank
(no inter-GPU traffic on the step path), NCCL gather of the float64 dataset rows to rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29533 \
        tools/config4_multi.py [--envs 1048576] [--out gpurun_out/config4_Ggpu.json]

Times are CUDA-event times, max over ranks.  Rank 0 then re-simulates a few 4096-env slices of other ranks' shards
alone (env_offset = global env index) and requires the gathered rows to be bit-identical: the dataset does not depend
on G."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from lerobot_mujoco_sim2real_b200 import builtin_tables, sharding
from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=1 << 20)
ap.add_argument("--out", default=None)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
tables = builtin_tables("scene_with_table_v.xml")
lo, hi = sharding.shard_range(args.envs, rank, world)
n = hi - lo
env = SOARM101VecEnv(tables=tables, num_envs=n, dtype="float64", device=local, seed=42)


def tmax(ms: float) -> float:
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


if world > 1:   # NCCL sets its connections up lazily: one small gather before anything is timed
    sharding.gather_rows(torch.zeros((n, 1, 13), dtype=torch.float64, device=dev), args.envs, dst=0)
    torch.cuda.synchronize()
res = {"n_gpus": world, "envs_total": args.envs, "envs_per_gpu": n, "dtype": "f64", "cases": {}}
for kind, T in (("random", 20), ("chirp", 200)):
    rows = torch.empty((n, T + 1, 13), dtype=torch.float64, device=dev)
    env.rollout(2, kind, seed=42, env_offset=lo, out=rows[:, :3].contiguous())   # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record()
    env.rollout(T, kind, seed=42, env_offset=lo, out=rows)
    b.record()
    full = sharding.gather_rows(rows, args.envs, dst=0)
    c.record()
    torch.cuda.synchronize()
    sim_ms, gather_ms, tot_ms = tmax(a.elapsed_time(b)), tmax(b.elapsed_time(c)), tmax(a.elapsed_time(c))
    flagged = torch.tensor([int((env.flags() != 0).sum().item())], device=dev)
    if world > 1:
        dist.all_reduce(flagged)
    case = {"control": kind, "T": T, "sim_ms": sim_ms, "gather_ms": gather_ms, "total_ms": tot_ms,
            "env_steps_per_s_sim": args.envs * T / (sim_ms * 1e-3), "physics_steps_per_s_sim": args.envs * T * 10 / (sim_ms * 1e-3),
            "env_steps_per_s_incl_gather": args.envs * T / (tot_ms * 1e-3),
            "dataset_bytes": args.envs * (T + 1) * 13 * 8,
            "gather_GBps_into_rank0": (args.envs - n) * (T + 1) * 13 * 8 / (gather_ms * 1e-3) / 1e9 if world > 1 else None,
            "envs_flagged_by_tripwire": int(flagged.item())}
    if rank == 0:
        assert full.shape == (args.envs, T + 1, 13)
        chk = SOARM101VecEnv(tables=tables, num_envs=4096, dtype="float64", device=local, seed=42)
        same = True
        for r in sorted({0, world // 2, world - 1}):
            off = sharding.shard_range(args.envs, r, world)[0] + 12345 % max(1, n - 4096)
            ref = chk.rollout(T, kind, seed=42, env_offset=off)
            same &= bool(torch.equal(ref, full[off:off + 4096]))
        case["gathered_rows_bitwise_equal_to_single_gpu_slices"] = same
        assert same
        del chk
    res["cases"][f"{kind}_T{T}"] = case
    del rows, full
    torch.cuda.empty_cache()
if rank == 0:
    line = json.dumps(res)
    print(line)
    if args.out:
        open(args.out, "w").write(line + "\n")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
