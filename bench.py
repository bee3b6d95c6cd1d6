#!/usr/bin/env python
"""bench.py — batched SO-ARM101 stepping: env-steps/s on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE.json configs[1] — 4096 envs per GPU, fp64, 1000 physics
steps per rollout (100 control steps x frame_skip 10), U(-0.3,0.3) reset, U(-0.5,0.5) random
controls (Philox, seed 42), scene_with_table_v.xml.  One bench "step" = one such rollout of the
whole batch = one launch of the fused kernel (`so101_batch_rollout`).

  value      env-steps/s (1 env-step = SOARM101Env.step = 10 mj_step), state and rows device resident
  e2e        same metric through the C ABI with HOST buffers: pinned control tensor + initial
             states H2D, rollout, dataset rows D2H, all inside the timed region
  roofline   dominant kernel vs the FP64 pipe (this path is FP64-bound): achieved =
             6.2 kFLOP x physics steps / CUDA-event time, peak = FMA loop measured in this run
  cpu_baseline  the CPU oracle (oracle/, a port of the mj_step restatement; real MuJoCo is not
             installable offline) on the box's host cores, bounded sample of the same workload

--impl reference times that CPU oracle with all host threads (rank 0 only under torchrun).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS = 4096          # per GPU (weak scaling)
T_CTRL = 100           # control steps per rollout
FRAME_SKIP = 10
SEED = 42
FLOP_PER_PHYSICS_STEP = 6.2e3   # SURVEY.md 8d / BASELINE.md (4.9 k + 0.63 k x 2 Newton iterations)
METRIC, UNIT = "env-steps/sec", "env-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary configs (fp32 65536, fp64 large batch)")
    return ap.parse_args()


def workload_config(n_gpus: int) -> dict:
    return {"workload": "BASELINE.json configs[1]: SO-ARM101 scene_with_table_v, 4096 envs/GPU fp64, "
                        "1000 physics steps (100 ctrl x frame_skip 10) random-control rollouts",
            "n_envs_per_gpu": N_ENVS, "n_envs_total": N_ENVS * n_gpus, "control_steps": T_CTRL,
            "frame_skip": FRAME_SKIP, "seed": SEED, "parallelism": f"env-sharded x{n_gpus}, no step-path collective",
            "l2": "flushed between timed iterations (256 MiB write)"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle on host cores
# ------------------------------------------------------------------------------------------------
def oracle_rate(n_envs: int, reps: int, warm: int, threads: int = 0):
    """-> (env-steps/s, seconds per rep, threads used) for `n_envs` x T_CTRL rollouts on the CPU oracle."""
    from lerobot_mujoco_sim2real_b200 import builtin_tables
    from oracle import oracle as O
    O.build()
    tables = builtin_tables("scene_with_table_v.xml")
    # torchrun exports OMP_NUM_THREADS=1: ask the OS for the cores this process may use instead
    nthr = threads or len(os.sched_getaffinity(0))
    times = []
    for r in range(warm + reps):
        spec = O.make_spec(kind=0, seed=SEED + r)
        t0 = time.perf_counter()
        O.rollout(tables, spec, n_envs, T_CTRL, FRAME_SKIP, nthreads=nthr, want_rows=True)
        dt = time.perf_counter() - t0
        if r >= warm:
            times.append(dt)
    mean = sum(times) / len(times)
    return n_envs * T_CTRL / mean, mean, nthr


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = 512   # x 1000 physics steps = 0.5 M mj_step-equivalents per step: a few seconds on 16 cores
    rate, sec, nthr = oracle_rate(n_sample, args.steps, args.warmup)
    sample = (f"{n_sample} of the {N_ENVS} envs x {T_CTRL} control steps x {FRAME_SKIP} sub-steps per step "
              f"(same reset/control distributions), OpenMP over envs")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "physics_steps_per_s": rate * FRAME_SKIP,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": nthr, "kind": "port", "sample": sample,
                         "note": "restated-oracle CPU baseline (oracle/so101_oracle.c), not MuJoCo: the mujoco "
                                 "wheel is not installable offline"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist
    from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv, fma_peak_tflops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        # stdout carries exactly one JSON line: NCCL's own version / debug lines go to stderr
        # (NCCL_DEBUG=VERSION makes NCCL printf its version to stdout; INFO and above honour NCCL_DEBUG_FILE)
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
            sys.stderr.write("bench.py: NCCL_DEBUG=VERSION -> WARN (keeps stdout to one JSON line)\n")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", torch.cuda.current_device())
    n_gpus = world
    K, W = args.steps, max(args.warmup, 0)
    tables = builtin_tables("scene_with_table_v.xml")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn, k: int, w: int):
        """k timed calls of fn(step_index); CUDA events around each call on the launching stream,
        L2 flushed in between.  -> (sum of per-call ms, wall ms of the bracketed region)"""
        for i in range(w):
            fn(i)
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        t0 = time.perf_counter()
        for i in range(k):
            flush_buf.fill_(i & 0xFF)
            evs[i][0].record()
            fn(w + i)
            evs[i][1].record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        return sum(a.elapsed_time(b) for a, b in evs), wall

    # ---- headline: device-resident fused rollout -----------------------------------------------
    env = SOARM101VecEnv(tables=tables, num_envs=N_ENVS, dtype="float64", device=dev.index, seed=SEED)
    rows = torch.empty((N_ENVS, T_CTRL + 1, T.ROW), dtype=torch.float64, device=dev)
    env_offset = rank * N_ENVS

    def step_dev(i):
        env.rollout(T_CTRL, "random", seed=SEED + i, env_offset=env_offset, out=rows)

    step_dev(0)
    torch.cuda.synchronize()
    env.stats()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    ms_sum, wall_ms = timed(step_dev, K, W)
    clocks = sampler.stop() if rank == 0 else None
    st = env.stats()
    flags = env.flags()
    n_trip = int((flags & T.FLAG_TRIP).ne(0).sum().item())
    n_bad = int((flags & T.FLAG_BADSTATE).ne(0).sum().item())
    ms_sum = max_over_ranks(ms_sum)
    ms_per_step = ms_sum / K
    env_steps_per_launch = N_ENVS * T_CTRL
    value = n_gpus * env_steps_per_launch / (ms_per_step * 1e-3)
    kernel_ms = ms_per_step            # the step IS one launch of k_rollout<double,double>
    phys_per_launch = env_steps_per_launch * FRAME_SKIP

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------
    g = torch.Generator().manual_seed(SEED + rank)
    U_host = ((torch.rand((T_CTRL + 1, T.NU_ENV, N_ENVS), generator=g, dtype=torch.float64) - 0.5)).pin_memory()
    q0_host = torch.zeros((T.NV, N_ENVS), dtype=torch.float64)
    q0_host[:5] = (torch.rand((5, N_ENVS), generator=g, dtype=torch.float64) - 0.5) * 0.6
    q0_host = q0_host.pin_memory()
    rows_host = torch.empty((N_ENVS, T_CTRL + 1, T.ROW), dtype=torch.float64).pin_memory()

    def step_e2e(i):
        # ONE C-ABI call with host buffers: H2D controls + initial angles, k_reset, k_rollout, D2H rows, sync
        env.rollout_host(T_CTRL, "tensor", u_host=U_host, qpos0_host=q0_host, out_host=rows_host)

    e2e_ms, _ = timed(step_e2e, K, W)
    e2e_ms = max_over_ranks(e2e_ms) / K
    e2e_value = n_gpus * env_steps_per_launch / (e2e_ms * 1e-3)
    h2d = U_host.numel() * 8 + q0_host.numel() * 8
    d2h = rows_host.numel() * 8

    # ---- dataset gather over NCCL (config 4's only collective), timed separately ------------------
    gather_ms = None
    if world > 1:
        from lerobot_mujoco_sim2real_b200 import sharding
        sharding.gather_rows(rows, N_ENVS * world, dst=0)     # untimed: first use sets up the NCCL channels
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sharding.gather_rows(rows, N_ENVS * world, dst=0)
        e1.record()
        barrier()
        gather_ms = max_over_ranks(e0.elapsed_time(e1))

    if rank != 0:
        dist.barrier()          # rank 0 is still measuring its CPU baseline / secondary configs
        dist.destroy_process_group()
        return

    # ---- rank 0 only: roofline denominators, secondary configs, CPU baseline -----------------------
    peak64 = fma_peak_tflops("float64", dev.index)
    peak32 = fma_peak_tflops("float32", dev.index)
    achieved_tf = FLOP_PER_PHYSICS_STEP * phys_per_launch / (kernel_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch_config1")
        except Exception:
            traffic = None
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = env_steps_per_launch * 208.0   # fused rollout, fp64: 208 B per env-step
    roofline = {
        "bound": "fp64_pipe", "kernel": "k_rollout<double,double,SPLIT=true> (three-warp team kernel, batches <= 9472 envs)", "achieved": achieved_tf, "peak": peak64,
        "unit": "TFLOP/s", "frac": achieved_tf / peak64, "traffic": traffic,
        "peak_source": "register-resident DFMA loop measured in this run (so101_fma_peak), 'of measured'",
        "algorithmic_flop_per_physics_step": FLOP_PER_PHYSICS_STEP, "physics_steps_per_launch": phys_per_launch,
        "kernel_ms": kernel_ms,
        "bound_note": "neither hbm nor tensor: ~300 FLOP per byte of state/row traffic, no contraction; the limiter is "
                      "the FP64 FMA pipe.  At this config (4096 envs = 128 three-warp teams, one per SM, on 148 SMs) the launch is "
                      "latency-bound: a physics step is one dependent chain of ~2.8 k instructions on the team's dynamics warp (ncu: 41 % of its samples are fixed-latency waits); the large-batch fraction is in extra (fp64 131072: ~0.50 - the accounting budgets 2 Newton iterations and a line search, the kernels need one direct solve; ncu: FP64 pipe 45 % busy)",
        "hbm_sanity": {"achieved_gbs": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                       "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
                       "algorithmic_bytes_per_env_step": 208},
        "newton_iters_per_physics_step": st["newton_iters"] / max(1, st["physics_steps"]),
        "ls_evals_per_physics_step": st["ls_evals"] / max(1, st["physics_steps"]),
    }

    extra = {}
    if not args.no_extra:
        def quick(n, dtype, t_ctrl, reps=3):
            e = SOARM101VecEnv(tables=tables, num_envs=n, dtype=dtype, device=dev.index, seed=SEED)
            e.rollout_discard(2, "random")
            torch.cuda.synchronize()
            best = 1e30
            for r in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                flush_buf.fill_(r)
                a.record()
                e.rollout_discard(t_ctrl, "random", seed=SEED + r)
                b.record()
                torch.cuda.synchronize()
                best = min(best, a.elapsed_time(b))
            phys = n * t_ctrl * FRAME_SKIP / (best * 1e-3)
            pk = peak64 if dtype == "float64" else peak32
            return {"n_envs": n, "dtype": dtype, "control_steps": t_ctrl, "ms": best, "env_steps_per_s": phys / FRAME_SKIP,
                    "physics_steps_per_s": phys, "roofline_frac": FLOP_PER_PHYSICS_STEP * phys / 1e12 / pk}
        extra["config3_fp32_65536"] = quick(65536, "float32", 200)
        extra["fp64_131072_per_gpu(config4 shard at 8 GPUs)"] = quick(131072, "float64", 20)
        extra["fp32_4096"] = quick(4096, "float32", 100)
        extra["fma_peak_tflops"] = {"fp64": peak64, "fp32": peak32}
        # config 4 shape on one GPU: train-shaped rollouts WITH the dataset rows written (T = 20)
        e4 = SOARM101VecEnv(tables=tables, num_envs=131072, dtype="float64", device=dev.index, seed=SEED)
        rows4 = torch.empty((131072, 21, T.ROW), dtype=torch.float64, device=dev)
        e4.rollout(20, "random", out=rows4)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); e4.rollout(20, "random", seed=SEED + 1, out=rows4); b.record(); torch.cuda.synchronize()
        f4 = e4.flags()
        extra["config4_shard_131072_T20_with_rows"] = {
            "ms": a.elapsed_time(b), "env_steps_per_s": 131072 * 20 / (a.elapsed_time(b) * 1e-3),
            "rows_bytes": rows4.numel() * 8, "envs_tripwire": int((f4 & T.FLAG_TRIP).ne(0).sum().item())}
        del e4, rows4
        # config 5: 8192 control sequences x 50 env-steps from one shared state
        e5 = SOARM101VecEnv(tables=tables, num_envs=8192, dtype="float64", device=dev.index, seed=SEED)
        U5 = (torch.rand((50, T.NU_ENV, 8192), dtype=torch.float64, device=dev) - 0.5).contiguous()
        s0 = [0.1, -0.2, 0.15, 0.05, -0.1, 0.0] + [0.0] * 12
        e5.shoot(s0, U5)
        torch.cuda.synchronize()
        best5 = 1e30
        for gflag in (0, T.ROLL_GRAVCOMP_HOLD):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); e5.shoot(s0, U5, flags=gflag); b.record(); torch.cuda.synchronize()
            best5 = min(best5, a.elapsed_time(b))
        extra["config5_shoot_8192x50"] = {"ms": best5, "env_steps_per_s": 8192 * 50 / (best5 * 1e-3),
                                          "rollouts_per_s": 8192 / (best5 * 1e-3)}
        del e5
        if rank == 0:   # single-GPU side measurements: no collectives, nothing on stdout
            import contextlib
            # BASELINE configs[0] workload through the drop-in API: the reference's whole data job with args.py defaults
            # (train 50000 x 20 random; val 2000 x 200 random; tests 2000 x 200 random / sin / chirp), host arrays out
            import types
            from lerobot_mujoco_sim2real_b200.SOARM101_DataCollection import SOARM101DataGenerator
            jargs = types.SimpleNamespace(xml_path="unused", x_dim=8, u_dim=5, device="cuda", seed=SEED, env="SOARM101")
            with contextlib.redirect_stdout(sys.stderr):
                jgen = SOARM101DataGenerator(jargs, tables=tables, device=dev.index)
            job = [(50000, 20, "random"), (2000, 200, "random"), (2000, 200, "random"), (2000, 200, "sin"), (2000, 200, "chirp")]
            jgen.generate_device(64, 2, "random")[0].cpu().numpy()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nbytes = 0
            for (jn, jt, jk) in job:
                arr = jgen.generate_device(jn, jt, jk, seed=SEED)[0].cpu().numpy()   # what generate_physics_based_data returns
                nbytes += arr.nbytes
            job_s = time.perf_counter() - t0
            job_steps = sum(jn * jt for jn, jt, _ in job)
            extra["config0_reference_data_job_args_py_defaults"] = {
                "datasets": [f"{jn}x{jt} {jk}" for jn, jt, jk in job], "env_steps": job_steps, "wall_ms": job_s * 1e3,
                "env_steps_per_s": job_steps / job_s, "host_bytes_out": nbytes,
                "note": "SOARM101DataGenerator: the five datasets generate_and_save_data builds before np.save (rank 0 only, no gather), "
                        "host wall clock incl. device->host copies; the reference runs this serially on one CPU MuJoCo env"}
            del jgen
            # row N3: 16384 reference curves (300 way-points each, position-only as Koopman_MPC.py runs them) -> joint tracks
            import numpy as np
            from lerobot_mujoco_sim2real_b200.TrajectoryGenerator import CartesianTrajectoryGenerator, reference_curve
            gen = CartesianTrajectoryGenerator(tables=tables, device=dev.index)
            nik = 16384
            base = np.stack([reference_curve(nm, ix)[0] for nm in ("Fig8", "Circle") for ix in (0, 1)])
            xyz = torch.as_tensor(base[np.arange(nik) % 4] + np.random.default_rng(SEED).uniform(-0.03, 0.03, (nik, 1, 3)),
                                  device=dev)
            gen.solve_tracks(xyz)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); qik, stik = gen.solve_tracks(xyz); b.record(); torch.cuda.synchronize()
            ms_ik = a.elapsed_time(b)
            extra["n3_ik_tracks_16384x300"] = {
                "ms": ms_ik, "waypoints_per_s": nik * 300 / (ms_ik * 1e-3),
                "gauss_newton_iters_per_waypoint": float((stik >> 8).double().mean().item()),
                "success_frac": float((stik & 1).double().mean().item()),
                "note": "k_ik_track<false>: one thread per track incl. the [n,P,3] -> [P,3,n] relayout"}
            del gen, xyz, qik, stik

    # CPU baseline: bounded sample sized for ~10-20 s of CPU work
    cal_rate, _, nthr = oracle_rate(64, 1, 1)
    n_sample = int(min(N_ENVS, max(128, 12.0 * cal_rate / T_CTRL)))
    cpu_rate, cpu_sec, nthr = oracle_rate(n_sample, 1, 0)
    cpu_baseline = {"value": cpu_rate, "unit": UNIT, "cores": nthr, "kind": "port",
                    "sample": f"{n_sample} of the {N_ENVS} envs x {T_CTRL} control steps x {FRAME_SKIP} sub-steps "
                              f"({cpu_sec:.1f} s), same distributions, OpenMP over envs",
                    "physics_steps_per_s": cpu_rate * FRAME_SKIP,
                    "note": "restated-oracle CPU baseline (oracle/so101_oracle.c), not MuJoCo"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(n_gpus),
        "physics_steps_per_s": value * FRAME_SKIP,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms,
                "path": "so101_batch_rollout_host: pinned host control tensor [T+1,5,N] + initial angles [6,N] -> "
                        "k_reset + up to 12 time-chunked k_rollout launches (SO101_CTRL_TENSOR) overlapped with the upload of the next chunk's controls and the download of the previous chunk's rows -> dataset rows [N,T+1,13] in pinned host memory, stream sync"},
        "gpu_launches": K,
        "gpu_launches_note": "one k_rollout<double,double,true> launch per step in the device-resident region; "
                             "the e2e region launches k_reset + 11 k_rollout chunks per step",
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
        "wall_ms_timed_region": wall_ms,
        "flags": {"envs_tripwire": n_trip, "envs_badstate": n_bad, "of": N_ENVS},
        "gather_ms": gather_ms, "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
