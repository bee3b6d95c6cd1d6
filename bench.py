#!/usr/bin/env python
"""bench.py — batched SO-ARM101 stepping: env-steps/s on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 1|3|4|5|ik]

Workloads (`config.workload`), all on scene_with_table_v.xml with U(-0.3,0.3) resets and Philox controls (seed 42):
  --config 1 (default at N = 1)  BASELINE.json configs[1]: 4096 envs per GPU, fp64, 1000 physics steps per rollout
             (100 control steps x frame_skip 10), random controls.  One "step" = one rollout of the batch = one launch
             of the fused kernel.  `e2e` = the same through ONE C-ABI call with HOST buffers (pinned controls + initial
             states H2D, rollout, dataset rows D2H inside the timed region).  N > 1: weak scaling, no collective.
  --config 4 (default at N > 1)  BASELINE.json configs[3], the multi-GPU job of north_star: 2^20 envs fp64 SHARDED over the
             N ranks (strong scaling), train-shaped T = 20 random-control rollouts.  `value` = the simulation (rows to
             each rank's own HBM); `e2e` = simulation + delivery of the whole [2^20, 21, 13] dataset into rank 0's HBM:
             every rank's row writer stores straight into rank 0's buffer over NVLink (sharding.SharedRows, CUDA IPC),
             so there is no collective to wait for - one barrier ends the step.  The test-shaped T = 200 chirp job is
             measured the same way and reported inside `e2e` / `config` (first-class keys, not `extra`).
  --config 3 | 5 | ik (N = 1)    configs[2] (65536 envs fp32), configs[4] (8192 x 50 shooting), row N3 (IK tracks) as
             first-class lines.

  roofline   dominant kernel vs the FP64 (fp32: FP32) pipe - this path is compute-bound: achieved =
             6.2 kFLOP x physics steps / CUDA-event time; `frac_executed` uses the FLOPs the kernels really execute;
             peak = FMA loop measured in this run ("of measured"), nominal peak at the sampled SM clock beside it
  cpu_baseline  the CPU oracle (oracle/, a port of the mj_step restatement; real MuJoCo is not installable offline) on
             the box's host cores, bounded sample of the same workload

--impl reference times that CPU oracle with all host threads on the same workload (rank 0 only under torchrun).
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

N_ENVS = 4096          # config 1: per GPU (weak scaling)
T_CTRL = 100           # config 1: control steps per rollout
N_TOTAL4 = 1 << 20     # config 4: envs over all ranks (strong scaling)
T4_TRAIN, T4_TEST = 20, 200
FRAME_SKIP = 10
SEED = 42
FLOP_PER_PHYSICS_STEP = 6.2e3   # SURVEY.md 8d / BASELINE.md (4.9 k + 0.63 k x 2 Newton iterations)
# FLOPs the kernels execute per physics step (ncu op counts, profiles/r1_fp_op_counts.json): the link-local formulation
# and the direct active-set solve need fewer than the algorithmic count
FLOP_EXECUTED = {"onewarp": 3.9e3, "team": 5.0e3}
METRIC, UNIT = "env-steps/sec", "env-steps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default=None, choices=["1", "3", "4", "5", "ik"],
                    help="workload; default: 1 on one GPU, 4 (the sharded 2^20-env job) on several")
    ap.add_argument("--no-extra", action="store_true", help="config 1: skip the secondary measurements")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.config is None:
        a.config = "1" if max(world, a.gpus) == 1 else "4"
    return a


def workload_config(cfg: str, n_gpus: int) -> dict:
    if cfg == "4":
        return {"workload": "BASELINE.json configs[3]: SO-ARM101 scene_with_table_v, 2^20 envs fp64 sharded over the ranks, "
                            "train-shaped T=20 random-control rollouts (200 physics steps per env), dataset rows "
                            "[2^20,21,13] delivered to rank 0",
                "n_envs_total": N_TOTAL4, "n_envs_per_gpu": N_TOTAL4 // n_gpus, "control_steps": T4_TRAIN,
                "frame_skip": FRAME_SKIP, "seed": SEED,
                "parallelism": f"env-sharded x{n_gpus} (strong scaling), no step-path collective; rows stored straight into "
                               "rank 0's HBM by every rank's kernel (CUDA IPC peer mapping over NVLink)",
                "l2": "every step writes 286 MB .. 2.3 GB of rows per rank (>> 126 MB L2); no separate flush"}
    if cfg == "3":
        return {"workload": "BASELINE.json configs[2]: SO-ARM101 65536 envs fp32, 200 control steps x frame_skip 10, "
                            "random controls, state resident", "n_envs_per_gpu": 65536, "control_steps": 200,
                "frame_skip": FRAME_SKIP, "seed": SEED, "l2": "flushed between timed iterations (256 MiB write)"}
    if cfg == "5":
        return {"workload": "BASELINE.json configs[4]: 8192 control sequences x 50 env-steps from one shared state "
                            "(so101_batch_shoot), fp64", "n_envs_per_gpu": 8192, "control_steps": 50,
                "frame_skip": FRAME_SKIP, "seed": SEED, "l2": "flushed between timed iterations (256 MiB write)"}
    if cfg == "ik":
        return {"workload": "SURVEY 8f N3: 16384 Cartesian reference curves x 300 way-points -> joint tracks "
                            "(so101_ik_track), fp64", "tracks": 16384, "waypoints": 300,
                "l2": "flushed between timed iterations (256 MiB write)"}
    return {"workload": "BASELINE.json configs[1]: SO-ARM101 scene_with_table_v, 4096 envs/GPU fp64, "
                        "1000 physics steps (100 ctrl x frame_skip 10) random-control rollouts",
            "n_envs_per_gpu": N_ENVS, "n_envs_total": N_ENVS * n_gpus, "control_steps": T_CTRL,
            "frame_skip": FRAME_SKIP, "seed": SEED, "parallelism": f"env-sharded x{n_gpus}, no step-path collective",
            "l2": "flushed between timed iterations (256 MiB write)"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle on host cores
# ------------------------------------------------------------------------------------------------
def oracle_rate(n_envs: int, reps: int, warm: int, threads: int = 0, t_ctrl: int = T_CTRL):
    """-> (env-steps/s, seconds per rep, threads used) for `n_envs` x t_ctrl rollouts on the CPU oracle."""
    from lerobot_mujoco_sim2real_b200 import builtin_tables
    from oracle import oracle as O
    O.build()
    tables = builtin_tables("scene_with_table_v.xml")
    from lerobot_mujoco_sim2real_b200 import tables as T_
    O.set_hulls(T_.builtin_hulls())      # same physics as the CUDA arm: table-plane contact simulated
    # torchrun exports OMP_NUM_THREADS=1: ask the OS for the cores this process may use instead
    nthr = threads or len(os.sched_getaffinity(0))
    times = []
    for r in range(warm + reps):
        spec = O.make_spec(kind=0, seed=SEED + r)
        t0 = time.perf_counter()
        O.rollout(tables, spec, n_envs, t_ctrl, FRAME_SKIP, nthreads=nthr, want_rows=True)
        dt = time.perf_counter() - t0
        if r >= warm:
            times.append(dt)
    mean = sum(times) / len(times)
    return n_envs * t_ctrl / mean, mean, nthr


def cpu_baseline_for(cfg: str, budget_s: float = 12.0) -> dict:
    """Bounded sample (~budget_s of CPU work) of the workload on the oracle, all host threads."""
    t_ctrl = {"1": T_CTRL, "4": T4_TRAIN, "3": 200, "5": 50}.get(cfg, T_CTRL)
    full = {"1": N_ENVS, "4": N_TOTAL4, "3": 65536, "5": 8192}.get(cfg, N_ENVS)
    cal_rate, _, nthr = oracle_rate(64, 1, 1, t_ctrl=t_ctrl)
    n_sample = int(min(full, max(128, budget_s * cal_rate / t_ctrl)))
    rate, sec, nthr = oracle_rate(n_sample, 1, 0, t_ctrl=t_ctrl)
    return {"value": rate, "unit": UNIT, "cores": nthr, "kind": "port",
            "sample": f"{n_sample} of the {full} envs x {t_ctrl} control steps x {FRAME_SKIP} sub-steps "
                      f"({sec:.1f} s), same distributions, OpenMP over envs, fp64",
            "physics_steps_per_s": rate * FRAME_SKIP,
            "note": "restated-oracle CPU baseline (oracle/so101_oracle.c), not MuJoCo: the mujoco wheel is not installable offline"}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config if args.config in ("1", "4") else "1"
    t_ctrl = T4_TRAIN if cfg == "4" else T_CTRL
    full = N_TOTAL4 if cfg == "4" else N_ENVS
    n_sample = 2560 if cfg == "4" else 512   # x t_ctrl x 10 physics steps = 0.5 M mj_step-equivalents per step
    rate, sec, nthr = oracle_rate(n_sample, args.steps, args.warmup, t_ctrl=t_ctrl)
    sample = (f"{n_sample} of the {full} envs x {t_ctrl} control steps x {FRAME_SKIP} sub-steps per step "
              f"(same reset/control distributions), OpenMP over envs")
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong" if cfg == "4" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(cfg, args.gpus),
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
        self.index, self.proc, self.lines, self.first = index, None, [], 0

    def start(self):
        """Starts nvidia-smi and waits for its first line (its start-up can take longer than a short timed region), then
        marks where the samples taken under load begin."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 3.0:
                time.sleep(0.01)
            self.first = len(self.lines)
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self, busy=None) -> dict:
        """busy: callable that keeps the GPU under the same load; used to extend the sampling when the timed region was
        too short to catch two samples (the extension is not timed)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t0 = time.perf_counter()
        while busy is not None and len(self.lines) - self.first < 3 and time.perf_counter() - t0 < 2.0:
            busy()
        self.lines = self.lines[self.first:] if len(self.lines) > self.first else self.lines
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
class Ctx:
    """Process-group plumbing, device timing and clock sampling shared by the workloads."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.cuda.set_device(local_rank)
            # stdout carries exactly one JSON line: NCCL's own version / debug lines go to stderr
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"
                sys.stderr.write("bench.py: NCCL_DEBUG=VERSION -> WARN (keeps stdout to one JSON line)\n")
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            torch.cuda.set_device(0)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.K, self.W = args.steps, max(args.warmup, 0)
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        from lerobot_mujoco_sim2real_b200 import builtin_tables
        self.tables = builtin_tables("scene_with_table_v.xml")
        self.sampler = ClockSampler(self.dev.index)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, fn, k: int, w: int, flush: bool = True):
        """k timed calls of fn(step_index) after w warm-up calls; CUDA events around each call on the launching stream,
        L2 flushed in between (unless the step's own traffic exceeds L2).  The timed region is bracketed by barrier +
        synchronize.  -> (max over ranks of the mean per-call ms, wall ms of the bracketed region on this rank)"""
        torch = self.torch
        for i in range(w):
            fn(i)
        self.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
        t0 = time.perf_counter()
        for i in range(k):
            if flush:
                self.flush_buf.fill_(i & 0xFF)
            evs[i][0].record()
            fn(w + i)
            evs[i][1].record()
        self.barrier()
        wall = (time.perf_counter() - t0) * 1e3
        return self.max_over_ranks(sum(a.elapsed_time(b) for a, b in evs) / k), wall

    def peaks(self):
        from lerobot_mujoco_sim2real_b200.vec_env import fma_peak_tflops
        return fma_peak_tflops("float64", self.dev.index), fma_peak_tflops("float32", self.dev.index)

    def hbm_peak(self):
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            return 6650.0, "fallback (B200_PROFILING.md)"

    def finish(self, line):
        if self.rank == 0:
            print(json.dumps(line), flush=True)
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def roofline_block(ctx, kernel: str, family: str, phys_per_launch: float, kernel_ms: float, peak: float, clocks: dict,
                   dtype: str, bytes_per_env_step: float, env_steps_per_launch: float, note: str, stats=None,
                   traffic=None) -> dict:
    achieved = FLOP_PER_PHYSICS_STEP * phys_per_launch / (kernel_ms * 1e-3) / 1e12
    executed = FLOP_EXECUTED[family] * phys_per_launch / (kernel_ms * 1e-3) / 1e12
    hbm_peak, hbm_src = ctx.hbm_peak()
    lanes = 64 if dtype == "f64" else 128
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    nominal = 148 * lanes * 2 * sm_mhz * 1e6 / 1e12
    gbs = env_steps_per_launch * bytes_per_env_step / (kernel_ms * 1e-3) / 1e9
    out = {
        "bound": "fp64_pipe" if dtype == "f64" else "fp32_pipe", "kernel": kernel, "achieved": achieved, "peak": peak,
        "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
        "achieved_executed": executed, "frac_executed": executed / peak,
        "executed_flop_per_physics_step": FLOP_EXECUTED[family],
        "peak_source": "register-resident FMA loop measured in this run (so101_fma_peak), 'of measured'",
        "peak_nominal_at_sampled_clock": nominal, "frac_of_nominal": achieved / nominal,
        "algorithmic_flop_per_physics_step": FLOP_PER_PHYSICS_STEP, "physics_steps_per_launch": phys_per_launch,
        "kernel_ms": kernel_ms, "bound_note": note,
        "hbm_sanity": {"achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak, "peak_source": hbm_src,
                       "algorithmic_bytes_per_env_step": bytes_per_env_step},
    }
    if stats:
        out["newton_iters_per_physics_step"] = stats["newton_iters"] / max(1, stats["physics_steps"])
        out["ls_evals_per_physics_step"] = stats["ls_evals"] / max(1, stats["physics_steps"])
    return out


# ---- config 1: the headline (and weak scaling when forced at N > 1) ------------------------------------------------
def wl_config1(ctx) -> None:
    torch, dist = ctx.torch, ctx.dist
    from lerobot_mujoco_sim2real_b200 import tables as T
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    dev, rank, world, K, W, tables = ctx.dev, ctx.rank, ctx.world, ctx.K, ctx.W, ctx.tables
    env = SOARM101VecEnv(tables=tables, num_envs=N_ENVS, dtype="float64", device=dev.index, seed=SEED)
    rows = torch.empty((N_ENVS, T_CTRL + 1, T.ROW), dtype=torch.float64, device=dev)
    env_offset = rank * N_ENVS

    def step_dev(i):
        env.rollout(T_CTRL, "random", seed=SEED + i, env_offset=env_offset, out=rows)

    step_dev(0)
    torch.cuda.synchronize()
    env.stats()
    if rank == 0:
        ctx.sampler.start()
    ms_per_step, wall_ms = ctx.timed(step_dev, K, W)
    clocks = ctx.sampler.stop(busy=lambda: (step_dev(0), torch.cuda.synchronize())) if rank == 0 else None
    st = env.stats()
    flags = env.flags()
    n_trip = int((flags & T.FLAG_TRIP).ne(0).sum().item())
    n_bad = int((flags & T.FLAG_BADSTATE).ne(0).sum().item())
    n_contact = int((flags & T.FLAG_CONTACT).ne(0).sum().item()) if hasattr(T, "FLAG_CONTACT") else None
    env_steps_per_launch = N_ENVS * T_CTRL
    value = world * env_steps_per_launch / (ms_per_step * 1e-3)
    phys_per_launch = env_steps_per_launch * FRAME_SKIP

    # ---- e2e: host buffers through the C ABI ------------------------------------------------------
    # Five different input sets, used in turn: the launch time depends on which envs reach the table (6.2 .. 8.2 ms over the
    # control seeds of the device-resident region), so a single fixed input would make e2e a statement about one seed.
    g = torch.Generator().manual_seed(SEED + rank)
    inputs = []
    for _ in range(5):
        U_host = ((torch.rand((T_CTRL + 1, T.NU_ENV, N_ENVS), generator=g, dtype=torch.float64) - 0.5)).pin_memory()
        q0_host = torch.zeros((T.NV, N_ENVS), dtype=torch.float64)
        q0_host[:5] = (torch.rand((5, N_ENVS), generator=g, dtype=torch.float64) - 0.5) * 0.6
        inputs.append((U_host, q0_host.pin_memory()))
    rows_host = torch.empty((N_ENVS, T_CTRL + 1, T.ROW), dtype=torch.float64).pin_memory()

    def step_e2e(i):
        # ONE C-ABI call with host buffers: H2D controls + initial angles, k_reset, k_rollout storing the rows to the host, sync
        U_host, q0_host = inputs[i % len(inputs)]
        env.rollout_host(T_CTRL, "tensor", u_host=U_host, qpos0_host=q0_host, out_host=rows_host)

    e2e_ms, _ = ctx.timed(step_e2e, K, W)
    e2e_value = world * env_steps_per_launch / (e2e_ms * 1e-3)
    # the same five input sets with everything resident: what the launch alone takes on them (no collective in here)
    zeros6 = torch.zeros((N_ENVS, T.NV), dtype=torch.float64, device=dev)
    same = []
    for U_host, q0_host in inputs:
        Ud, q0d = U_host.to(dev), q0_host.to(dev).t().contiguous()
        for rep in range(2):
            env.set_state(q0d, zeros6, zeros6)
            ctx.flush_buf.fill_(rep)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            env.rollout(T_CTRL, "tensor", u=Ud, out=rows, flags=T.ROLL_NO_RESET)
            b.record()
            torch.cuda.synchronize()
        same.append(a.elapsed_time(b))
    e2e_kernel_ms = sum(same) / len(same)
    h2d = U_host.numel() * 8 + q0_host.numel() * 8
    d2h = rows_host.numel() * 8

    if rank != 0:
        ctx.finish(None)
        return

    peak64, peak32 = ctx.peaks()
    traffic = None
    for name in ("r2_kernel_traffic.json", "r1_kernel_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch_config1")
                break
            except Exception:
                traffic = None
    roofline = roofline_block(
        ctx, "k_rollout<double,double,SPLIT=true> (three-warp team kernel, batches <= 9472 envs)", "team",
        phys_per_launch, ms_per_step, peak64, clocks, "f64", 208.0, env_steps_per_launch,
        "neither hbm nor tensor: ~300 FLOP per byte of state/row traffic, no contraction; the limiter is the FP64 FMA "
        "pipe.  At this config (4096 envs = 128 three-warp teams, one per SM, on 148 SMs) the launch is latency-bound: a "
        "physics step is one dependent instruction chain on the team's dynamics warp; the large-batch fraction is in "
        "extra / --config 4", st, traffic)

    # like for like with round 1 (contact-free physics: envs that reach the table are only flagged): the same batch and seeds
    # with the hull data left out.  Not the product's physics; it separates what the contact path costs from the rest.
    env_cf = SOARM101VecEnv(tables=tables, num_envs=N_ENVS, dtype="float64", device=dev.index, seed=SEED, hulls=None)
    cf = []
    for i in range(3 + min(K, 10)):                        # rank 0 only: no barrier in here
        ctx.flush_buf.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        env_cf.rollout(T_CTRL, "random", seed=SEED + i, env_offset=env_offset, out=rows)
        b.record()
        torch.cuda.synchronize()
        cf.append(a.elapsed_time(b))
    cf_ms = sum(cf[3:]) / len(cf[3:])
    cf_flagged = int((env_cf.flags() & T.FLAG_TRIP).ne(0).sum().item())
    del env_cf
    roofline["contact_free_same_batch"] = {
        "kernel_ms": cf_ms, "frac": FLOP_PER_PHYSICS_STEP * phys_per_launch / (cf_ms * 1e-3) / 1e12 / peak64,
        "envs_flagged_not_simulated": cf_flagged,
        "note": "round-1 physics (table contact only flagged) on the same batch and seeds: the launch lasts as long as its "
                "slowest team, and with contact simulated that team carries an env that keeps touching the table "
                "(a contact step costs its team 2.5x a plain one, DESIGN 4a)"}

    extra = {}
    if not ctx.args.no_extra:
        extra = config1_extras(ctx, peak64, peak32)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config("1", world),
        "physics_steps_per_s": value * FRAME_SKIP,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms, "kernel_ms_same_inputs": e2e_kernel_ms,
                "path": "so101_batch_rollout_host: pinned host control tensor [T+1,5,N] + initial angles [6,N] uploaded -> "
                        "k_reset + ONE k_rollout launch (SO101_CTRL_TENSOR) whose row writer stores the dataset rows "
                        "[N,T+1,13] straight into the caller's pinned host buffer over PCIe while it computes (no staging "
                        "copy, no download phase; datasets >= 128 MB take the time-chunked copy-engine pipeline instead) "
                        "-> stream sync.  d2h_bytes_per_step = the bytes those stores carry"},
        "gpu_launches": K,
        "gpu_launches_note": "one k_rollout<double,double,true> launch per step in the device-resident region; "
                             "the e2e region launches k_reset + one k_rollout per step",
        "roofline": roofline, "cpu_baseline": cpu_baseline_for("1"), "clocks": clocks,
        "wall_ms_timed_region": wall_ms,
        "flags": {"envs_tripwire": n_trip, "envs_badstate": n_bad, "envs_in_contact": n_contact, "of": N_ENVS},
        "extra": extra,
    }
    ctx.finish(line)


def config1_extras(ctx, peak64, peak32) -> dict:
    """Secondary single-GPU measurements riding on the default run (the driver drops `extra`; every one of them is also
    a first-class workload through --config)."""
    torch = ctx.torch
    from lerobot_mujoco_sim2real_b200 import tables as T
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    dev, tables = ctx.dev, ctx.tables
    extra = {}

    def quick(n, dtype, t_ctrl, reps=3):
        e = SOARM101VecEnv(tables=tables, num_envs=n, dtype=dtype, device=dev.index, seed=SEED)
        e.rollout_discard(2, "random")
        torch.cuda.synchronize()
        best = 1e30
        for r in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.flush_buf.fill_(r)
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
    import contextlib
    import types
    # BASELINE configs[0] workload through the drop-in API: the reference's whole data job with args.py defaults
    from lerobot_mujoco_sim2real_b200.SOARM101_DataCollection import SOARM101DataGenerator
    jargs = types.SimpleNamespace(xml_path="unused", x_dim=8, u_dim=5, device="cuda", seed=SEED, env="SOARM101")
    with contextlib.redirect_stdout(sys.stderr):
        jgen = SOARM101DataGenerator(jargs, tables=tables, device=dev.index)
        job = [(50000, 20, "random"), (2000, 200, "random"), (2000, 200, "random"), (2000, 200, "sin"), (2000, 200, "chirp")]
        jgen.generate_physics_based_data(64, 2, "random")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nbytes, replaced = 0, 0
        for (jn, jt, jk) in job:
            arr = jgen.generate_physics_based_data(jn, jt, jk, seed=SEED)
            nbytes += arr.nbytes
            replaced += jgen.last_replaced
        job_s = time.perf_counter() - t0
    job_steps = sum(jn * jt for jn, jt, _ in job)
    extra["config0_reference_data_job_args_py_defaults"] = {
        "datasets": [f"{jn}x{jt} {jk}" for jn, jt, jk in job], "env_steps": job_steps, "wall_ms": job_s * 1e3,
        "env_steps_per_s": job_steps / job_s, "host_bytes_out": nbytes, "trajectories_replaced": replaced,
        "note": "SOARM101DataGenerator.generate_physics_based_data for the five datasets generate_and_save_data builds "
                "(host numpy arrays out, flagged trajectories regenerated); the reference runs this serially on one CPU MuJoCo env"}
    del jgen
    return extra


# ---- config 4: the sharded 2^20-env dataset job (strong scaling) ----------------------------------------------------
def wl_config4(ctx) -> None:
    torch, dist = ctx.torch, ctx.dist
    from lerobot_mujoco_sim2real_b200 import sharding, tables as T
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    dev, rank, world, K, W, tables = ctx.dev, ctx.rank, ctx.world, ctx.K, ctx.W, ctx.tables
    lo, hi = sharding.shard_range(N_TOTAL4, rank, world)
    n_local = hi - lo
    env = SOARM101VecEnv(tables=tables, num_envs=n_local, dtype="float64", device=dev.index, seed=SEED)
    res = {}
    clocks = None
    for name, kind, Tn, k, w in (("train_T20_random", "random", T4_TRAIN, K, W),
                                 ("test_T200_chirp", "chirp", T4_TEST, max(1, min(K, 3)), 1)):
        row_bytes = (Tn + 1) * T.ROW * 8
        local = torch.empty((n_local, Tn + 1, T.ROW), dtype=torch.float64, device=dev)

        def step_sim(i, kind=kind, Tn=Tn, local=local):
            env.rollout(Tn, kind, seed=SEED + i, env_offset=lo, out=local)

        step_sim(0)
        torch.cuda.synchronize()
        env.stats()
        if rank == 0 and name.startswith("train"):
            ctx.sampler.start()
        sim_ms, sim_wall = ctx.timed(step_sim, k, w, flush=False)
        if rank == 0 and name.startswith("train"):
            clocks = ctx.sampler.stop()
        st = env.stats()
        fl = env.flags()
        n_trip = ctx.sum_over_ranks(float((fl & T.FLAG_TRIP).ne(0).sum().item()))
        n_contact = (ctx.sum_over_ranks(float((fl & T.FLAG_CONTACT).ne(0).sum().item()))
                     if hasattr(T, "FLAG_CONTACT") else None)
        del local
        torch.cuda.empty_cache()
        # e2e: every rank's kernel stores its rows straight into rank 0's buffer; a step ends when all rows are there
        shared = sharding.SharedRows(N_TOTAL4, (Tn + 1, T.ROW), torch.float64, dev.index, dst=0)

        def step_e2e(i, kind=kind, Tn=Tn, shared=shared):
            env.rollout(Tn, kind, seed=SEED + i, env_offset=lo, out_ptr=shared.local_ptr)

        for i in range(w):
            step_e2e(i)
        ctx.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for i in range(k):
            step_e2e(w + i)
            if world > 1:
                torch.cuda.current_stream().synchronize()   # my rows have landed ...
                dist.barrier()                              # ... and everybody else's: the dataset is complete on rank 0
        b.record()
        ctx.barrier()
        e2e_wall = (time.perf_counter() - t0) * 1e3 / k
        e2e_ms = ctx.max_over_ranks(max(a.elapsed_time(b) / k, e2e_wall if world > 1 else 0.0))
        # check the delivered dataset against a re-simulated slice of the LAST rank's shard (rank 0 alone)
        full = shared.finish()
        check = None
        if rank == 0:
            nchk = 2048
            lo_last = sharding.shard_range(N_TOTAL4, world - 1, world)[0]
            e_chk = SOARM101VecEnv(tables=tables, num_envs=nchk, dtype="float64", device=dev.index, seed=SEED)
            ref = e_chk.rollout(Tn, kind, seed=SEED + w + k - 1, env_offset=lo_last)
            check = bool(torch.equal(ref, full[lo_last:lo_last + nchk]))
            del e_chk, ref
        del full
        shared.close()
        env_steps = N_TOTAL4 * Tn
        res[name] = {"control_steps": Tn, "kind": kind, "steps_timed": k, "sim_ms": sim_ms, "e2e_ms": e2e_ms,
                     "sim_env_steps_per_s": env_steps / (sim_ms * 1e-3), "e2e_env_steps_per_s": env_steps / (e2e_ms * 1e-3),
                     "rows_bytes_total": N_TOTAL4 * row_bytes,
                     "nvlink_bytes_into_rank0_per_step": (N_TOTAL4 - (sharding.shard_range(N_TOTAL4, 0, world)[1])) * row_bytes,
                     "delivery_overhead_ms": e2e_ms - sim_ms, "last_rank_slice_bit_identical": check,
                     "envs_tripwire": n_trip, "envs_in_contact": n_contact, "stats": st, "wall_ms_sim_region": sim_wall}
    if rank != 0:
        ctx.finish(None)
        return
    # same workload on ONE GPU in this run (rank 0 alone, train job): the base of the strong-scaling ratio
    single = None
    if world > 1:
        e1 = SOARM101VecEnv(tables=tables, num_envs=N_TOTAL4, dtype="float64", device=dev.index, seed=SEED)
        r1 = torch.empty((N_TOTAL4, T4_TRAIN + 1, T.ROW), dtype=torch.float64, device=dev)
        e1.rollout(T4_TRAIN, "random", seed=SEED, out=r1)
        torch.cuda.synchronize()
        best = 1e30
        for i in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); e1.rollout(T4_TRAIN, "random", seed=SEED + i, out=r1); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        single = {"ms": best, "env_steps_per_s": N_TOTAL4 * T4_TRAIN / (best * 1e-3)}
        del e1, r1
    peak64, _ = ctx.peaks()
    tr = res["train_T20_random"]
    phys_per_launch = n_local * T4_TRAIN * FRAME_SKIP
    roofline = roofline_block(
        ctx, "k_rollout<double,double,SPLIT=false> (one thread per env, 256-thread blocks)", "onewarp", phys_per_launch,
        tr["sim_ms"], peak64, clocks, "f64", 208.0, n_local * T4_TRAIN,
        "FP64-pipe bound (ncu: pipe ~45 % busy, issue slots ~39 %, top stall = fixed-latency FP64 dependencies at 8 "
        "warps/SM x 255 registers); HBM traffic is the rows only", tr["stats"])
    line = {
        "metric": METRIC, "value": tr["sim_env_steps_per_s"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": tr["sim_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config("4", world), single_gpu_same_workload=single,
                       test_T200_chirp={k_: v for k_, v in res["test_T200_chirp"].items() if k_ != "stats"}),
        "physics_steps_per_s": tr["sim_env_steps_per_s"] * FRAME_SKIP,
        "e2e": {"value": tr["e2e_env_steps_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "ms_per_step": tr["e2e_ms"], "gather_ms": tr["delivery_overhead_ms"],
                "gather_bytes_over_nvlink": tr["nvlink_bytes_into_rank0_per_step"],
                "rows_bytes_total": tr["rows_bytes_total"], "delivered_bit_identical": tr["last_rank_slice_bit_identical"],
                "test_T200_chirp": {"value": res["test_T200_chirp"]["e2e_env_steps_per_s"],
                                    "sim_value": res["test_T200_chirp"]["sim_env_steps_per_s"],
                                    "ms_per_step": res["test_T200_chirp"]["e2e_ms"],
                                    "gather_ms": res["test_T200_chirp"]["delivery_overhead_ms"],
                                    "gather_bytes_over_nvlink": res["test_T200_chirp"]["nvlink_bytes_into_rank0_per_step"]},
                "path": "simulation + delivery of the whole dataset to rank 0: controls are generated on the device "
                        "(Philox / chirp generators keyed by global env id: no host inputs), every rank's k_rollout writes "
                        "its rows straight into rank 0's [2^20,T+1,13] buffer through a CUDA-IPC peer mapping (NVLink), "
                        "stream sync + one host barrier per step; the result stays in rank 0's HBM (np.save would add a "
                        "D2H copy of the dataset, not part of the metric)"},
        "gpu_launches": K,
        "gpu_launches_note": "one k_rollout<double,double,false> launch per rank and step, in both regions",
        "roofline": roofline, "cpu_baseline": cpu_baseline_for("4"), "clocks": clocks,
        "flags": {"envs_tripwire": tr["envs_tripwire"], "envs_in_contact": tr["envs_in_contact"], "of": N_TOTAL4},
    }
    ctx.finish(line)


# ---- configs 3 / 5 / ik as first-class single-GPU lines --------------------------------------------------------------
def wl_single(ctx, cfg: str) -> None:
    torch = ctx.torch
    from lerobot_mujoco_sim2real_b200 import tables as T
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    if ctx.world != 1:
        raise SystemExit(f"--config {cfg} is a single-GPU workload")
    dev, K, W, tables = ctx.dev, ctx.K, ctx.W, ctx.tables
    peak64, peak32 = ctx.peaks()
    ctx.sampler.start()
    if cfg == "3":
        n, Tn = 65536, 200
        env = SOARM101VecEnv(tables=tables, num_envs=n, dtype="float32", device=dev.index, seed=SEED)
        rows = torch.empty((n, Tn + 1, T.ROW), dtype=torch.float32, device=dev)
        ms, wall = ctx.timed(lambda i: env.rollout_discard(Tn, "random", seed=SEED + i), K, W)
        clocks = ctx.sampler.stop()
        st = env.stats()
        ms_rows, _ = ctx.timed(lambda i: env.rollout(Tn, "random", seed=SEED + i, out=rows, flags=T.ROLL_ROWS_F32), max(1, K // 4), 1)
        rows_host = torch.empty((n, Tn + 1, T.ROW), dtype=torch.float32).pin_memory()
        e2e_ms, _ = ctx.timed(lambda i: env.rollout_host(Tn, "random", seed=SEED + i, out_host=rows_host, flags=T.ROLL_ROWS_F32),
                              max(1, K // 4), 1)
        value, steps_per = n * Tn / (ms * 1e-3), n * Tn
        roofline = roofline_block(ctx, "k_rollout<float,*,SPLIT=false> (one thread per env, 512-thread blocks)", "onewarp",
                                  steps_per * FRAME_SKIP, ms, peak32, clocks, "f32", 52.0, steps_per,
                                  "FP32-pipe bound; state stays in registers for the whole launch", st)
        line = {"dtype": "f32", "value": value, "ms_per_step": ms,
                "e2e": {"value": steps_per / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": rows_host.numel() * 4, "ms_per_step": e2e_ms,
                        "device_rows_value": steps_per / (ms_rows * 1e-3),
                        "path": "so101_batch_rollout_host, controls generated on the device, float32 rows [N,T+1,13] "
                                "downloaded into pinned host memory in pipelined time chunks"},
                "roofline": roofline, "cpu_baseline": cpu_baseline_for("3")}
    elif cfg == "5":
        n, H = 8192, 50
        env = SOARM101VecEnv(tables=tables, num_envs=n, dtype="float64", device=dev.index, seed=SEED)
        U = (torch.rand((H, T.NU_ENV, n), dtype=torch.float64, device=dev) - 0.5).contiguous()
        s0 = [0.1, -0.2, 0.15, 0.05, -0.1, 0.0] + [0.0] * 12
        ms, wall = ctx.timed(lambda i: env.shoot(s0, U, flags=T.ROLL_GRAVCOMP_HOLD), K, W)
        clocks = ctx.sampler.stop()
        st = env.stats()
        U_host = U.cpu().pin_memory()
        X_host = torch.empty((n, H + 1, T.NOBS), dtype=torch.float32).pin_memory()

        def e2e(i):
            Ud = U_host.to(dev, non_blocking=True)
            X = env.shoot(s0, Ud, flags=T.ROLL_GRAVCOMP_HOLD)
            X_host.copy_(X, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e_ms, _ = ctx.timed(e2e, K, W)
        value, steps_per = n * H / (ms * 1e-3), n * H
        roofline = roofline_block(ctx, "k_shoot<double,SPLIT=true> (three-warp team kernel)", "team", steps_per * FRAME_SKIP,
                                  ms, peak64, clocks, "f64", 72.0, steps_per,
                                  "latency-bound like config 1 (8192 envs = 256 teams on 148 SMs)", st)
        line = {"dtype": "f64", "value": value, "ms_per_step": ms, "rollouts_per_s": n / (ms * 1e-3),
                "e2e": {"value": steps_per / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": U_host.numel() * 8,
                        "d2h_bytes_per_step": X_host.numel() * 4, "ms_per_step": e2e_ms,
                        "path": "pinned control sequences [H,5,B] H2D, so101_batch_shoot, observations [B,H+1,8] f32 D2H"},
                "roofline": roofline, "cpu_baseline": cpu_baseline_for("5")}
    else:
        import numpy as np
        from lerobot_mujoco_sim2real_b200.TrajectoryGenerator import CartesianTrajectoryGenerator, reference_curve
        gen = CartesianTrajectoryGenerator(tables=tables, device=dev.index)
        nik = 16384
        base = np.stack([reference_curve(nm, ix)[0] for nm in ("Fig8", "Circle") for ix in (0, 1)])
        xyz_h = torch.as_tensor(base[np.arange(nik) % 4] + np.random.default_rng(SEED).uniform(-0.03, 0.03, (nik, 1, 3))).pin_memory()
        xyz = xyz_h.to(dev)
        out = {}
        ms, wall = ctx.timed(lambda i: out.__setitem__("r", gen.solve_tracks(xyz)), K, W)
        clocks = ctx.sampler.stop()
        qik, stik = out["r"]

        def e2e(i):
            q, s_ = gen.solve_tracks(xyz_h.to(dev, non_blocking=True))
            out["h"] = (q.cpu(), s_.cpu())
        e2e_ms, _ = ctx.timed(e2e, max(1, K // 4), 1)
        iters = float((stik >> 8).double().mean().item())
        line = {"metric": "way-points/sec", "unit": "way-points/s", "dtype": "f64", "value": nik * 300 / (ms * 1e-3),
                "ms_per_step": ms, "gauss_newton_iters_per_waypoint": iters,
                "success_frac": float((stik & 1).double().mean().item()),
                "e2e": {"value": nik * 300 / (e2e_ms * 1e-3), "unit": "way-points/s", "h2d_bytes_per_step": xyz_h.numel() * 8,
                        "d2h_bytes_per_step": out["h"][0].numel() * 8 + out["h"][1].numel() * 4, "ms_per_step": e2e_ms,
                        "path": "pinned way-points H2D, relayout + so101_ik_track, joint tracks + status D2H"},
                "roofline": {"bound": "fp64_pipe", "kernel": "k_ik_track<false>", "achieved": 1.18e3 * 64 * nik * 300 * iters / (ms * 1e-3) / 1e12,
                             "peak": peak64, "unit": "TFLOP/s", "frac": 1.18e3 * 64 * nik * 300 * iters / (ms * 1e-3) / 1e12 / peak64 / 32,
                             "traffic": None, "note": "1.18 k warp instructions per FK + solve iteration (ncu), counted as 2 FLOP per lane "
                                                      "of FP64 work: an instruction-count bound, not an algorithmic FLOP count"},
                "cpu_baseline": None}
    base_line = {"metric": METRIC, "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W, "higher_is_better": True,
                 "scaling": "weak", "vs_baseline": None, "data": "synthetic", "config": workload_config(cfg, 1),
                 "gpu_launches": K, "clocks": clocks, "wall_ms_timed_region": wall}
    base_line.update(line)
    if cfg != "ik":
        base_line["physics_steps_per_s"] = base_line["value"] * FRAME_SKIP
    ctx.finish(base_line)


def run_b200(args) -> None:
    ctx = Ctx(args)
    if args.config == "1":
        wl_config1(ctx)
    elif args.config == "4":
        wl_config4(ctx)
    else:
        wl_single(ctx, args.config)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
