"""GPU: the BASELINE.json configurations at their FULL sizes, through the C ABI.

Where the CPU oracle finishes in seconds (config 1) the comparison is direct; at the sizes it cannot reach
(65 536 fp32 envs, 2^20 envs, 8192 x 50 shooting) the checks are size-independent properties: determinism,
invariance to how the batch is sharded / which kernel family runs, agreement of independent kernels with each
other (shoot vs rollout), plus oracle comparisons on sampled sub-ranges.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _vec(tables, n, dtype="float64", **kw):
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    return SOARM101VecEnv(tables=tables, num_envs=n, dtype=dtype, **kw)


def _rel(a, b):
    return np.abs(a - b) / (1e-3 + np.maximum(np.abs(a), np.abs(b)))


# ------------------------------------------------------------------------------------------------
# config 1: 4096 envs fp64, 1000-step random-control rollouts - teacher-forced against the oracle at full size
# ------------------------------------------------------------------------------------------------
def test_config1_full_size_teacher_forced(oracle_mod, tables_v):
    O = oracle_mod
    n, steps = 4096, 1000
    rng = np.random.default_rng(42)
    state = np.zeros((n, 18)); state[:, :5] = rng.uniform(-0.3, 0.3, (n, 5))
    s_in = np.empty((steps, n, 18)); s_ref = np.empty((steps, n, 18)); ctrls = np.empty((steps, n, 6))
    ctrl = np.zeros((n, 6))
    for t in range(steps):
        if t % 10 == 0:
            ctrl = np.zeros((n, 6)); ctrl[:, :5] = rng.uniform(-0.5, 0.5, (n, 5))
        s_in[t], ctrls[t] = state, ctrl
        state, _, _ = O.step_batch(tables_v, state, ctrl, 1)
        s_ref[t] = state
    env = _vec(tables_v, n * steps)                       # 4.1 M states replayed in one launch
    flat = s_in.reshape(-1, 18)
    env.set_state(flat[:, :6], flat[:, 6:12], flat[:, 12:18])
    u = torch.as_tensor(ctrls.reshape(-1, 6).T.copy(), device=env.device).contiguous()
    env.step_soa(u, 1)
    q, v, w = env.get_state()
    out = np.concatenate([q.cpu().numpy(), v.cpu().numpy(), w.cpu().numpy()], axis=1)
    trip = (env.flags().cpu().numpy() & 0x6) != 0          # contact tripwire: claims are over flag-free envs
    err = _rel(out, s_ref.reshape(-1, 18))[~trip]
    print(f"config 1 full size: {(~trip).sum()} flag-free states of {n * steps}; max rel err qpos {err[:, :6].max():.2e} "
          f"qvel {err[:, 6:12].max():.2e} (99.9 %: {np.quantile(err[:, 6:12], 0.999):.2e}) qacc {err[:, 12:].max():.2e}")
    assert (~trip).mean() > 0.9
    assert err[:, :6].max() < 1e-12
    assert np.quantile(err[:, 6:12], 0.999) < 1e-12 and err[:, 6:12].max() < 1e-10
    assert np.quantile(err[:, 12:], 0.99) < 1e-10 and err[:, 12:].max() < 1e-7


# ------------------------------------------------------------------------------------------------
# config 2: 65 536 envs fp32, 200 env-steps
# ------------------------------------------------------------------------------------------------
def test_config2_full_size_fp32(oracle_mod, tables_v):
    from lerobot_mujoco_sim2real_b200 import tables as T_
    O = oracle_mod
    n, T = 65536, 200
    env = _vec(tables_v, n, "float32")
    rows = env.rollout(T, "random", seed=42, flags=T_.ROLL_ROWS_F32)
    q, v, _ = env.get_state()
    fl = env.flags()
    assert env.stats()["physics_steps"] == n * T * 10
    assert torch.isfinite(rows).all() and torch.isfinite(v).all()
    assert int((fl & T_.FLAG_BADSTATE).ne(0).sum()) == 0
    # determinism, and invariance to sharding / kernel family: 8 shards of 8192 envs run the team kernels
    rows2 = _vec(tables_v, n, "float32").rollout(T, "random", seed=42, flags=T_.ROLL_ROWS_F32)
    assert torch.equal(rows, rows2)
    del rows2
    for r in (0, 5):
        shard = _vec(tables_v, n // 8, "float32").rollout(T, "random", seed=42, env_offset=r * (n // 8), flags=T_.ROLL_ROWS_F32)
        assert torch.equal(rows[r * (n // 8):(r + 1) * (n // 8)], shard)
    # joint angles stay inside the joint ranges (+ the limit rows' penetration allowance)
    lo = torch.tensor([tables_v.jnt_range[k][0] for k in range(6)], device=q.device) - 0.02
    hi = torch.tensor([tables_v.jnt_range[k][1] for k in range(6)], device=q.device) + 0.02
    assert bool(((q >= lo) & (q <= hi)).all())
    # against the fp64 oracle before chaos amplifies fp32 rounding: first 3 control steps, 512 envs
    ref, _, _ = O.rollout(tables_v, O.make_spec(kind=0, seed=42, env_offset=1000), 512, 3, 10)
    got = rows[1000:1512, :4].cpu().numpy().astype(np.float64)
    np.testing.assert_allclose(got[:, :, :5], ref[:, :, :5], atol=1e-7)      # controls (f32-rounded Philox stream)
    assert np.abs(got[:, :, 5:] - ref[:, :, 5:]).max() < 5e-5                # stated fp32 tolerance on observations


# ------------------------------------------------------------------------------------------------
# config 3/4: 2^20 envs fp64, train-shaped T=20, sharded 8 ways
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["random", "chirp"])
def test_config4_full_size_one_million_envs(oracle_mod, tables_v, kind):
    O = oracle_mod
    n, T, G = 1 << 20, 20, 8
    rows = _vec(tables_v, n).rollout(T, kind, seed=42)                       # 2.3 GB of rows, one launch
    assert rows.shape == (n, T + 1, 13)
    shard = n // G
    for r in (0, 3, 7):                                                      # what rank r of 8 would generate
        part = _vec(tables_v, shard).rollout(T, kind, seed=42, env_offset=r * shard)
        assert torch.equal(rows[r * shard:(r + 1) * shard], part)
        del part
    # a 4096-env rank (team kernels) generates the same bits as the 2^20-env launch (one-warp kernels)
    small = _vec(tables_v, 4096).rollout(T, kind, seed=42, env_offset=777_777)
    assert torch.equal(rows[777_777:777_777 + 4096], small)
    # sampled sub-range against the CPU restatement of generate_physics_based_data
    k = {"random": 0, "sin": 1, "chirp": 2}[kind]
    ref, _, _ = O.rollout(tables_v, O.make_spec(kind=k, seed=42, env_offset=500_000), 256, T, 10)
    got = rows[500_000:500_256].cpu().numpy()
    if kind == "random":
        np.testing.assert_array_equal(got[:, :, :5], ref[:, :, :5])
    else:
        np.testing.assert_allclose(got[:, :, :5], ref[:, :, :5], atol=1e-15)
    err = np.abs(got[:, :, 5:] - ref[:, :, 5:])
    assert err[:, :6].max() <= 6e-8                        # first 50 physics steps: at most one float32 ulp flip
    # 200 free-running steps on scene A: chaos has started to act (P2 curve: a 1-ulp perturbation is 5e-11 at
    # step 100 and 2e-3 at step 300), so only the bulk is required to agree
    print(f"config 4 {kind}: obs err vs oracle, rows 0-10: max {err[:, :11].max():.1e}; rows 11-20: median {np.median(err[:, 11:]):.1e} "
          f"99 % {np.quantile(err[:, 11:], 0.99):.1e} max {err.max():.1e}")
    assert np.quantile(err[:, :11], 0.999) <= 1e-6 and err[:, :11].max() < 1e-3
    assert np.median(err[:, 11:]) <= 1e-4 and err.max() < 5e-2


# ------------------------------------------------------------------------------------------------
# config 5: 8192 control sequences x 50-step horizon from a shared state
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("gravcomp", [False, True])
def test_config5_full_size_shoot_equals_rollout(oracle_mod, tables_v, gravcomp):
    from lerobot_mujoco_sim2real_b200 import tables as T_
    O = oracle_mod
    B, H = 8192, 50
    flags = T_.ROLL_GRAVCOMP_HOLD if gravcomp else 0
    rng = np.random.default_rng(42)
    s0 = np.zeros(18); s0[:5] = rng.uniform(-0.3, 0.3, 5)
    s0 = O.step_batch(tables_v, s0[None], np.zeros((1, 6)), 100)[0][0]       # 10 warm-up env-steps
    U = torch.as_tensor(rng.uniform(-0.5, 0.5, (H, 5, B))).cuda().contiguous()
    env = _vec(tables_v, B)
    X = env.shoot(s0, U, flags=flags)
    assert X.shape == (B, H + 1, 8) and env.stats()["physics_steps"] == B * H * 10
    # the same sequences through the dataset kernel: state broadcast, controls as a tensor (u_t applied at step t+1)
    env2 = _vec(tables_v, B)
    bc = lambda x: torch.as_tensor(np.tile(x, (B, 1)))
    env2.set_state(bc(s0[:6]), bc(s0[6:12]), bc(s0[12:18]))
    Upad = torch.cat([U, torch.zeros((1, 5, B), dtype=U.dtype, device=U.device)]).contiguous()
    rows = env2.rollout(H, "tensor", u=Upad, flags=flags | T_.ROLL_NO_RESET)
    assert torch.equal(X, rows[:, :, 5:].to(torch.float32))
    for a, b in zip(env.get_state(), env2.get_state()):
        assert torch.equal(a, b)
    # and 32 of the sequences against the oracle over the first 5 env-steps (before chaos acts)
    ref = O.shoot(tables_v, s0, U[:5, :, :32].cpu().numpy(), 10, flags=flags)
    assert np.abs(X[:32, :6].cpu().numpy() - ref).max() <= 2e-7


# ------------------------------------------------------------------------------------------------
# config 4's wording, end to end: TrajectoryGenerator-driven closed-loop data at 2^20 curves
# ------------------------------------------------------------------------------------------------
def _curve_params(lo, hi):
    """curve family / centre / scale as a pure function of the GLOBAL curve index (what makes the dataset shard-invariant)"""
    i = np.arange(lo, hi, dtype=np.uint64)
    h = (i * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(11)
    u = [((h >> np.uint64(8 * k)) & np.uint64(0xFF)).astype(np.float64) / 255.0 for k in range(4)]
    centers = np.stack([0.4 + 0 * u[1], 0.03 * (u[1] - 0.5), 0.2 + 0.03 * (u[2] - 0.5)], axis=1)    # y-z plane (idx = 1)
    return u[0] < 0.5, centers, 0.4 + 0.2 * u[3]


def _mpc_pipeline(tables, lo, hi, frames, P, MPC_type="delta_mpc"):
    """curves [lo, hi) -> way-points on the device -> IK (one launch) -> Koopman_MPC loop -> rows [n, frames, 13]"""
    from lerobot_mujoco_sim2real_b200.Koopman_MPC import BatchedKoopmanMPC
    from lerobot_mujoco_sim2real_b200.TrajectoryGenerator import CartesianTrajectoryGenerator
    from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    n = hi - lo
    km = KoopmanModel.from_npz(os.path.join(ROOT, "tests", "golden", "koopman_dkuc.npz"))
    gen = CartesianTrajectoryGenerator(tables=tables)
    names, centers, scale = _curve_params(lo, hi)
    xyz, q, st = gen.generate_batch(names, idx=np.ones(n, dtype=np.int64), traj_scale=scale, centers=centers)
    xyz, q = xyz[:, :P].contiguous(), q[:, :P].contiguous()                 # the first P way-points of every curve
    env = SOARM101VecEnv(tables=tables, num_envs=n, gravity_compensation=True)
    loop = BatchedKoopmanMPC(env, km, xyz, q, H=10, MPC_type=MPC_type)
    actual = loop.run(frames)
    return loop.dataset_rows(), actual, xyz, st[:, :P]


def test_config4_size_trajectory_generator_pipeline(tables_v):
    """2^20 Cartesian curves -> batched IK -> the reference's Koopman_MPC loop (delta_mpc, gravity compensation) ->
    dataset rows, on one GPU in one piece; the same rows come out of what rank 3 of 8 would run alone (one-warp kernels)
    and out of a 4096-curve slice (team kernels), bit for bit, and a re-run reproduces them."""
    n, frames, P, G = 1 << 20, 20, 32, 8
    rows, actual, xyz, st = _mpc_pipeline(tables_v, 0, n, frames, P)
    assert rows.shape == (n, frames, 13) and rows.dtype == torch.float64
    assert float((st & 1).double().mean()) > 0.99
    assert torch.isfinite(rows).all() and float(rows[:, :, :5].abs().max()) <= 0.5
    shard = n // G
    part = _mpc_pipeline(tables_v, 3 * shard, 4 * shard, frames, P)[0]
    assert torch.equal(rows[3 * shard:4 * shard], part)
    del part
    small = _mpc_pipeline(tables_v, 777_777, 777_777 + 4096, frames, P)[0]
    assert torch.equal(rows[777_777:777_777 + 4096], small)
    again = _mpc_pipeline(tables_v, 777_777, 777_777 + 4096, frames, P)[0]
    assert torch.equal(small, again)
    # the loop tracks: after the first frames the end effector is on the curve (y-z plane curves, the model's range)
    ee_err = (actual[:, 10:, :3] - xyz[:, 11:frames + 1]).norm(dim=2)
    print(f"2^20 curves x {frames} frames: ee error mean {float(ee_err.mean()) * 1e3:.2f} mm, p99 {float(ee_err.flatten()[::97].quantile(0.99)) * 1e3:.2f} mm")
    assert float(ee_err.mean()) < 0.01
