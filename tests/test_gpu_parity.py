"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Protocol (because scene A is chaotic under kv=50, F3):
  P1  teacher-forced: every one of 1000 physics steps of an oracle trajectory is replayed on
      the GPU from the oracle's own (qpos, qvel, qacc_warmstart, ctrl); next state must agree
      to 1e-12 relative (fp64).
  P2  free-running on scene A: error-vs-step against the oracle's own 1-ulp self-divergence.
  P3  free-running on scene B (position servos, contractive): 1e-9 relative after 1000 steps.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    """relative error with the denominator floored at 1e-3 (absolute for smaller values)"""
    return np.abs(a - b) / (1e-3 + np.maximum(np.abs(a), np.abs(b)))


def _rel_true(a, b, floor=1e-9):
    """true relative error |a - b| / max(|a|, |b|) over the entries whose magnitude exceeds `floor` (an absolute error of
    one ulp of 1.0 on a value of 1e-9 is a relative error of 1e-7: below the floor only the absolute error means anything)"""
    m = np.maximum(np.abs(a), np.abs(b))
    ok = m > floor
    return (np.abs(a - b)[ok] / m[ok]) if ok.any() else np.zeros(1)


def _oracle_traj(O, tables, n, steps, seed, ctrl_hold=10, qscale=0.3, uscale=0.5):
    rng = np.random.default_rng(seed)
    state = np.zeros((n, 18))
    state[:, :5] = rng.uniform(-qscale, qscale, (n, 5))
    states, ctrls = [], []
    ctrl = np.zeros((n, 6))
    for t in range(steps):
        if t % ctrl_hold == 0:
            ctrl = np.zeros((n, 6))
            ctrl[:, :5] = rng.uniform(-uscale, uscale, (n, 5))
        states.append(state)
        ctrls.append(ctrl)
        state, _, _ = O.step_batch(tables, state, ctrl, 1)
    states.append(state)
    return np.stack(states), np.stack(ctrls)   # [steps+1, n, 18], [steps, n, 6]


def _gpu_step(tables, state, ctrl, nsub=1, dtype="float64", family=0):
    from lerobot_mujoco_sim2real_b200 import tables as T_
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    n = state.shape[0]
    env = SOARM101VecEnv(tables=tables, num_envs=n, dtype=dtype)
    env.set_option(T_.OPT_KERNEL_FAMILY, family)
    env.set_state(state[:, :6], state[:, 6:12], state[:, 12:18])
    u = torch.as_tensor(ctrl.T.copy(), dtype=env.torch_dtype, device=env.device).contiguous()
    obs = env.step_soa(u, nsub).t().cpu().numpy()
    q, v, w = env.get_state()
    out = np.concatenate([q.cpu().numpy(), v.cpu().numpy(), w.cpu().numpy()], axis=1).astype(np.float64)
    return out, obs, env


@pytest.mark.parametrize("scene", ["v", "p"])
def test_p1_teacher_forced_fp64(oracle_mod, tables_v, tables_p, scene):
    O = oracle_mod
    tables = tables_v if scene == "v" else tables_p
    n, steps = 128, 1000
    uscale = 0.5 if scene == "v" else 1.5
    states, ctrls = _oracle_traj(O, tables, n, steps, seed=7, uscale=uscale)
    s_in = states[:-1].reshape(-1, 18)
    s_ref = states[1:].reshape(-1, 18)
    out, _, env = _gpu_step(tables, s_in, ctrls.reshape(-1, 6), 1)
    err = _rel(out, s_ref)
    st = env.stats()
    print(f"scene {scene}: max rel err qpos {err[:, :6].max():.3e} qvel {err[:, 6:12].max():.3e} "
          f"qacc {err[:, 12:].max():.3e}; newton/step {st['newton_iters'] / st['physics_steps']:.3f}")
    print("   qvel rel err quantiles 50/99/99.9/max:",
          " ".join(f"{np.quantile(err[:, 6:12], x):.2e}" for x in (0.5, 0.99, 0.999, 1.0)))
    print("   qacc rel err quantiles 50/99/99.9/max:",
          " ".join(f"{np.quantile(err[:, 12:], x):.2e}" for x in (0.5, 0.99, 0.999, 1.0)))
    assert err[:, :6].max() < 1e-12
    # the same without the floor: qpos entries above 1e-6 rad in magnitude agree to 1e-9 relative, all to 1e-15 absolute
    tr = _rel_true(out[:, :6], s_ref[:, :6], 1e-6)
    print(f"   qpos true relative error (|q| > 1e-6): max {tr.max():.2e}; absolute: max {np.abs(out[:, :6] - s_ref[:, :6]).max():.2e}")
    assert tr.max() < 1e-9 and np.abs(out[:, :6] - s_ref[:, :6]).max() < 1e-14
    # qacc is the Newton solver's output: both implementations stop at opt.tolerance = 1e-8, so it
    # agrees to rounding when the last step lands on the exact minimiser of the piecewise-quadratic
    # cost (the bulk) and to ~1e-10 otherwise; qvel inherits h * that.
    assert np.quantile(err[:, 6:12], 0.99) < 1e-12
    assert err[:, 6:12].max() < 1e-10
    assert np.quantile(err[:, 12:], 0.99) < 1e-10
    assert err[:, 12:].max() < 1e-7


def test_p1_teacher_forced_fp32(oracle_mod, tables_v):
    O = oracle_mod
    n, steps = 64, 500
    states, ctrls = _oracle_traj(O, tables_v, n, steps, seed=11)
    out, _, _ = _gpu_step(tables_v, states[:-1].reshape(-1, 18), ctrls.reshape(-1, 6), 1, dtype="float32")
    ref = states[1:].reshape(-1, 18)
    dq = np.abs(out[:, :6] - ref[:, :6]).max()
    dv = np.abs(out[:, 6:12] - ref[:, 6:12])
    print(f"fp32 one-step: |dq| max {dq:.3e}, |dqvel| max {dv.max():.3e} median {np.median(dv):.3e}")
    # stated fp32 tolerance for ONE teacher-forced physics step
    assert dq < 2e-6
    assert np.quantile(dv, 0.99) < 2e-3


def test_obs_ee_lag_and_rounding(oracle_mod, tables_v):
    """step() observation = site of the last sub-step's forward (lags qpos by one sub-step), float32."""
    O = oracle_mod
    n = 256
    states, ctrls = _oracle_traj(O, tables_v, n, 50, seed=3)
    s_in, c_in = states[-1], ctrls[-1]
    ref_state, ref_obs, _ = O.step_batch(tables_v, s_in, c_in, 10)
    out, obs, _ = _gpu_step(tables_v, s_in, c_in, 10)
    assert obs.dtype == np.float32
    assert np.abs(obs - ref_obs.astype(np.float32)).max() <= 2e-7
    # the lag: the ee position is NOT the site at the returned qpos
    from lerobot_mujoco_sim2real_b200 import mjcf
    lagfree = np.stack([mjcf.site_numpy(tables_v, ref_state[i, :6]) for i in range(8)])
    assert np.abs(lagfree - ref_obs[:8, :3]).max() > 1e-7


def test_p3_free_running_scene_b(oracle_mod, tables_p):
    """BASELINE.json bar taken literally on the contractive position-servo scene."""
    O = oracle_mod
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    n, T_ctrl = 512, 100
    rng = np.random.default_rng(5)
    q0 = np.zeros((n, 6)); q0[:, :5] = rng.uniform(-0.3, 0.3, (n, 5))
    # random-walk position targets around the initial pose
    U = np.cumsum(rng.uniform(-0.05, 0.05, (T_ctrl + 1, 5, n)), axis=0) + q0[:, :5].T[None]
    spec = O.make_spec(kind=3, u=np.ascontiguousarray(U))
    _, fin, _ = O.rollout(tables_p, spec, n, T_ctrl, 10, qpos0=q0, want_rows=False)
    env = SOARM101VecEnv(tables=tables_p, num_envs=n, dtype="float64")
    env.set_state(q0, np.zeros((n, 6)), np.zeros((n, 6)))
    Ud = torch.as_tensor(U, dtype=torch.float64, device=env.device).contiguous()
    from lerobot_mujoco_sim2real_b200 import tables as T
    env.rollout(T_ctrl, "tensor", u=Ud, flags=T.ROLL_NO_RESET)
    q, v, _ = env.get_state()
    eq = _rel(q.cpu().numpy(), fin[:, :6]).max()
    ev = np.abs(v.cpu().numpy() - fin[:, 6:12]).max() / max(1e-3, np.abs(fin[:, 6:12]).max())
    print(f"scene B free-running 1000 physics steps: rel err qpos {eq:.3e} qvel {ev:.3e}")
    assert eq < 1e-9 and ev < 1e-9


def test_p2_free_running_scene_a_curve(oracle_mod, tables_v):
    """Chaotic scene: GPU-vs-oracle divergence must track the oracle's own 1-ulp self-divergence."""
    O = oracle_mod
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    from lerobot_mujoco_sim2real_b200 import tables as T
    n, T_ctrl = 256, 100
    rng = np.random.default_rng(9)
    q0 = np.zeros((n, 6)); q0[:, :5] = rng.uniform(-0.3, 0.3, (n, 5))
    q0p = q0.copy(); q0p[:, :5] = np.nextafter(q0[:, :5], 1.0)   # 1-ulp perturbation
    U = rng.uniform(-0.5, 0.5, (T_ctrl + 1, 5, n))
    spec = O.make_spec(kind=3, u=np.ascontiguousarray(U))
    env = SOARM101VecEnv(tables=tables_v, num_envs=n, dtype="float64")
    Ud = torch.as_tensor(U, dtype=torch.float64, device=env.device).contiguous()
    curve = []
    for Tc in (1, 3, 5, 7, 10, 30, 100):
        _, fin, _ = O.rollout(tables_v, spec, n, Tc, 10, qpos0=q0, want_rows=False)
        _, finp, _ = O.rollout(tables_v, spec, n, Tc, 10, qpos0=q0p, want_rows=False)
        env.set_state(q0, np.zeros((n, 6)), np.zeros((n, 6)))
        env.rollout(Tc, "tensor", u=Ud[: Tc + 1].contiguous(), flags=T.ROLL_NO_RESET)
        q, _, _ = env.get_state()
        self_div = np.median(np.abs(fin[:, :6] - finp[:, :6]).max(axis=1))
        gpu_div = np.median(np.abs(q.cpu().numpy() - fin[:, :6]).max(axis=1))
        curve.append((Tc * 10, self_div, gpu_div))
    for s, a, b in curve:
        print(f"physics step {s:5d}: oracle 1-ulp self-divergence {a:.3e}   gpu-vs-oracle {b:.3e}")
    # before chaos amplifies rounding: tight agreement
    assert curve[0][2] < 1e-12
    # the same curve within 10x of what a 1-ulp perturbation does to the oracle itself, and <= 1e-9 up to the
    # step where that self-divergence crosses 1e-10
    for s, a, b in curve:
        assert b <= max(1e-12, 10 * a), (s, a, b)
        if a < 1e-10:
            assert b <= 1e-9, (s, a, b)


def test_free_running_against_joint_limits_scene_b(contact_free, tables_p):
    """Position targets far beyond the joint ranges (ctrl clamp, force clamp, limit rows active for
    hundreds of steps): free-running agreement with the oracle on the contractive scene (contact-free pipeline on both
    sides; resting on the table under load is tests/test_contact.py's free-running case)."""
    O = contact_free
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    from lerobot_mujoco_sim2real_b200 import tables as T
    n, T_ctrl = 256, 60
    rng = np.random.default_rng(12)
    q0 = np.zeros((n, 6)); q0[:, :5] = rng.uniform(-0.3, 0.3, (n, 5))
    U = np.repeat(rng.choice([-3.0, 3.0], (1, 5, n)), T_ctrl + 1, axis=0)     # clamped to the ctrlrange = joint range
    U += rng.uniform(-0.01, 0.01, U.shape)
    spec = O.make_spec(kind=3, u=np.ascontiguousarray(U))
    _, fin, _ = O.rollout(tables_p, spec, n, T_ctrl, 10, qpos0=q0, want_rows=False)
    env = SOARM101VecEnv(tables=tables_p, num_envs=n, dtype="float64", hulls=None)
    env.set_state(q0, np.zeros((n, 6)), np.zeros((n, 6)))
    env.rollout(T_ctrl, "tensor", u=torch.as_tensor(U).cuda().contiguous(), flags=T.ROLL_NO_RESET)
    q, v, _ = env.get_state()
    q, v = q.cpu().numpy(), v.cpu().numpy()
    lim = np.array([[tables_p.jnt_range[k][0], tables_p.jnt_range[k][1]] for k in range(6)])
    at_limit = (np.abs(q[:, :5] - lim[:5, 0]) < 0.02) | (np.abs(q[:, :5] - lim[:5, 1]) < 0.02)
    frac_lim = at_limit.mean()
    st = env.stats()
    eq = _rel(q, fin[:, :6]).max()
    ev = np.abs(v - fin[:, 6:12]).max()
    print(f"joints resting at a limit: {frac_lim:.2f}; steps with limit rows {st['limit_steps'] / st['physics_steps']:.2f}; "
          f"rel err qpos {eq:.2e}, abs err qvel {ev:.2e}")
    assert frac_lim > 0.5 and st["limit_steps"] > 0.3 * st["physics_steps"]
    assert int((env.flags() & T.FLAG_LIMIT).ne(0).sum()) > 0.9 * n
    assert eq < 1e-9 and ev < 1e-8


@pytest.mark.parametrize("variant", ["frictionless", "undamped", "mixed", "heavy"])
@pytest.mark.parametrize("split", ["0", "1"])
def test_model_variants_teacher_forced(oracle_mod, tables_v, monkeypatch, variant, split):
    """Paths the two reference scenes never take: no friction rows at all (unconstrained unless a limit is active), no
    joint damping (explicit Euler), friction / damping on some dofs only, other inertial constants - one teacher-forced
    step from 4000 random states against the oracle, in both kernel families."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    O = oracle_mod
    t = T_.tables_from_dict(T_.tables_to_dict(tables_v))
    if variant == "frictionless":
        for k in range(6): t.dof_frictionloss[k] = 0.0
    elif variant == "undamped":
        for k in range(6): t.dof_damping[k] = 0.0
    elif variant == "mixed":
        for k in (0, 2, 4): t.dof_frictionloss[k] = 0.0
        t.dof_damping[1] = 0.0; t.dof_damping[3] = 2.0
    else:
        for k in range(6): t.dof_armature[k] = 0.005 + 0.01 * k
        for b in range(2, 8): t.body_mass[b] *= 1.7
        t.gravity[0], t.gravity[2] = 1.0, -5.0
    rng = np.random.default_rng(3)
    n = 4000
    state = np.zeros((n, 18))
    state[:, :5] = rng.uniform(-0.6, 0.6, (n, 5)); state[:, 5] = rng.uniform(-0.1, 1.0, n)
    state[:, 6:12] = rng.uniform(-1.0, 1.0, (n, 6))
    ctrl = np.zeros((n, 6)); ctrl[:, :5] = rng.uniform(-0.5, 0.5, (n, 5))
    state, _, _ = O.step_batch(t, state, ctrl, 3)          # three oracle steps: realistic qacc_warmstart
    ref, _, _ = O.step_batch(t, state, ctrl, 1)
    out, _, env = _gpu_step(t, state, ctrl, 1, family=T_.FAMILY_TEAM if split == "1" else T_.FAMILY_ONEWARP)
    err = _rel(out, ref)
    print(f"{variant} split={split}: qpos {err[:, :6].max():.2e} qvel {err[:, 6:12].max():.2e} qacc {err[:, 12:].max():.2e}")
    assert err[:, :6].max() < 1e-12 and err[:, 6:12].max() < 1e-9 and np.quantile(err[:, 12:], 0.99) < 1e-9


@pytest.mark.gpu
def test_mechanical_energy_is_dissipated(tables_p):
    """A property that needs no oracle.  Position-servo scene, constant targets: every force on the arm is conservative (gravity,
    the servos' kp (q* - q) = a spring) or dissipative (kv, joint damping, friction loss), so E = 1/2 qd' M(q) qd + V_gravity(q)
    + 1/2 kp |q - q*|^2 can only fall.  M and V come from the textbook formulas in mjcf.py (world-frame Jacobians), not from the
    kernels' link-local recursion, so this checks mass matrix, bias forces, actuation and integrator of the CUDA path against
    mechanics itself."""
    from lerobot_mujoco_sim2real_b200 import mjcf
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    t = tables_p
    n, steps = 24, 80
    rng = np.random.default_rng(3)
    q0 = rng.uniform(-0.3, 0.3, (n, 6)); q0[:, 5] = 0.2
    v0 = rng.uniform(-1.0, 1.0, (n, 6))
    target = q0 + rng.uniform(-0.1, 0.1, (n, 6))
    kp = np.array([t.act_gain[i] for i in range(6)])
    grav = np.array(t.gravity[:])

    def energy(q, v, qs):
        xpos, xmat, _, _ = mjcf.fk_numpy(t, q)
        V = -sum(t.body_mass[b] * grav @ (xpos[b] + xmat[b] @ np.array(t.body_ipos[b][:])) for b in range(1, t.nbody))
        return 0.5 * v @ mjcf.mass_matrix_numpy(t, q) @ v + V + 0.5 * np.sum(kp * (q - qs) ** 2)

    env = SOARM101VecEnv(tables=t, num_envs=n, hulls=None)
    env.set_state(q0, v0, np.zeros((n, 6)))
    u = torch.as_tensor(target[:, :5].T.copy(), dtype=torch.float64, device="cuda")
    # the gripper's servo target is 0 in step() (ctrl[5] stays 0): its spring is anchored there
    target[:, 5] = 0.0
    E = np.zeros((steps + 1, n))
    E[0] = [energy(q0[i], v0[i], target[i]) for i in range(n)]
    for s in range(steps):
        env.step_soa(u, 1)
        q, v, _ = env.get_state()
        q, v = q.cpu().numpy(), v.cpu().numpy()
        E[s + 1] = [energy(q[i], v[i], target[i]) for i in range(n)]
    drop = E[0] - E[-1]
    rise = np.diff(E, axis=0).max(axis=0)
    print(f"energy: start {E[0].mean():.4f} J, after {steps} steps {E[-1].mean():.4f} J; largest single-step rise / total drop: "
          f"{(rise / drop).max():.2e}")
    assert (drop > 0.01).all()                       # most of the kinetic energy is gone
    assert (rise <= 1e-3 * drop).all()               # and it never comes back (semi-implicit Euler: O(h^2) wiggles at most)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 2e-5)])
def test_gravity_compensation_is_a_static_equilibrium(tables_v, dtype, tol):
    """A property that needs no oracle: qfrc_applied = qfrc_bias [REF Koopman_MPC.py:119] at rest cancels gravity exactly, the
    velocity servos command zero velocity, so the arm does not move - for any pose, over 500 physics steps.  Without the
    compensation the same arms sag (the servos are dampers, not position holds)."""
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    n = 512
    rng = np.random.default_rng(12)
    q0 = rng.uniform(-0.6, 0.6, (n, 6)); q0[:, 5] = rng.uniform(0.0, 0.5, n)
    u = torch.zeros((5, n), dtype=getattr(torch, dtype), device="cuda")
    moved = {}
    for comp in (True, False):
        env = SOARM101VecEnv(tables=tables_v, num_envs=n, dtype=dtype, hulls=None, gravity_compensation=comp)
        env.set_state(q0, np.zeros((n, 6)), np.zeros((n, 6)))
        for _ in range(50):
            env.step_soa(u, 10)
        q, v, _ = env.get_state()
        moved[comp] = (np.abs(q.cpu().numpy().astype(np.float64) - q0).max(), float(v.abs().max()))
    print(f"{dtype}: with compensation moved {moved[True][0]:.2e} rad (max |qvel| {moved[True][1]:.2e}); without {moved[False][0]:.2e} rad")
    assert moved[True][0] < tol and moved[True][1] < 100 * tol
    assert moved[False][0] > 1e-3


@pytest.mark.gpu
def test_bias_forces_against_the_lagrangian(tables_v):
    """A property that needs no oracle: mj_forward's qfrc_bias on the CUDA path (link-local RNEA) equals the bias force of the
    Lagrangian built from the textbook M(q) and V(q) of mjcf.py (world-frame Jacobians), by central differences:
    c_i = dV/dq_i + sum_jk (dM_ij/dq_k - 1/2 dM_jk/dq_i) qd_j qd_k."""
    from lerobot_mujoco_sim2real_b200 import mjcf
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    t = tables_v
    n, h = 48, 1e-5
    rng = np.random.default_rng(21)
    q = rng.uniform(-1.0, 1.0, (n, 6)); q[:, 5] = rng.uniform(0.0, 1.0, n)
    v = rng.uniform(-2.0, 2.0, (n, 6))
    grav = np.array(t.gravity[:])

    def V(x):
        xpos, xmat, _, _ = mjcf.fk_numpy(t, x)
        return -sum(t.body_mass[b] * grav @ (xpos[b] + xmat[b] @ np.array(t.body_ipos[b][:])) for b in range(1, t.nbody))

    env = SOARM101VecEnv(tables=t, num_envs=n, hulls=None)
    env.set_state(q, v, np.zeros((n, 6)))
    _, bias = env.forward()
    bias = bias.cpu().numpy()
    worst = 0.0
    for e in range(n):
        dM = np.zeros((6, 6, 6)); dV = np.zeros(6)
        for k in range(6):
            d = np.zeros(6); d[k] = h
            dM[k] = (mjcf.mass_matrix_numpy(t, q[e] + d) - mjcf.mass_matrix_numpy(t, q[e] - d)) / (2 * h)
            dV[k] = (V(q[e] + d) - V(q[e] - d)) / (2 * h)
        c = dV + np.einsum("kij,j,k->i", dM, v[e], v[e]) - 0.5 * np.einsum("ijk,j,k->i", dM, v[e], v[e])
        worst = max(worst, np.abs(bias[e] - c).max() / (1e-3 + np.abs(c).max()))
    print(f"qfrc_bias vs Lagrangian by central differences: worst relative difference {worst:.2e}")
    assert worst < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("damped", [False, True])
@pytest.mark.parametrize("family", ["onewarp", "team"])
def test_one_step_obeys_newtons_law(tables_v, family, damped):
    """A property that needs no oracle: without friction loss (a model variant; no constraint row is active away from the
    limits) one physics step from rest gives qvel' = h a with (M(q) + h B) a = gain * u - g(q), B = the joint damping that
    mj_Euler treats implicitly (or none): the textbook mass matrix and gravity vector of mjcf.py against the CUDA path's CRBA,
    factorisation, actuation and Euler step."""
    from lerobot_mujoco_sim2real_b200 import mjcf, tables as T_
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    t = T_.tables_from_dict(T_.tables_to_dict(tables_v))
    for k in range(6):
        t.dof_frictionloss[k] = 0.0
        if not damped:
            t.dof_damping[k] = 0.0
    B = np.diag([t.dof_damping[k] for k in range(6)])      # damped: mj_Euler's implicit step solves (M + h B) a = force
    n = 64
    rng = np.random.default_rng(31)
    q = rng.uniform(-1.0, 1.0, (n, 6)); q[:, 5] = rng.uniform(0.0, 1.0, n)
    u = np.zeros((n, 6)); u[:, :5] = rng.uniform(-0.05, 0.05, (n, 5))      # small: the force clamp (+-3.5) stays inactive
    env = SOARM101VecEnv(tables=t, num_envs=n, hulls=None)
    env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_ONEWARP if family == "onewarp" else T_.FAMILY_TEAM)
    env.set_state(q, np.zeros((n, 6)), np.zeros((n, 6)))
    env.step_soa(torch.as_tensor(u[:, :5].T.copy(), dtype=torch.float64, device="cuda"), 1)
    q1, v1, _ = [x.cpu().numpy() for x in env.get_state()]
    h = t.timestep
    worst = 0.0
    for e in range(n):
        M = mjcf.mass_matrix_numpy(t, q[e])
        tau = np.array([t.act_gain[i] * u[e, i] for i in range(6)]) - mjcf.gravity_bias_numpy(t, q[e])   # qvel = 0: no velocity feedback, no Coriolis
        a = np.linalg.solve(M + t.timestep * B, tau)
        worst = max(worst, np.abs(v1[e] / h - a).max() / (1e-3 + np.abs(a).max()))
        assert np.abs(q1[e] - (q[e] + h * v1[e])).max() < 1e-15        # semi-implicit Euler: the new velocity moves the position
    print(f"{family}: one step from rest, qvel'/h against M^-1 (tau - g): worst relative difference {worst:.2e}")
    assert worst < 1e-10
