"""the reference's Koopman model next to the batched stepper - lift, model roll-out and MPC cost on the
device (`so101_koopman_score`), the MPC problem solved in closed form, and the closed loop on the CUDA simulator.
Weights: tests/golden/koopman_dkuc.npz = the reference's shipped checkpoint (results/SOARM101/11_27/DKUC/best_model.pt)."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "koopman_dkuc.npz")


def _weights():
    return {k: v.astype(np.float64) for k, v in np.load(GOLD).items()}


def test_closed_form_mpc_matches_normal_equations_cpu():
    """mpc_gains (what the product uses) == the oracle's normal-equation solve of the reference's NLP, and the solution
    is a stationary point of the reference's cost [REF MPC_Controler.py:65-98]."""
    from oracle import koopman_oracle as KO
    from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
    W = _weights()
    km = KoopmanModel(W, device="cpu")
    rng = np.random.default_rng(0)
    H = 10
    x = rng.uniform(-0.3, 0.3, 8); xref = rng.uniform(-0.3, 0.3, (H, 8))
    z0, zref = KO.lift(W, x), KO.lift(W, xref)
    np.testing.assert_allclose(km.lift(torch.as_tensor(x)).numpy(), z0, rtol=0, atol=1e-13)
    u_star = KO.mpc_solve(W, z0, zref, H)
    Kz, Kr = km.mpc_gains(H)
    u_gain = (Kz.numpy() @ z0 + Kr.numpy() @ zref.reshape(-1)).reshape(H, 5)
    np.testing.assert_allclose(u_gain, u_star, rtol=0, atol=1e-9)
    # stationarity: central differences of the oracle cost vanish at u*
    c0 = KO.score(W, z0, u_star[:, :, None], zref)[1][0]
    for _ in range(20):
        d = rng.standard_normal((H, 5)) * 1e-3
        cp = KO.score(W, z0, (u_star + d)[:, :, None], zref)[1][0]
        cm = KO.score(W, z0, (u_star - d)[:, :, None], zref)[1][0]
        assert cp > c0 and cm > c0 and abs(cp - cm) < 1e-6 * (cp - c0 + cm - c0) + 1e-12
    # batch control = first block of the gain solution, clipped
    u1 = km.mpc_control(torch.as_tensor(x)[None], torch.as_tensor(xref)[None], H).numpy()[0]
    np.testing.assert_allclose(u1, np.clip(u_star[0], -0.5, 0.5), atol=1e-9)


def test_delta_mpc_gains_match_the_restated_problem_cpu():
    """The reference's default formulation, 'delta_mpc' [REF args.py:75, control/MPC_Controler.py:100-141]: decision
    variable delta_u, u_t = u_prev + cumulative sum, cost on delta_u.  Closed-form gains == least-squares solve of the
    problem as the reference rolls it, for random (z0, zref, u_prev); and with u_prev = 0 and r -> the same weights the
    two formulations differ (so the test can tell them apart)."""
    from oracle import koopman_oracle as KO
    from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
    W = _weights()
    km = KoopmanModel(W, device="cpu")
    rng = np.random.default_rng(3)
    H = 10
    Kz, Kr, Ku = [K.numpy() for K in km.mpc_gains3(H, mpc_type="delta_mpc")]
    assert Ku.shape == (H * 5, 5) and np.abs(Ku).max() > 1e-3
    for _ in range(5):
        x = rng.uniform(-0.3, 0.3, 8); xref = rng.uniform(-0.3, 0.3, (H, 8)); up = rng.uniform(-0.6, 0.6, 5)
        z0, zref = KO.lift(W, x), KO.lift(W, xref)
        d_star = KO.mpc_solve(W, z0, zref, H, mpc_type="delta_mpc", u_prev=up)
        d_gain = (Kz @ z0 + Kr @ zref.reshape(-1) + Ku @ up).reshape(H, 5)
        np.testing.assert_allclose(d_gain, d_star, rtol=0, atol=1e-9)
        u_mpc = KO.mpc_solve(W, z0, zref, H, mpc_type="mpc")
        assert np.abs(np.cumsum(d_star, 0) + up - u_mpc).max() > 1e-3
    # 'mpc' through the same entry point: Ku = 0
    Kz2, Kr2, Ku2 = [K.numpy() for K in km.mpc_gains3(H, mpc_type="mpc")]
    assert not Ku2.any() and np.array_equal(Kz2, km.mpc_gains(H)[0].numpy())
    # the controller's u_prev hand-off [REF MPC_Controler.py:143-152, Koopman_MPC.py:217]: u_prev <- u0 (unclipped)
    ctl = KO.Controller(W, H, "delta_mpc")
    x = rng.uniform(-0.3, 0.3, 8); zref = KO.lift(W, rng.uniform(-0.3, 0.3, (H, 8)) * 3)
    u0, a = ctl.step(x, zref)
    assert np.array_equal(ctl.u_prev, u0) and np.array_equal(a, np.clip(u0, -0.5, 0.5))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
def test_koopman_score_matches_numpy(dtype):
    from oracle import koopman_oracle as KO
    from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
    W = _weights()
    km = KoopmanModel(W)
    rng = np.random.default_rng(1)
    H, n = 50, 8192 + 37
    U = rng.uniform(-0.5, 0.5, (H, 5, n))
    Ud = torch.as_tensor(U, dtype=dtype).cuda().contiguous()
    Uh = Ud.cpu().numpy().astype(np.float64)                      # what the kernel sees
    z0 = KO.lift(W, rng.uniform(-0.3, 0.3, 8)); zref = KO.lift(W, rng.uniform(-0.3, 0.3, (H, 8)))
    for ref in (zref, None):
        X, c = km.score(z0, Ud, ref)
        Xo, co = KO.score(W, z0, Uh, ref)
        assert X.shape == (n, H + 1, 8) and X.dtype == torch.float32 and c.dtype == torch.float64
        np.testing.assert_allclose(X.cpu().numpy(), Xo.astype(np.float32), rtol=0, atol=2e-7)
        np.testing.assert_allclose(c.cpu().numpy(), co, rtol=1e-12)
    # optimality through the GPU scorer: no perturbation of the closed-form solution costs less
    Kz, Kr = km.mpc_gains(10)
    u_star = (Kz.cpu().numpy() @ z0 + Kr.cpu().numpy() @ zref[:10].reshape(-1)).reshape(10, 5)
    cand = u_star[:, :, None] + np.concatenate([np.zeros((10, 5, 1)), rng.standard_normal((10, 5, 4095)) * 0.05], -1)
    _, cc = km.score(z0, torch.as_tensor(cand, dtype=dtype).cuda().contiguous(), zref[:10], want_pred=False)
    tol = 0.0 if dtype == torch.float64 else 1e-5 * float(cc[0])
    assert float(cc.min()) >= float(cc[0]) - tol and int(cc.argmin()) == 0 or dtype == torch.float32


@pytest.mark.gpu
def test_model_prediction_vs_physics_config5(tables_v):
    """BASELINE config 5 as the reference uses it: the same 8192 x 50 control sequences through the physics (`shoot`,
    gravity compensation held per control step as Koopman_MPC.py:119 does) and through the model (`score`)."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    km = KoopmanModel(_weights())
    B, H = 8192, 50
    env = SOARM101VecEnv(tables=tables_v, num_envs=B)
    g = torch.Generator(device="cuda").manual_seed(5)
    U = (torch.rand((H, 5, B), generator=g, dtype=torch.float64, device="cuda") - 0.5).contiguous()
    s0 = np.zeros(18); s0[:5] = [0.1, -0.2, 0.15, 0.05, -0.1]
    X = env.shoot(s0, U, flags=T_.ROLL_GRAVCOMP_HOLD)                        # physics  [B, H+1, 8]
    Xhat, cost = km.score(km.lift(X[0, 0].double()), U, None)                # model    [B, H+1, 8]
    err = (Xhat - X).abs()
    mae10, mae50 = float(err[:, 1:11].mean()), float(err[:, 1:].mean())
    print(f"model-vs-physics MAE over 8192 sequences: horizon 10: {mae10:.2e}, horizon 50: {mae50:.2e}")
    assert torch.equal(Xhat[:, 0], X[:, 0].expand(B, 8)) or float((Xhat[:, 0] - X[:, 0]).abs().max()) < 1e-7
    assert mae10 < 1.5e-3 and mae50 < 3.5e-3        # measured 7.3e-4 / 1.6e-3; published 200-step MAE: 6.84e-3
    np.testing.assert_allclose(cost.cpu().numpy(), 0.5 * (U ** 2).sum((0, 1)).cpu().numpy(), rtol=1e-12)


@pytest.mark.gpu
def test_closed_loop_mpc_tracks_on_the_cuda_simulator(tables_v):
    """The Koopman_MPC.py loop [REF Koopman_MPC.py:119,197-222] for 512 envs at once: lift, closed-form MPC, clip,
    env.step with gravity compensation.  The reference's model was trained on real MuJoCo data; that its MPC tracks on
    this simulator is an end-to-end check of the stepper against the reference's own artefact."""
    from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    km = KoopmanModel(_weights())
    N, H, steps = 512, 10, 250
    dt = 0.02
    g = torch.Generator(device="cuda").manual_seed(11)
    amp = (torch.rand((N, 5), generator=g, dtype=torch.float64, device="cuda") - 0.5) * 0.5      # +-0.25 rad
    freq = 0.1 + 0.2 * torch.rand((N, 5), generator=g, dtype=torch.float64, device="cuda")       # 0.1-0.3 Hz
    t_all = torch.arange(steps + H + 1, device="cuda", dtype=torch.float64) * dt
    qref = amp[:, None, :] * torch.sin(2 * np.pi * freq[:, None, :] * t_all[None, :, None])      # [N, steps+H+1, 5]
    # reference observations [ee, q] from the simulator's own forward kinematics
    scratch = SOARM101VecEnv(tables=tables_v, num_envs=N * (steps + H + 1))
    qfull = torch.zeros((N * (steps + H + 1), 6), dtype=torch.float64, device="cuda")
    qfull[:, :5] = qref.reshape(-1, 5)
    scratch.set_state(qfull, torch.zeros_like(qfull), torch.zeros_like(qfull))
    xref = scratch.forward()[0].double().reshape(N, steps + H + 1, 8).clone()
    env = SOARM101VecEnv(tables=tables_v, num_envs=N, gravity_compensation=True)
    q0 = torch.zeros((N, 6), dtype=torch.float64, device="cuda")
    env.set_state(q0, torch.zeros_like(q0), torch.zeros_like(q0))
    x = env.forward()[0].double().clone()
    errs = []
    for k in range(steps):
        u = km.mpc_control(x, xref[:, k + 1:k + 1 + H], H)
        x = env.step(u)[0].double()
        errs.append((x[:, 3:] - qref[:, k + 1]).abs().mean().item())
    tail = float(np.mean(errs[100:]))
    ee = float((x[:, :3] - xref[:, steps, :3]).norm(dim=-1).mean())
    print(f"closed-loop Koopman MPC on 512 envs: mean |q - qref| after settling {tail:.3e} rad, final ee error {ee*1e3:.2f} mm")
    assert tail < 0.01 and ee < 0.006        # measured: 4.2e-3 rad, 2.0 mm


@pytest.mark.gpu
def test_shooting_mpc_improves_on_the_model_solution(tables_v):
    """Sampling MPC through the physics: 8192 candidates around the closed-form (model) solution in one k_shoot
    launch, costed with the reference's MPC cost on the simulated outcome.  Candidate 0 IS the model's solution, so
    the winner is never worse; re-shooting the winner alone reproduces its trajectory and cost (determinism,
    batch-position invariance); in closed loop the sampled controls track at least as well as the model's."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    km = KoopmanModel(_weights())
    B, H = 8192, 10
    shooter = SOARM101VecEnv(tables=tables_v, num_envs=B)
    plant = SOARM101VecEnv(tables=tables_v, num_envs=2, gravity_compensation=True)     # env 0: sampled, env 1: model
    q0 = torch.zeros((2, 6), dtype=torch.float64, device="cuda")
    q0[:, :5] = torch.tensor([0.1, -0.2, 0.15, 0.05, -0.1], dtype=torch.float64)
    plant.set_state(q0, torch.zeros_like(q0), torch.zeros_like(q0))
    x = plant.forward()[0].double().clone()                                          # [2, 8]
    # reference: joints swing sinusoidally; reference observations from the simulator's own kinematics
    steps = 40
    t = torch.arange(steps + H + 1, dtype=torch.float64, device="cuda") * 0.02
    qref = q0[0, :5][None] + 0.2 * torch.sin(2 * np.pi * 0.4 * t)[:, None] * torch.tensor([1.0, -0.7, 0.5, 0.8, -1.0],
                                                                                         dtype=torch.float64, device="cuda")
    scratch = SOARM101VecEnv(tables=tables_v, num_envs=steps + H + 1)
    qf = torch.zeros((steps + H + 1, 6), dtype=torch.float64, device="cuda"); qf[:, :5] = qref
    scratch.set_state(qf, torch.zeros_like(qf), torch.zeros_like(qf))
    xref = scratch.forward()[0].double().clone()                                     # [steps+H+1, 8]
    err_s, err_m, gains = [], [], []
    for k in range(steps):
        qp, qv, qw = plant.get_state()
        s0 = torch.cat([qp[0], qv[0], qw[0]]).cpu().numpy()
        Ub, costs, best = km.shooting_mpc(shooter, s0, x[0], xref[k + 1:k + 1 + H], sigma=0.05,
                                          flags=T_.ROLL_GRAVCOMP_HOLD, seed=k)
        assert float(costs[best]) <= float(costs[0])
        gains.append(1.0 - float(costs[best]) / float(costs[0]))
        if k == 0:
            # the winner re-shot in another batch position: same observations, same cost
            U1 = Ub[:, :, None].expand(H, 5, B).contiguous()
            X1 = shooter.shoot(s0, U1, flags=T_.ROLL_GRAVCOMP_HOLD)
            assert torch.equal(X1[0], X1[B - 1])
            Z = km.lift(X1[0, 1:].double())
            c1 = 50.0 * ((Z - km.lift(xref[1:1 + H])) ** 2).sum() + 0.5 * (Ub ** 2).sum()
            assert abs(float(c1) - float(costs[best])) <= 1e-9 * float(c1)
        u_model = km.mpc_control(x[1:2], xref[None, k + 1:k + 1 + H], H)[0]
        x = plant.step(torch.stack([Ub[0], u_model]))[0].double().clone()
        err_s.append(float((x[0, 3:] - qref[k + 1]).abs().mean()))
        err_m.append(float((x[1, 3:] - qref[k + 1]).abs().mean()))
    print(f"shooting MPC (8192 candidates x {H} steps through the physics): mean cost reduction vs the model's solution "
          f"{100 * np.mean(gains):.1f} %; closed-loop |q - qref| {np.mean(err_s[10:]):.2e} rad (sampled) vs "
          f"{np.mean(err_m[10:]):.2e} rad (model)")
    assert np.mean(err_s[10:]) < 1.25 * np.mean(err_m[10:]) + 1e-3


@pytest.mark.parametrize("u_z", [False, True])
def test_bilinear_variant_mpc_matches_the_restated_problem_cpu(u_z):
    """The bilinear model variants (DBKN / IBKN: z+ = A z + B u + H (z (x) u)) under the reference's MPC: `linearize_B`
    freezes B_total(z0) over the horizon [REF control/MPC_Controler.py:46-63, models/KoopmanBase.py:62-110], gains then
    depend on z0 (one solve per env).  The reference ships no bilinear checkpoint: the shipped linear model plus a random H
    layer.  Batched solve == the problem restated and solved by least squares, both formulations, both Kronecker orders."""
    from oracle import koopman_oracle as KO
    from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
    W = _weights()
    rng = np.random.default_rng(12)
    nz, nu = W["lB.weight"].shape
    W["H.weight"] = rng.standard_normal((nz, nz * nu)) * 0.05
    km = KoopmanModel({k: v for k, v in W.items() if k != "H.weight"}, device="cpu")
    km.set_bilinear(W["H.weight"], u_z=u_z)
    H, N = 10, 6
    x = rng.uniform(-0.3, 0.3, (N, 8)); xref = rng.uniform(-0.3, 0.3, (N, H, 8)); up = rng.uniform(-0.4, 0.4, (N, 5))
    for mpc_type in ("delta_mpc", "mpc"):
        u0, a = km.mpc_control_bilinear(torch.as_tensor(x), torch.as_tensor(xref), torch.as_tensor(up), H, mpc_type)
        for e in range(N):
            z0, zref = KO.lift(W, x[e]), KO.lift(W, xref[e])
            sol = KO.mpc_solve(W, z0, zref, H, mpc_type=mpc_type, u_prev=up[e] if mpc_type == "delta_mpc" else None, u_z=u_z)
            np.testing.assert_allclose(u0[e].numpy(), sol[0] + up[e], rtol=0, atol=1e-8)
        assert np.array_equal(a.numpy(), np.clip(u0.numpy(), -0.5, 0.5))
    # and it differs from the linear model's answer (the H layer matters)
    lin = KO.mpc_solve({k: v for k, v in W.items() if k != "H.weight"}, KO.lift(W, x[0]), KO.lift(W, xref[0]), H, mpc_type="mpc")
    assert np.abs(lin[0] + up[0] - km.mpc_control_bilinear(torch.as_tensor(x[:1]), torch.as_tensor(xref[:1]),
                                                            torch.as_tensor(up[:1]), H, "mpc")[0][0].numpy()).max() > 1e-4


@pytest.mark.gpu
def test_lift_and_mpc_step_kernels_match_numpy():
    """The fused encoder kernel (`so101_koopman_lift`), the reference fold (`so101_koopman_feedforward`) and one MPC frame
    (`so101_koopman_mpc_step`) against the numpy restatement, for both formulations, both observation layouts / dtypes and
    ragged batch sizes; float64 arithmetic on both sides: 1e-12."""
    from oracle import koopman_oracle as KO
    from lerobot_mujoco_sim2real_b200.koopman import KoopmanModel
    W = _weights()
    km = KoopmanModel(W)
    rng = np.random.default_rng(5)
    H, P = 10, 37
    for n in (1, 63, 64, 1000):
        x = rng.uniform(-0.6, 0.6, (n, 8))
        z_ref = KO.lift(W, x)
        xt = torch.as_tensor(x).cuda()
        np.testing.assert_allclose(km.lift_device(xt).cpu().numpy(), z_ref, rtol=0, atol=1e-12)
        np.testing.assert_allclose(km.lift_device(xt.t().contiguous(), soa=True).cpu().numpy(), z_ref, rtol=0, atol=1e-12)
        x32 = xt.float()
        np.testing.assert_allclose(km.lift_device(x32.t().contiguous(), soa=True).cpu().numpy(),
                                   KO.lift(W, x32.double().cpu().numpy()), rtol=0, atol=1e-12)
        np.testing.assert_allclose(km.lift_device(xt).cpu().numpy(), km.lift(xt).cpu().numpy(), rtol=0, atol=1e-12)
    n = 130
    xref = rng.uniform(-0.4, 0.4, (n, P, 8))
    x = rng.uniform(-0.4, 0.4, (n, 8)); up = rng.uniform(-0.7, 0.7, (n, 5))
    for mpc_type in ("delta_mpc", "mpc"):
        Kz, Kr, Ku = [K.cpu().numpy() for K in km.mpc_gains3(H, mpc_type=mpc_type)]
        uff = km.feedforward(torch.as_tensor(xref).cuda(), H, mpc_type).cpu().numpy()
        for k in (0, 5, P - 4, P - 1):
            zr = np.zeros((n, H, km.nz))
            m = min(H, P - 1 - k)
            if m > 0:
                zr[:, :m] = KO.lift(W, xref[:, k + 1:k + 1 + m])
            np.testing.assert_allclose(uff[:, k], zr.reshape(n, -1) @ Kr[:5].T, rtol=0, atol=1e-11)
        k = 5
        u_prev = torch.as_tensor(up.T.copy()).cuda()
        ctrl = torch.zeros((5, n), dtype=torch.float64, device="cuda")
        a_out = torch.zeros((n, 5), dtype=torch.float64, device="cuda")
        km.mpc_step(torch.as_tensor(x).cuda(), False, torch.as_tensor(uff).cuda(), k, u_prev, ctrl, a_out, H, mpc_type)
        u0_ref = KO.lift(W, x) @ Kz[:5].T + uff[:, k] + up @ Ku[:5].T + up
        np.testing.assert_allclose(u_prev.cpu().numpy().T, u0_ref, rtol=0, atol=1e-11)
        np.testing.assert_allclose(a_out.cpu().numpy(), np.clip(u0_ref, -0.5, 0.5), rtol=0, atol=1e-11)
        assert np.array_equal(ctrl.cpu().numpy().T, a_out.cpu().numpy())
        # one whole controller step against the restated NLP (least squares) for a few envs
        for e in (0, 77):
            zr = np.zeros((H, km.nz)); zr[:] = KO.lift(W, xref[e, k + 1:k + 1 + H])
            ctl = KO.Controller(W, H, mpc_type); ctl.u_prev = up[e].copy()
            u0_o, a_o = ctl.step(x[e], zr)
            np.testing.assert_allclose(u_prev.cpu().numpy()[:, e], u0_o, rtol=0, atol=1e-9)
