"""CPU oracle (oracle/so101_oracle.c): independent re-derivations, invariants, pins, fingerprint."""
import os

import numpy as np
import pytest

from lerobot_mujoco_sim2real_b200 import mjcf

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_philox_known_answers(oracle_mod):
    """Random123 kat_vectors for philox4x32-10."""
    O = oracle_mod
    assert list(O.philox4x32_10([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(O.philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(O.philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = O.uniform8(42, 7, 3, 1)
    assert np.all((u >= 0) & (u < 1)) and len(set(u)) == 8
    assert not np.array_equal(u, O.uniform8(42, 8, 3, 1))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_smooth_dynamics_against_textbook_forms(oracle_mod, tables_v, seed):
    """M from Jacobians, gravity torque from the potential, Coriolis from dM/dq (all numpy)."""
    O, t = oracle_mod, tables_v
    rng = np.random.default_rng(seed)
    q, qd = rng.uniform(-1.5, 1.5, 6), rng.uniform(-3, 3, 6)
    o = O.Oracle(t)
    o.reset(); o.set("qpos", q); o.set("qvel", qd); o.forward()
    M = o.full_M()
    np.testing.assert_allclose(M, mjcf.mass_matrix_numpy(t, q), atol=1e-15)
    assert np.all(np.linalg.eigvalsh(M) > 0.02)   # SPD, armature-dominated
    np.testing.assert_allclose(o.arr("site_xpos"), mjcf.site_numpy(t, q), atol=1e-15)
    eps = 1e-6
    dM = np.zeros((6, 6, 6))
    for k in range(6):
        d = np.zeros(6); d[k] = eps
        dM[k] = (mjcf.mass_matrix_numpy(t, q + d) - mjcf.mass_matrix_numpy(t, q - d)) / (2 * eps)
    cor = np.einsum("kij,k,j->i", dM, qd, qd) - 0.5 * np.einsum("kij,i,j->k", dM, qd, qd)
    np.testing.assert_allclose(o.arr("qfrc_bias"), mjcf.gravity_bias_numpy(t, q) + cor, atol=2e-10)
    # qacc_smooth solves M a = qfrc_smooth
    np.testing.assert_allclose(M @ o.arr("qacc_smooth"), o.arr("qfrc_smooth"), atol=1e-13)


def test_constraint_rows_and_solver_optimality(oracle_mod, tables_v):
    O, t = oracle_mod, tables_v
    rng = np.random.default_rng(3)
    for trial in range(20):
        o = O.Oracle(t)
        o.reset()
        q = rng.uniform(-1.0, 1.0, 6)
        q[5] = abs(q[5])      # the gripper's lower limit is -0.1745
        if trial % 4 == 0:
            q[trial % 6] = t.jnt_range[trial % 6][1] + 0.0005   # inside the 1 mm impedance ramp of the upper limit
        o.set("qpos", q); o.set("qvel", rng.uniform(-2, 2, 6)); o.set("ctrl", rng.uniform(-2, 2, 6))
        o.set("qacc_warmstart", rng.uniform(-20, 20, 6))
        o.forward()
        nefc = o.d.nefc
        assert o.d.nf == 6 and nefc == 6 + (1 if trial % 4 == 0 else 0) + 4 * o.d.ncon   # friction, limit, pyramid rows
        R = o.arr("efc_R")[:nefc]
        np.testing.assert_allclose(R[:6], np.array(t.dof_invweight0[:]) / 9.0, rtol=1e-12)   # imp = 0.9
        np.testing.assert_allclose(o.arr("efc_aref")[:6], -(2 / (0.95 * 0.02)) * o.arr("qvel"), rtol=1e-13)
        f = o.arr("efc_force")[:nefc]
        assert np.all(np.abs(f[:6]) <= 0.052 + 1e-15)                 # friction within the loss
        assert np.all(f[6:] >= 0)                                     # limit forces push inward
        if trial % 4 == 0:
            x = 0.0005 / 0.001
            imp = 0.9 + 0.05 * (2 * x * x)
            np.testing.assert_allclose(R[6], (1 - imp) / imp * t.dof_invweight0[trial % 6], rtol=1e-12)
        # first-order optimality of the convex cost: M a - qfrc_smooth - J^T f = 0
        M = o.full_M()
        J = o.arr("efc_J")[:nefc]
        grad = M @ o.arr("qacc") - o.arr("qfrc_smooth") - J.T @ f
        assert np.linalg.norm(grad) / (t.meaninertia * 6) < 1e-6
        assert 0 <= o.d.solver_niter <= 10


def test_energy_decays_without_control(oracle_mod, tables_p):
    """ctrl = hold pose on scene B, no gravity work done by actuators beyond PD: velocities die out."""
    O, t = oracle_mod, tables_p
    o = O.Oracle(t)
    o.reset()
    q0 = np.array([0.2, -0.3, 0.4, 0.1, -0.2, 0.0])
    o.set("qpos", q0); o.set("qvel", np.array([1.0, -1.0, 0.5, 0.5, -0.5, 0.2])); o.set("ctrl", q0)
    o.step(2000)
    assert np.abs(o.arr("qvel")).max() < 1e-4      # soft (impedance 0.9) friction rows let the arm creep
    assert np.abs(o.arr("qpos") - q0).max() < 0.05      # PD + friction holds against gravity
    assert abs(o.d.time - 4.0) < 1e-9


def test_forward_does_not_touch_state_or_warmstart(oracle_mod, tables_v):
    O = oracle_mod
    o = O.Oracle(tables_v)
    o.reset(); o.set("qpos", [0.1] * 6); o.set("qacc_warmstart", [1.0] * 6)
    o.forward()
    assert list(o.arr("qpos")) == [0.1] * 6 and list(o.arr("qacc_warmstart")) == [1.0] * 6
    o.step()
    np.testing.assert_array_equal(o.arr("qacc_warmstart"), o.arr("qacc"))


@pytest.mark.parametrize("tag", ["v", "p"])
def test_oracle_pins(oracle_mod, tables_v, tables_p, tag):
    """Regression pins (tools/gen_golden.py).  They pin the oracle, not MuJoCo (parity unpinned)."""
    O = oracle_mod
    t = tables_v if tag == "v" else tables_p
    g = np.load(os.path.join(GOLD, f"oracle_{tag}.npz"))
    for i in range(4):
        o = O.Oracle(t)
        o.reset(); o.set("qpos", g["q"][i]); o.set("qvel", g["v"][i]); o.set("ctrl", g["u"][i])
        o.set("qacc_warmstart", g["w"][i]); o.forward()
        assert o.d.nefc == g["nefc"][i]
        np.testing.assert_allclose(o.full_M(), g["M"][i], rtol=0, atol=1e-16)
        for k in ("qfrc_bias", "site_xpos", "qacc_smooth", "qacc"):
            np.testing.assert_allclose(o.arr(k), g[k][i], rtol=1e-12, atol=1e-13)
    assert g["nefc"][3] > 6   # the fourth state sits beyond joint limits
    rows, final, iters = O.rollout(t, O.make_spec(kind=0, seed=7), 8, 20, 10)
    np.testing.assert_array_equal(rows[:, :, :5], g["rows"][:, :, :5])
    np.testing.assert_allclose(rows, g["rows"], atol=1e-7)
    np.testing.assert_allclose(final[:, :6], g["final"][:, :6], atol=1e-6)
    for kind, key in ((1, "rows_sin"), (2, "rows_chirp")):
        r, _, _ = O.rollout(t, O.make_spec(kind=kind, seed=7), 4, 20, 10)
        np.testing.assert_allclose(r, g[key], atol=1e-7)


def test_dataset_semantics(oracle_mod, tables_v):
    """Row layout, float32 rounding, ee lag, control ranges [REF SOARM101_DataCollection.py:90-136]."""
    O, t = oracle_mod, tables_v
    rows, final, _ = O.rollout(t, O.make_spec(kind=0, seed=11), 16, 5, 10)
    assert rows.shape == (16, 6, 13) and rows.dtype == np.float64
    u, ee, q = rows[:, :, :5], rows[:, :, 5:8], rows[:, :, 8:]
    assert np.all(np.abs(u) <= 0.5)
    assert np.all(np.abs(q[:, 0]) <= 0.3 + 1e-7)
    np.testing.assert_array_equal(rows[:, :, 5:], rows[:, :, 5:].astype(np.float32).astype(np.float64))
    # row 0: observation right after reset -> ee consistent with qpos
    for e in range(4):
        q6 = np.concatenate([q[e, 0], [0.0]])
        np.testing.assert_allclose(ee[e, 0], mjcf.site_numpy(t, q6), atol=2e-7)
    np.testing.assert_allclose(q[:, -1], final[:, :5], atol=1e-7)
    # sin / chirp generators stay within the amplitude range
    for kind in (1, 2):
        r, _, _ = O.rollout(t, O.make_spec(kind=kind, seed=5), 8, 30, 10)
        assert np.all(np.abs(r[:, :, :5]) <= 0.5 + 1e-12)
        assert np.abs(np.diff(r[:, :, :5], axis=1)).max() < 0.3   # smooth in time


def test_koopman_fingerprint(oracle_mod, tables_v):
    """The reference's trained DKUC model [REF results/SOARM101/11_27/DKUC/best_model.pt] predicts the
    oracle's one-step dynamics as well as it fitted real MuJoCo data (train pred_loss 9.4e-7,
    best_scores.json:5).  Wrong gains / friction / dt / observation order miss this by > 10x."""
    O = oracle_mod
    W = {k: v.astype(np.float64) for k, v in np.load(os.path.join(GOLD, "koopman_dkuc.npz")).items()}
    rows, _, _ = O.rollout(tables_v, O.make_spec(kind=0, seed=1), 1500, 20, 10)
    x, u, xn = rows[:, :-1, 5:].reshape(-1, 8), rows[:, :-1, :5].reshape(-1, 5), rows[:, 1:, 5:].reshape(-1, 8)
    h = x
    for i in range(5):
        h = h @ W[f"x_encode_net.linear_{i}.weight"].T + W[f"x_encode_net.linear_{i}.bias"]
        if i != 4:
            h = np.maximum(h, 0)
    z = np.concatenate([x, h], -1)
    pred = (z @ W["lA.weight"].T + u @ W["lB.weight"].T) @ W["lC.weight"].T
    mse, persistence = ((pred - xn) ** 2).mean(), ((x - xn) ** 2).mean()
    assert mse < 1.5e-6, mse                       # reference training value 9.4e-7; survey probe 7.6e-7
    assert persistence > 10 * mse
    dq = xn[:, 3:] - x[:, 3:]
    for j in range(5):
        A = np.stack([u[:, j], np.ones(len(u))], 1)
        slope = np.linalg.lstsq(A, dq[:, j], rcond=None)[0][0]
        assert abs(slope - W["lB.weight"][3 + j, j]) < 0.0015     # ~0.0175 rad per unit control


def test_koopman_long_horizon_fingerprint(oracle_mod, tables_v):
    """Second, stronger pin against the reference's own artefacts: the shipped model's 200-step open-loop MAE on
    its validation set is 6.84e-3 [REF results/SOARM101/11_27/DKUC/best_scores.json:7, train.py:35-87,
    models/losses.py:132-173].  Re-running that evaluation on oracle data reproduces it (6.8e-3) if and only if
    the data is generated with the gravity-compensation line of SOARM101Env.step active
    (`qfrc_applied = qfrc_bias`, [REF SOARM101_Env.py:120], commented out in the current tree): without it the
    arm sags 7e-4 rad per control step on the two gravity-loaded joints, which the model (trained on real MuJoCo
    data) does not predict, and the MAE is 3.2e-2.  So the artefact was trained with that line on, and the
    oracle's chatter-regime dynamics agree with real MuJoCo statistically over 2000 physics steps."""
    from lerobot_mujoco_sim2real_b200 import tables as T
    O = oracle_mod
    W = {k: v.astype(np.float64) for k, v in np.load(os.path.join(GOLD, "koopman_dkuc.npz")).items()}

    def predict(x, u):
        h = x
        for i in range(5):
            h = h @ W[f"x_encode_net.linear_{i}.weight"].T + W[f"x_encode_net.linear_{i}.bias"]
            if i != 4:
                h = np.maximum(h, 0)
        return (np.concatenate([x, h], -1) @ W["lA.weight"].T + u @ W["lB.weight"].T) @ W["lC.weight"].T

    out = {}
    for name, flags in (("plain", 0), ("gravcomp", T.ROLL_GRAVCOMP_HOLD)):
        rows, _, _ = O.rollout(tables_v, O.make_spec(kind=0, seed=3), 600, 200, 10, flags=flags)
        x, u = rows[:, :, 5:], rows[:, :, :5]
        x0, mae = x[:, 0], 0.0
        for i in range(200):
            x0 = predict(x0, u[:, i])
            mae += np.abs(x0 - x[:, i + 1]).mean()
        drift = (x[:, 1:, 3:] - x[:, :-1, 3:]).mean(axis=(0, 1))
        out[name] = (mae / 200, drift)
    mae_g, drift_g = out["gravcomp"]
    mae_p, drift_p = out["plain"]
    assert 5.5e-3 < mae_g < 8.5e-3, mae_g                 # reference: 6.84e-3
    assert np.abs(drift_g).max() < 1.5e-4                 # no sag, as the trained model implies
    assert mae_p > 3 * mae_g and drift_p[1] > 4e-4 and drift_p[2] > 4e-4



from oracle.koopman_oracle import PUBLISHED_TRAIN, check_training_loss_fingerprints  # noqa: E402


def test_training_loss_fingerprints(oracle_mod, tables_v):
    """koopman_loss, recon_loss, stable_Loss and train_total_loss of best_scores.json reproduced on oracle data (with
    the gravity-compensation line active, as for the evaluation metric above); without it they are 3x off."""
    from lerobot_mujoco_sim2real_b200 import tables as T
    O = oracle_mod
    W = {k: v.astype(np.float64) for k, v in np.load(os.path.join(GOLD, "koopman_dkuc.npz")).items()}
    rows, _, _ = O.rollout(tables_v, O.make_spec(kind=0, seed=11), 3000, 20, 10, flags=T.ROLL_GRAVCOMP_HOLD)
    check_training_loss_fingerprints(W, rows[:, :, 5:], rows[:, :, :5], "oracle:")
    rows, _, _ = O.rollout(tables_v, O.make_spec(kind=0, seed=11), 1000, 20, 10, flags=0)
    from oracle import koopman_oracle as KO
    plain = KO.k_linear_loss(W, rows[:, :, 5:], rows[:, :, :5], 7)
    assert plain["pred_loss"] > 2.0 * PUBLISHED_TRAIN["pred_loss"]


@pytest.mark.parametrize("scene", ["v", "p"])
def test_exact_line_search_is_equivalent(oracle_mod, tables_v, tables_p, scene):
    """The CUDA kernels replace MuJoCo's iterative PrimalSearch by the exact root of the piecewise-linear
    slope (csrc/so101_physics.cuh line_search).  On the CPU, with everything else identical: same Newton
    iteration counts, step results equal to ~1e-13, fewer than half the slope evaluations."""
    O = oracle_mod
    t = tables_v if scene == "v" else tables_p
    rng = np.random.default_rng(0)
    n = 256
    state = np.zeros((n, 18)); state[:, :5] = rng.uniform(-0.3, 0.3, (n, 5))
    S, U = [], []
    for s in range(300):
        if s % 10 == 0:
            ctrl = np.zeros((n, 6))
            ctrl[:, :5] = rng.uniform(-0.5, 0.5, (n, 5)) if scene == "v" else state[:, :5] + rng.uniform(-0.3, 0.3, (n, 5))
        if s % 50 == 0 and s:   # push some envs against a joint limit
            state[: n // 8, 2] = t.jnt_range[2][1] + rng.uniform(-1e-3, 5e-3, n // 8)
        S.append(state); U.append(ctrl)
        state, _, _ = O.step_batch(t, state, ctrl, 1)
    S, U = np.concatenate(S), np.concatenate(U)
    ref, _, aux = O.step_batch(t, S, U, 1)
    try:
        O.set_line_search(1)
        alt, _, aux2 = O.step_batch(t, S, U, 1)
    finally:
        O.set_line_search(0)
    assert (aux[:, 2] > 6).sum() > 100                      # limit rows were exercised
    assert np.array_equal(aux[:, 0], aux2[:, 0])            # identical Newton iteration counts
    assert aux2[:, 1].mean() < 0.75 * aux[:, 1].mean()
    rel = np.abs(alt - ref) / (1e-3 + np.abs(ref))
    assert rel[:, :6].max() < 1e-14 and rel[:, 6:12].max() < 5e-12 and rel[:, 12:].max() < 1e-9
    assert np.quantile(rel[:, 6:12], 0.999) < 1e-12


@pytest.mark.parametrize("scene", ["v", "p"])
def test_prox_start_is_equivalent(oracle_mod, tables_v, tables_p, scene):
    """The CUDA kernels start Newton at the per-dof prox point instead of MuJoCo's warm-start pick and use
    the exact line search: same unique minimiser, one iteration instead of three."""
    O = oracle_mod
    t = tables_v if scene == "v" else tables_p
    rng = np.random.default_rng(1)
    n = 256
    state = np.zeros((n, 18)); state[:, :5] = rng.uniform(-0.3, 0.3, (n, 5))
    S, U = [], []
    for s in range(300):
        if s % 10 == 0:
            ctrl = np.zeros((n, 6))
            ctrl[:, :5] = rng.uniform(-0.5, 0.5, (n, 5)) if scene == "v" else state[:, :5] + rng.uniform(-0.3, 0.3, (n, 5))
        if s % 50 == 0 and s:
            state[: n // 8, 2] = t.jnt_range[2][1] + rng.uniform(-1e-3, 5e-3, n // 8)
        S.append(state); U.append(ctrl)
        state, _, _ = O.step_batch(t, state, ctrl, 1)
    S, U = np.concatenate(S[20:]), np.concatenate(U[20:])
    ref, _, aux = O.step_batch(t, S, U, 1)
    try:
        O.set_line_search(1); O.set_solver_start(1)
        alt, _, aux2 = O.step_batch(t, S, U, 1)
    finally:
        O.set_line_search(0); O.set_solver_start(0)
    free = aux[:, 2] == 6                                   # no limit row active
    assert (~free).sum() > 100
    assert aux[free, 0].mean() > 1.7 and aux2[free, 0].mean() < 1.02     # 1 iteration instead of ~1.9-2.3
    assert aux2[:, 0].reshape(-1, 32).max(axis=1).mean() < 0.6 * aux[:, 0].reshape(-1, 32).max(axis=1).mean()
    rel = np.abs(alt - ref) / (1e-3 + np.abs(ref))
    assert rel[:, :6].max() < 1e-13
    assert np.quantile(rel[:, 6:12], 0.999) < 2e-12 and rel[:, 6:12].max() < 1e-10
    assert np.quantile(rel[:, 12:], 0.999) < 1e-10 and rel[:, 12:].max() < 1e-7


@pytest.mark.parametrize("scene", ["v", "p"])
def test_direct_active_set_is_equivalent(oracle_mod, tables_v, tables_p, scene):
    """The CUDA kernels first try a direct active-set solve (zones of the friction rows guessed per dof, one
    factorisation, KKT check) and only fall back to Newton from the prox point if the check fails
    (csrc/so101_physics.cuh active_set_guess / active_set_accept).  On the CPU, everything else identical: accepted in
    > 98 % of the steps without limit rows (> 99.8 % with the one retry from the zones of the rejected candidate),
    and the step result equals the literal MuJoCo solver's to ~1e-13."""
    O = oracle_mod
    t = tables_v if scene == "v" else tables_p
    rng = np.random.default_rng(2)
    n = 256
    state = np.zeros((n, 18)); state[:, :5] = rng.uniform(-0.3, 0.3, (n, 5))
    S, U = [], []
    for s in range(300):
        if s % 10 == 0:
            ctrl = np.zeros((n, 6))
            ctrl[:, :5] = rng.uniform(-0.5, 0.5, (n, 5)) if scene == "v" else state[:, :5] + rng.uniform(-0.3, 0.3, (n, 5))
        if s % 50 == 0 and s:
            state[: n // 8, 2] = t.jnt_range[2][1] + rng.uniform(-1e-3, 5e-3, n // 8)
        S.append(state); U.append(ctrl)
        state, _, _ = O.step_batch(t, state, ctrl, 1)
    S, U = np.concatenate(S[20:]), np.concatenate(U[20:])
    ref, _, aux = O.step_batch(t, S, U, 1)
    free = aux[:, 2] == 6                                   # no limit row active
    rates = []
    for retries in (0, 1):
        try:
            O.set_line_search(1); O.set_solver_start(2); O.lib().so101o_set_direct_retries(retries)
            alt, _, aux2 = O.step_batch(t, S, U, 1)
        finally:
            O.set_line_search(0); O.set_solver_start(0); O.lib().so101o_set_direct_retries(0)
        direct = free & (aux2[:, 0] == 1) & (aux2[:, 1] == 0)   # accepted: no Newton iteration, no line search
        rates.append(direct.sum() / free.sum())
    # first guess > 98 %; with one active-set retry (what the kernels do) > 99.8 %
    assert (~free).sum() > 100 and rates[0] > 0.98 and rates[1] > 0.998 and rates[1] >= rates[0]
    rel = np.abs(alt - ref) / (1e-3 + np.abs(ref))
    assert rel[:, :6].max() < 1e-13
    assert np.quantile(rel[:, 6:12], 0.999) < 2e-12 and rel[:, 6:12].max() < 1e-10
    assert np.quantile(rel[:, 12:], 0.999) < 1e-10 and rel[:, 12:].max() < 1e-7


def test_step_batch_matches_single_env(oracle_mod, tables_v):
    O, t = oracle_mod, tables_v
    rng = np.random.default_rng(0)
    state = np.zeros((8, 18)); state[:, :6] = rng.uniform(-0.5, 0.5, (8, 6)); state[:, 6:12] = rng.uniform(-1, 1, (8, 6))
    ctrl = rng.uniform(-1, 1, (8, 6))
    out, obs, aux = O.step_batch(t, state, ctrl, 3)
    o = O.Oracle(t)
    o.reset(); o.set("qpos", state[2, :6]); o.set("qvel", state[2, 6:12]); o.set("ctrl", ctrl[2]); o.step(3)
    np.testing.assert_array_equal(out[2, :6], o.arr("qpos"))
    np.testing.assert_array_equal(obs[2, :3], o.arr("site_xpos"))
    assert aux[2, 2] == 6
