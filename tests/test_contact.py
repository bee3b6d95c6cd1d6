"""SURVEY 8f N1, table-plane contact: host tables, hull data and the oracle's restatement (CPU); the CUDA contact path
against the oracle (`-m gpu`).  The reference scene enables contacts [REF SOARM101/SO101/scene_with_table_v.xml:28,31,
so101_new_calib_v.xml:53-117]; MuJoCo itself is not available, so the contact rows are restated from its published
algorithm (mjc_Convex -> one contact per geom pair, mj_instantiateContact pyramidal rows, mj_diagApprox,
mj_makeImpedance) and pinned by invariants here; tools/gen_mujoco_golden.py records `ct_*` vectors for the day a MuJoCo
is at hand."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def hulls():
    from lerobot_mujoco_sim2real_b200 import tables as T
    return T.builtin_hulls()


@pytest.fixture()
def contact_oracle(oracle_mod, hulls):
    oracle_mod.set_hulls(hulls)      # the session default; stated here because these tests are about it
    return oracle_mod


def _sample_states(n, seed, qmax=1.0):
    rng = np.random.default_rng(seed)
    s = np.zeros((n, 18))
    s[:, :6] = rng.uniform(-qmax, qmax, (n, 6))
    s[:, 5] = rng.uniform(-0.17, 1.0, n)
    s[:, 6:12] = rng.uniform(-1.0, 1.0, (n, 6))
    u = np.zeros((n, 6))
    u[:, :5] = rng.uniform(-0.5, 0.5, (n, 5))
    return s, u


def test_contact_tables(tables_v, tables_p, hulls):
    for t in (tables_v, tables_p):
        assert t.con_enabled == 1 and t.con_condim == 3 and t.ntrip == 10
        assert list(t.con_friction) == [1.0, 0.005, 0.0001] and list(t.con_solref) == [0.02, 1.0]
        assert list(t.con_solimp) == [0.9, 0.95, 0.001, 0.5, 2.0] and t.con_margin == 0.0
        assert list(t.con_box) == [-0.61, 0.61, -0.37, 0.37] and abs(t.trip_plane_z + 0.0009) < 1e-15
        w = np.ctypeslib.as_array(t.body_invweight0)
        assert np.all(w[:2] == 0) and np.all(np.diff(w[2:8, 0]) > 0) and np.all(w[2:8] > 0)   # lighter towards the tip
    assert hulls["vert_start"][-1] == hulls["vert"].shape[0] and hulls["adj_start"][-1] == hulls["adj"].shape[0]
    assert np.array_equal(hulls["trip_body"], np.ctypeslib.as_array(tables_v.trip_body)[:10])


def test_hull_walk_finds_the_support_vertex(hulls):
    """Steepest ascent on the hull's edge graph from the cube-map start == brute force, for random directions (what
    the CUDA path does against what the oracle does)."""
    from lerobot_mujoco_sim2real_b200 import tripwire
    res, V, AS, A = int(hulls["cube_res"]), hulls["vert"], hulls["adj_start"], hulls["adj"]
    rng = np.random.default_rng(1)
    steps = []
    for _ in range(4000):
        g = int(rng.integers(0, len(hulls["vert_start"]) - 1))
        d = rng.standard_normal(3)
        d /= np.linalg.norm(d)
        ax = int(np.argmax(np.abs(d)))
        u, w = d[(ax + 1) % 3] / abs(d[ax]), d[(ax + 2) % 3] / abs(d[ax])
        iu, iw = min(res - 1, max(0, int((u + 1) / 2 * res))), min(res - 1, max(0, int((w + 1) / 2 * res)))
        cur = int(hulls["cube"][g, 2 * ax + (d[ax] < 0), iu, iw])
        best, n = V[cur] @ d, 0
        while True:
            nb = A[AS[cur]:AS[cur + 1]]
            vals = V[nb] @ d
            k = int(np.argmax(vals))
            if vals[k] <= best:
                break
            best, cur, n = vals[k], int(nb[k]), n + 1
        assert cur == tripwire.support_numpy(hulls, g, d)[0]
        steps.append(n)
    assert np.mean(steps) < 2.0, np.mean(steps)


def test_oracle_contact_rows(contact_oracle, tables_v):
    """Row count, pyramid structure, Jacobian against finite differences of the witness point, one-sided forces, and
    agreement of the row parameters with the formulas the header of the oracle states."""
    O = contact_oracle
    o = O.Oracle(tables_v)
    o.reset(); o.set("qpos", [0.2, 0.45, 0.4, 0.35, 0.3, 0.1]); o.set("qvel", [0.3, -0.2, 0.1, 0.4, -0.5, 0.2])
    o.forward()
    d = o.d
    assert d.ncon >= 1 and d.con_unsupported == 0 and d.nefc == 6 + 4 * d.ncon
    assert list(d.efc_type[:6]) == [0] * 6 and set(d.efc_type[6:d.nefc]) == {2}
    J = o.arr("efc_J").copy()
    q0 = o.arr("qpos").copy()
    for c in range(d.ncon):
        g, vid, dist = d.con_geom[c], d.con_vert[c], d.con_dist[c]
        b = tables_v.trip_body[g]
        assert dist < 0 and abs(d.con_pos[c][2] - (tables_v.trip_plane_z + 0.5 * dist)) < 1e-15
        r0 = 6 + 4 * c
        Jn = 0.5 * (J[r0] + J[r0 + 1])
        Jt1, Jt2 = 0.5 * (J[r0] - J[r0 + 1]), 0.5 * (J[r0 + 2] - J[r0 + 3])           # mu = 1
        np.testing.assert_allclose(0.5 * (J[r0 + 2] + J[r0 + 3]), Jn, atol=1e-15)
        # finite differences of the world position of the point that sits at con_pos in the body frame
        from lerobot_mujoco_sim2real_b200 import mjcf
        xpos, xmat, _, _ = mjcf.fk_numpy(tables_v, q0)
        local = xmat[b].T @ (np.array(d.con_pos[c]) - xpos[b])
        jac = np.zeros((3, 6))
        for j in range(6):
            dq = np.zeros(6); dq[j] = 1e-6
            xp, xm, _, _ = mjcf.fk_numpy(tables_v, q0 + dq)
            xn, xmn, _, _ = mjcf.fk_numpy(tables_v, q0 - dq)
            jac[:, j] = ((xp[b] + xm[b] @ local) - (xn[b] + xmn[b] @ local)) / 2e-6
        np.testing.assert_allclose(Jn, jac[2], atol=2e-9)           # normal +z
        np.testing.assert_allclose(Jt1, jac[1], atol=2e-9)          # tangent (0, 1, 0)
        np.testing.assert_allclose(Jt2, -jac[0], atol=2e-9)         # tangent (-1, 0, 0)
        # impedance, R, aref of the four rows
        x = abs(dist) / 0.001
        y = 1.0 if x >= 1 else (2 * x * x if x <= 0.5 else 1 - 2 * (1 - x) ** 2)
        imp = 0.9 + 0.05 * y
        tran = tables_v.body_invweight0[b][0]
        R = 2.0 * max(1e-15, (1 - imp) / imp * 2 * tran)
        np.testing.assert_allclose(o.arr("efc_R")[r0:r0 + 4], R, rtol=1e-14)
        vel = J[r0:r0 + 4] @ o.arr("qvel")
        K, B = 1 / (0.95 ** 2 * 0.02 ** 2), 2 / (0.95 * 0.02)
        np.testing.assert_allclose(o.arr("efc_aref")[r0:r0 + 4], -B * vel - K * imp * dist, rtol=1e-13)
    f = o.arr("efc_force")[6:d.nefc]
    assert np.all(f >= 0) and f.max() > 0                           # pyramid rows only push
    # the solution is the minimiser: first-order optimality M a - qfrc_smooth - J'f = 0
    res = o.full_M() @ o.arr("qacc") - o.arr("qfrc_smooth") - o.arr("qfrc_constraint")
    assert np.abs(res).max() < 1e-6 * max(1.0, np.abs(o.arr("qfrc_smooth")).max())
    np.testing.assert_allclose(o.arr("qfrc_constraint"), J[:d.nefc].T @ o.arr("efc_force")[:d.nefc], atol=1e-12)


def test_oracle_arm_comes_to_rest_on_the_table(contact_oracle, tables_p):
    """Scene B (position servos): the arm is driven into the table and stays ON it - penetration stays at
    millimetres (soft contact, solref 0.02) where the contact-free model sinks 9 cm into it."""
    O = contact_oracle
    target = np.array([0.0, 0.6, 0.5, 0.5, 0.0, 0.0])
    depth = {}
    for with_hulls in (True, False):
        O.set_hulls(hulls_data() if with_hulls else None)
        o = O.Oracle(tables_p)
        o.reset(); o.set("ctrl", target)
        for _ in range(1500):
            o.step()
        O.set_hulls(hulls_data())
        s = np.zeros((1, 18)); s[0, :6] = o.arr("qpos")
        depth[with_hulls] = O.contact_probe(tables_p, s)[0, 3]
        if with_hulls:
            assert np.abs(o.arr("qvel")).max() < 1e-3 and o.d.ncon >= 1
    assert -5e-3 < depth[True] < 0 and depth[False] < -0.02, depth      # 1.7 mm under a 50 N push vs 92 mm


def hulls_data():
    from lerobot_mujoco_sim2real_b200 import tables as T
    return T.builtin_hulls()


def test_contact_statistics_match_the_survey(contact_oracle, tables_v):
    """SURVEY F5's probe: no contact in the reset box |q| <= 0.3, ~5 % of uniform poses in |q| <= 0.6, ~19 % in
    |q| <= 1.0 touch the table (exact hulls)."""
    O = contact_oracle
    frac = {}
    for qmax in (0.3, 0.6, 1.0):
        s, _ = _sample_states(3000, 5, qmax)
        s[:, 5] = np.clip(s[:, 5], -0.17, qmax)
        frac[qmax] = float((O.contact_probe(tables_v, s)[:, 0] > 0).mean())
    assert frac[0.3] == 0.0 and 0.02 < frac[0.6] < 0.09 and 0.15 < frac[1.0] < 0.24, frac


# ---------------------------------------------------------------------------------------------------------------------
# CUDA contact path vs the oracle
# ---------------------------------------------------------------------------------------------------------------------
def _gpu_step(tables, state, ctrl, nsub=1, family=0, dtype="float64"):
    import torch
    from lerobot_mujoco_sim2real_b200 import tables as T_
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    env = SOARM101VecEnv(tables=tables, num_envs=state.shape[0], dtype=dtype)
    assert env.model.has_contact
    env.set_option(T_.OPT_KERNEL_FAMILY, family)
    env.set_state(state[:, :6], state[:, 6:12], state[:, 12:18])
    u = torch.as_tensor(np.ascontiguousarray(ctrl.T), dtype=env.torch_dtype, device=env.device).contiguous()
    env.step_soa(u, nsub)
    q, v, w = env.get_state()
    out = np.concatenate([q.cpu().numpy(), v.cpu().numpy(), w.cpu().numpy()], axis=1).astype(np.float64)
    return out, env.flags().cpu().numpy(), env


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["v", "p"])
def test_contact_teacher_forced_parity(contact_oracle, tables_v, tables_p, scene):
    """One mj_step from 6000 states sampled in |q| <= 1.0 (19 % of them touch the table, up to five contacts at once):
    CUDA == oracle to 1e-12 (qpos) / 1e-9 (qvel) whether in contact or not; the CONTACT flag is raised exactly for the
    envs the oracle finds in contact; nothing is left to the tripwire."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    O = contact_oracle
    t = tables_v if scene == "v" else tables_p
    s, u = _sample_states(6000, 11)
    if scene == "p":
        u[:, :5] = s[:, :5] + np.random.default_rng(2).uniform(-0.3, 0.3, (6000, 5))
    s, _, _ = O.step_batch(t, s, u, 2)                  # two oracle steps: realistic warm starts, some envs settle in
    probe = O.contact_probe(t, s)
    ref, _, aux = O.step_batch(t, s, u, 1)
    out, flags, env = _gpu_step(t, s, u, 1)
    inc = probe[:, 0] > 0
    assert inc.mean() > 0.1 and probe[:, 1].max() == 0
    assert np.array_equal((flags & T_.FLAG_CONTACT) != 0, inc)
    assert not np.any(flags & T_.FLAG_TRIP_TABLE)
    den = np.maximum(np.abs(ref), 1e-3)
    err = np.abs(out - ref) / den
    for name, m in (("free", ~inc), ("contact", inc)):
        print(f"scene {scene} {name}: n={int(m.sum())} qpos {err[m, :6].max():.2e} qvel {err[m, 6:12].max():.2e} "
              f"qacc max {err[m, 12:].max():.2e} 99% {np.quantile(err[m, 12:], 0.99):.2e}")
    assert err[:, :6].max() < 1e-12 and err[:, 6:12].max() < 1e-9
    assert np.quantile(err[inc, 12:], 0.99) < 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_contact_both_kernel_families_same_bits(tables_v, dtype):
    from lerobot_mujoco_sim2real_b200 import tables as T_
    s, u = _sample_states(2000, 3)
    res = []
    for fam in (T_.FAMILY_ONEWARP, T_.FAMILY_TEAM):
        out, flags, _ = _gpu_step(tables_v, s, u, 10, family=fam, dtype=dtype)
        res.append((out, flags))
    assert (res[0][1] & T_.FLAG_CONTACT).any()
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])


@pytest.mark.gpu
def test_contact_free_running_scene_b_rests_on_table(contact_oracle, tables_p):
    """Free-running 1000 steps on the contractive scene with position targets that press the arm onto the table:
    CUDA ends where the oracle ends (stated tolerance 1e-6: stick-slip at the contact makes the late trajectory more
    sensitive than the contact-free P3 case), resting on the table with millimetre penetration."""
    O = contact_oracle
    n = 256
    rng = np.random.default_rng(9)
    s = np.zeros((n, 18)); s[:, :5] = rng.uniform(-0.3, 0.3, (n, 5))
    u = np.zeros((n, 6)); u[:, 1:4] = rng.uniform(0.4, 0.8, (n, 3)); u[:, 0] = rng.uniform(-0.5, 0.5, n)
    ref, _, _ = O.step_batch(tables_p, s, u, 1000)
    out, flags, _ = _gpu_step(tables_p, s, u, 1000)
    probe = O.contact_probe(tables_p, ref)
    assert (probe[:, 0] > 0).mean() > 0.5 and probe[:, 3].min() > -5e-3
    err = np.abs(out[:, :12] - ref[:, :12])
    print(f"resting on the table after 1000 steps: max |dq| {err[:, :6].max():.2e}, max |dqvel| {err[:, 6:].max():.2e}, "
          f"in contact {float((probe[:, 0] > 0).mean()):.2f}, deepest {probe[:, 3].min() * 1e3:.3f} mm")
    assert err[:, :6].max() < 1e-6 and err[:, 6:].max() < 1e-4


@pytest.mark.gpu
def test_chirp_test_set_needs_no_tripwire_for_the_table(tables_v):
    """8.8 % of a chirp T=200 test set passed through the table.  With the contact path the table flag is
    raised only for contacts the kernels cannot represent (an edge of the table, more than 6 at once)."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    n = 16384
    env = SOARM101VecEnv(tables=tables_v, num_envs=n)
    env.rollout(200, "chirp", seed=42)
    fl = env.flags().cpu().numpy()
    inc, table = (fl & T_.FLAG_CONTACT) != 0, (fl & T_.FLAG_TRIP_TABLE) != 0
    print(f"chirp T=200, {n} envs: in contact at some time {inc.mean():.3f}, unsimulated table contact {table.mean():.5f}, "
          f"self-collision box left {((fl & T_.FLAG_TRIP_SELF) != 0).mean():.3f}, bad {((fl & T_.FLAG_BADSTATE) != 0).mean():.5f}")
    assert inc.mean() > 0.03 and table.mean() < 2e-3 and not np.any(fl & T_.FLAG_BADSTATE)
    # without hull data the same run only flags those envs (round-1 behaviour)
    env2 = SOARM101VecEnv(tables=tables_v, num_envs=n, hulls=None)
    env2.rollout(200, "chirp", seed=42)
    fl2 = env2.flags().cpu().numpy()
    assert not np.any(fl2 & T_.FLAG_CONTACT) and ((fl2 & T_.FLAG_TRIP_TABLE) != 0).mean() > 0.03


@pytest.mark.gpu
def test_contact_statistics_free_running_chirp_match_the_oracle(contact_oracle, tables_v):
    """Scene A is chaotic, so free-running trajectories cannot be compared state by state beyond ~100 steps;
    their STATISTICS can: 2048 chirp rollouts of 200 control steps on both sides - the fraction of envs that ever touch the
    table, the fraction touching at the end, and the deepest penetration at the end (soft contact, solref 0.02: millimetres)
    agree, and nothing sinks through the table on either side."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    O = contact_oracle
    n, Tn = 2048, 200
    env = SOARM101VecEnv(tables=tables_v, num_envs=n)
    env.rollout(Tn, "chirp", seed=42)
    q, v, w = env.get_state()
    fin_gpu = np.concatenate([q.cpu().numpy(), v.cpu().numpy(), w.cpu().numpy()], axis=1)
    fl = env.flags().cpu().numpy()
    _, fin_ref, _ = O.rollout(tables_v, O.make_spec(kind=2, seed=42), n, Tn, 10, want_rows=False)
    pg, pr = O.contact_probe(tables_v, fin_gpu), O.contact_probe(tables_v, fin_ref)
    end_gpu, end_ref = float((pg[:, 0] > 0).mean()), float((pr[:, 0] > 0).mean())
    ever_gpu = float(((fl & T_.FLAG_CONTACT) != 0).mean())
    print(f"chirp T=200, {n} envs: touching at the end gpu {end_gpu:.4f} oracle {end_ref:.4f}; ever in contact (gpu) {ever_gpu:.4f}; "
          f"deepest penetration at the end gpu {pg[:, 3].min() * 1e3:.2f} mm oracle {pr[:, 3].min() * 1e3:.2f} mm")
    assert end_ref > 0.01 and abs(end_gpu - end_ref) < 0.25 * end_ref + 0.005
    assert ever_gpu >= end_gpu and ever_gpu > 0.02
    assert pg[:, 3].min() > -5e-3 and pr[:, 3].min() > -5e-3          # soft contact holds: no arm below the table top by > 5 mm
    assert not np.any(fl & T_.FLAG_BADSTATE) and np.isfinite(fin_gpu).all()


@pytest.mark.gpu
def test_the_table_holds_at_full_size(tables_v, hulls):
    """A property that needs no oracle: after a test-shaped chirp rollout of a config-4 shard (131072 envs, 2000 physics steps,
    regrouped launch) no collision hull of an env that touched the table lies deeper in it than the soft contact allows
    (millimetres: solref / solimp of the scene), measured with the exact hull vertices in numpy at the final pose - while the
    same rollout without the contact path leaves arms centimetres inside the table."""
    from lerobot_mujoco_sim2real_b200 import mjcf, tables as T_
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    t = tables_v
    n = 131072

    def deepest(q):
        xpos, xmat, _, _ = mjcf.fk_numpy(t, q)
        z = np.inf
        for g in range(t.ntrip):
            b = t.trip_body[g]
            v = hulls["vert"][hulls["vert_start"][g]:hulls["vert_start"][g + 1]]
            z = min(z, float((xpos[b][2] + v @ xmat[b][2]).min()))
        return z - t.trip_plane_z

    res = {}
    for name, h in (("on", "auto"), ("off", None)):
        env = SOARM101VecEnv(tables=t, num_envs=n, hulls=h)
        env.rollout_discard(200, "chirp", seed=42)
        fl = env.flags().cpu().numpy()
        q = env.get_state()[0].cpu().numpy()
        touched = np.nonzero(fl & (T_.FLAG_CONTACT if h else T_.FLAG_TRIP_TABLE))[0]
        assert len(touched) > 0.03 * n and not np.any(fl & T_.FLAG_BADSTATE)
        pick = touched[np.linspace(0, len(touched) - 1, 1200).astype(int)]
        res[name] = np.array([deepest(q[i]) for i in pick])
    print(f"deepest hull point below the table top, 1200 envs that touched it: contact on {res['on'].min() * 1e3:.2f} mm, "
          f"off {res['off'].min() * 1e3:.1f} mm")
    assert res["on"].min() > -5e-3
    assert res["off"].min() < -2e-2
