"""GPU: the reference-facing API (SOARM101Env / SOARM101VecEnv / SOARM101DataGenerator / shoot)
through the C ABI, against the CPU oracle and against size-independent properties."""
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _vec(tables, n, dtype="float64", **kw):
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    return SOARM101VecEnv(tables=tables, num_envs=n, dtype=dtype, **kw)


# ------------------------------------------------------------------------------------------------
# rollout kernel == CPU restatement of generate_physics_based_data
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,name", [(0, "random"), (1, "sin"), (2, "chirp")])
def test_rollout_rows_match_oracle(oracle_mod, tables_v, kind, name):
    O = oracle_mod
    n, T = 256, 5          # 50 physics steps: before chaos amplifies rounding
    env = _vec(tables_v, n)
    rows = env.rollout(T, name, seed=123, env_offset=1000).cpu().numpy()
    ref, final, _ = O.rollout(tables_v, O.make_spec(kind=kind, seed=123, env_offset=1000), n, T, 10)
    assert rows.shape == (n, T + 1, 13) and rows.dtype == np.float64
    if kind == 0:
        np.testing.assert_array_equal(rows[:, :, :5], ref[:, :, :5])       # Philox stream is bit-exact
    else:
        np.testing.assert_allclose(rows[:, :, :5], ref[:, :, :5], atol=1e-15)  # device sin() vs libm
    np.testing.assert_array_equal(rows[:, 0, 5:], ref[:, 0, 5:])            # reset observation
    # float32-rounded observations: equal up to one float32 ulp flip
    assert np.abs(rows[:, :, 5:] - ref[:, :, 5:]).max() <= 6e-8
    q, v, w = env.get_state()
    np.testing.assert_allclose(q.cpu().numpy(), final[:, :6], atol=1e-11)
    np.testing.assert_allclose(v.cpu().numpy(), final[:, 6:12], atol=1e-8)
    st = env.stats()
    assert st["physics_steps"] == n * T * 10


def test_rollout_is_deterministic_and_shard_invariant_full_size(tables_v):
    """BASELINE config 2 size (4096 envs x 1000 physics steps): bit-identical reruns, and two
    half batches with env_offset == one full batch (what makes the multi-GPU gather exact)."""
    n, T = 4096, 100
    env = _vec(tables_v, n)
    a = env.rollout(T, "random", seed=42)
    b = env.rollout(T, "random", seed=42)
    assert torch.equal(a, b)
    half = _vec(tables_v, n // 2)
    lo = half.rollout(T, "random", seed=42, env_offset=0)
    hi = half.rollout(T, "random", seed=42, env_offset=n // 2)
    assert torch.equal(torch.cat([lo, hi]), a)
    assert not torch.equal(a, env.rollout(T, "random", seed=43))
    r = a.cpu().numpy()
    assert np.isfinite(r).all()
    assert np.abs(r[:, :, :5]).max() <= 0.5 and np.abs(r[:, 0, 8:]).max() <= 0.3 + 1e-7
    assert np.array_equal(r[:, :, 5:], r[:, :, 5:].astype(np.float32).astype(np.float64))
    from lerobot_mujoco_sim2real_b200 import tables as T_
    assert int((env.flags() & T_.FLAG_BADSTATE).ne(0).sum()) == 0


def test_step_calls_equal_fused_rollout(tables_v):
    """T x SOARM101VecEnv.step(u_t) == one fused rollout with the same controls (bitwise)."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    n, T = 512, 6
    g = torch.Generator().manual_seed(0)
    U = (torch.rand((T + 1, 5, n), generator=g, dtype=torch.float64) - 0.5).cuda().contiguous()
    init = torch.zeros((n, 10), dtype=torch.float64)
    init[:, :5] = (torch.rand((n, 5), generator=g, dtype=torch.float64) - 0.5) * 0.6
    a = _vec(tables_v, n)
    a.reset(options={"initial_state": init})
    rows = a.rollout(T, "tensor", u=U, flags=T_.ROLL_NO_RESET).cpu().numpy()
    b = _vec(tables_v, n)
    obs0, info = b.reset(options={"initial_state": init})
    assert info == {} and obs0.shape == (n, 8) and obs0.dtype == torch.float32
    np.testing.assert_array_equal(obs0.cpu().numpy().astype(np.float64), rows[:, 0, 5:])
    for t in range(T):
        obs, rew, term, trunc, info = b.step(U[t].t())
        assert (rew, term, trunc, info) == (0.0, False, False, {})
        np.testing.assert_array_equal(obs.cpu().numpy().astype(np.float64), rows[:, t + 1, 5:])
    for x, y in zip(a.get_state(), b.get_state()):
        assert torch.equal(x, y)


def test_shoot_matches_oracle(oracle_mod, tables_v):
    """Config 5 shape (reduced horizon for the oracle): B sequences from one shared state."""
    O = oracle_mod
    from lerobot_mujoco_sim2real_b200 import tables as T_
    B, H = 1024, 4
    rng = np.random.default_rng(1)
    s0 = np.zeros(18); s0[:5] = rng.uniform(-0.3, 0.3, 5)
    s0_out, _, _ = O.step_batch(tables_v, s0[None], np.zeros((1, 6)), 30)   # warm-up: non-trivial qvel / warmstart
    s0 = s0_out[0]
    U = rng.uniform(-0.5, 0.5, (H, 5, B))
    for flags in (0, T_.ROLL_GRAVCOMP_HOLD):
        env = _vec(tables_v, B)
        X = env.shoot(s0, torch.as_tensor(U).cuda().contiguous(), flags=flags).cpu().numpy()
        ref = O.shoot(tables_v, s0, U, 10, flags=flags)
        assert X.shape == (B, H + 1, 8) and X.dtype == np.float32
        assert np.abs(X - ref).max() <= 2e-7
        assert np.abs(X[:, 0] - X[0, 0]).max() == 0        # shared initial observation
    # gravity compensation changes the motion
    assert np.abs(X - env.shoot(s0, torch.as_tensor(U).cuda().contiguous(), flags=0).cpu().numpy()).max() > 1e-4


def test_joint_limits_teacher_forced(contact_free, tables_v):
    """States beyond / inside the 1 mm impedance ramp of joint limits: limit rows vs the oracle (contact-free pipeline
    on both sides: a joint at its limit puts half the arm through the table, tests/test_contact.py covers that)."""
    O = contact_free
    from lerobot_mujoco_sim2real_b200 import tables as T_
    t = tables_v
    n = 2048
    rng = np.random.default_rng(4)
    state = np.zeros((n, 18))
    state[:, :6] = rng.uniform(-0.5, 0.5, (n, 6)); state[:, 5] = np.abs(state[:, 5])
    j = rng.integers(0, 6, n); side = rng.integers(0, 2, n)
    depth = rng.choice([1e-5, 3e-4, 7e-4, 2e-3, 2e-2], n)
    for e in range(n):
        state[e, j[e]] = t.jnt_range[j[e]][side[e]] + (depth[e] if side[e] else -depth[e])
    state[:, 6:12] = rng.uniform(-2, 2, (n, 6))
    state[:, 12:] = rng.uniform(-30, 30, (n, 6))
    ctrl = rng.uniform(-2.5, 2.5, (n, 6))
    ref, _, aux = O.step_batch(t, state, ctrl, 1)
    assert (aux[:, 2] >= 7).all()
    env = _vec(t, n, hulls=None)
    env.set_state(state[:, :6], state[:, 6:12], state[:, 12:])
    env.step_soa(torch.as_tensor(ctrl.T.copy()).cuda().contiguous(), 1)
    q, v, w = [x.cpu().numpy() for x in env.get_state()]
    assert np.abs(q - ref[:, :6]).max() < 1e-13
    rel = np.abs(v - ref[:, 6:12]) / (1e-3 + np.abs(ref[:, 6:12]))
    assert np.quantile(rel, 0.99) < 1e-11 and rel.max() < 1e-8
    assert int((env.flags() & T_.FLAG_LIMIT).ne(0).sum()) == n
    assert env.stats()["limit_steps"] == n


def test_bad_state_is_flagged_and_reset(tables_v):
    from lerobot_mujoco_sim2real_b200 import tables as T_
    env = _vec(tables_v, 64)
    q = np.zeros((64, 6)); q[3, 2] = np.nan; q[7, 0] = 1e11
    env.set_state(q, np.zeros((64, 6)), np.zeros((64, 6)))
    env.step_soa(torch.zeros((5, 64), dtype=torch.float64, device="cuda"), 2)
    f = env.flags().cpu().numpy()
    assert set(np.nonzero(f & T_.FLAG_BADSTATE)[0]) == {3, 7}
    qq = env.get_state()[0].cpu().numpy()
    assert np.isfinite(qq).all()


# ------------------------------------------------------------------------------------------------
# SOARM101Env drop-in (N = 1, host numpy in/out)
# ------------------------------------------------------------------------------------------------
def _write_scene(tmp_path, tables):
    """A self-contained MJCF equivalent to the packed tables is not needed: compile from the XML
    when the reference tree is present, else skip (the GPU box has no /root/reference)."""
    import os
    p = "/root/reference/SOARM101/SO101/scene_with_table_v.xml"
    if not os.path.exists(p):
        pytest.skip("reference XML not present on this box")
    return p


def test_soarm101env_drop_in_with_tables(oracle_mod, tables_v, monkeypatch):
    """The Env class behind pre-compiled tables (no XML on the GPU box): reference semantics."""
    O = oracle_mod
    from lerobot_mujoco_sim2real_b200 import SOARM101_Env as E, mjcf

    def fake_compile(path, site_name="gripperframe"):
        return mjcf.CompiledModel(
            tables=tables_v, body_names=["world", "base", "shoulder", "upper_arm", "lower_arm", "wrist", "gripper", "jaw"],
            joint_names=["shoulder_pan", "shoulder_lift", "elbow_flex", "wrist_flex", "wrist_roll", "gripper"],
            actuator_names=[], site_names=["baseframe", "gripperframe"], site_body=[1, 6], site_pos=[], key_names=["home"],
            geoms=[], mesh_files={}, meshdir="", xml_dir="", M0=np.eye(6))
    monkeypatch.setattr(mjcf, "compile_mjcf", fake_compile)
    env = E.SOARM101Env(xml_path="scene_with_table_v.xml")
    assert env.frame_skip == 10 and abs(env.dt - 0.02) < 1e-15
    assert env.udim == 5 and env.xdim == 8 and env.joint_ids == [0, 1, 2, 3, 4] and env.ee_site_id == 1
    assert env.action_space.shape == (5,) and env.observation_space.shape == (8,)
    init = np.array([0.1, -0.2, 0.3, -0.1, 0.2, 0.5, -0.5, 0.25, 0.0, 0.1])
    obs, info = env.reset(options={"initial_state": init})
    assert obs.dtype == np.float32 and obs.shape == (8,) and info == {}
    q6 = np.concatenate([init[:5], [0.0]])
    np.testing.assert_allclose(obs[:3], mjcf.site_numpy(tables_v, q6), atol=2e-7)     # ee first ...
    np.testing.assert_array_equal(obs[3:], init[:5].astype(np.float32))               # ... then qpos
    state = np.zeros((1, 18)); state[0, :5] = init[:5]; state[0, 6:11] = init[5:]
    ctrl = np.zeros((1, 6))
    for k in range(3):
        a = np.array([0.4, -0.3, 0.2, 0.1, -0.5]) * (k + 1) / 3
        obs, rew, term, trunc, info = env.step(a)
        ctrl[0, :5] = a
        state, ref_obs, _ = O.step_batch(tables_v, state, ctrl, 10)
        assert (rew, term, trunc, info) == (0.0, False, False, {})
        assert np.abs(obs - ref_obs[0].astype(np.float32)).max() <= 1.2e-7
    # env.data views [REF Koopman_MPC.py:89,119,186]
    np.testing.assert_allclose(np.asarray(env.data.qpos), state[0, :6], atol=1e-12)
    o = O.Oracle(tables_v)
    o.reset(); o.set("qpos", state[0, :6]); o.set("qvel", state[0, 6:12]); o.forward()
    np.testing.assert_allclose(env.data.qfrc_bias, o.arr("qfrc_bias"), atol=1e-12)
    env.data.qfrc_applied[:] = env.data.qfrc_bias[:]            # gravity compensation idiom
    env.data.qpos[:5] = [0.0, 0.1, 0.2, 0.3, 0.4]               # write-through
    np.testing.assert_allclose(env.data.qpos[:5], [0.0, 0.1, 0.2, 0.3, 0.4])
    np.testing.assert_allclose(env.model.key_qpos[0], [-0.04137, -1.68611, 1.69453, 0.5, 0.01833, 0])
    assert env.model.opt.timestep == 0.002 and env.model.joint("elbow_flex").id == 2
    # seeded default reset: gymnasium-style np_random, U(-0.3, 0.3), zero velocity
    o1, _ = env.reset(seed=5)
    o2, _ = env.reset(seed=5)
    np.testing.assert_array_equal(o1, o2)
    assert np.abs(o1[3:]).max() <= 0.3 and np.allclose(np.asarray(env.data.qvel), 0)
    env.close()


def test_data_generator_drop_in(oracle_mod, tables_v, tmp_path):
    """generate_physics_based_data: shape/dtype/columns, chunking invariance, cache files, loaders."""
    from lerobot_mujoco_sim2real_b200.SOARM101_DataCollection import SOARM101DataGenerator
    d = tmp_path / "data"
    args = types.SimpleNamespace(
        xml_path="unused", x_dim=8, u_dim=5, device="cuda", seed=42, env="SOARM101",
        train_samples=300, train_steps=4, test_samples=50, test_steps=6, batch_size=32, eval_batch_size=16,
        data_dir_save=str(d), data_dir_load_train=str(d / "train_data_300_4.npy"),
        data_dir_load_val=str(d / "val_data_50_6.npy"))
    gen = SOARM101DataGenerator(args, tables=tables_v)
    data = gen.generate_physics_based_data(300, 4, "random", seed=9)
    assert isinstance(data, np.ndarray) and data.shape == (300, 5, 13) and data.dtype == np.float64
    ref, _, _ = oracle_mod.rollout(tables_v, oracle_mod.make_spec(kind=0, seed=9), 300, 4, 10)
    np.testing.assert_array_equal(data[:, :, :5], ref[:, :, :5])
    assert np.abs(data - ref).max() <= 6e-8
    gen2 = SOARM101DataGenerator(args, tables=tables_v)
    gen2.max_batch = 128                                     # 3 launches instead of 1
    np.testing.assert_array_equal(gen2.generate_physics_based_data(300, 4, "random", seed=9), data)
    gen.generate_and_save_data()
    names = sorted(p.name for p in d.iterdir())
    assert names == ["test_data_chirp_50_6.npy", "test_data_random_50_6.npy", "test_data_sin_50_6.npy",
                     "train_data_300_4.npy", "val_data_50_6.npy"]
    assert np.load(d / "train_data_300_4.npy").shape == (300, 5, 13)
    train_loader, val_loader = gen.get_train_loader()
    batch = next(iter(train_loader))
    assert batch["x"].shape == (32, 5, 8) and batch["u"].shape == (32, 5, 5) and batch["x"].dtype == torch.float32
    assert next(iter(gen.get_test_loader("sin")))["x"].shape == (16, 7, 8)
    # cached files are reused, not regenerated
    before = (d / "train_data_300_4.npy").stat().st_mtime_ns
    SOARM101DataGenerator(args, tables=tables_v).generate_and_save_data()
    assert (d / "train_data_300_4.npy").stat().st_mtime_ns == before


def test_data_generator_never_returns_flagged_trajectories(tables_v, tmp_path):
    """The scene has contacts enabled: a trajectory that reaches a contact the simulator does not model is not MuJoCo
    data.  Default policy: such trajectories are replaced by fresh flag-free ones (deterministically, same shape);
    'keep' returns them with the mask (and generate_and_save_data writes the mask file); 'raise' raises."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    from lerobot_mujoco_sim2real_b200.SOARM101_DataCollection import ContactError, SOARM101DataGenerator
    d = tmp_path / "data"
    args = types.SimpleNamespace(
        xml_path="unused", x_dim=8, u_dim=5, device="cuda", seed=42, env="SOARM101",
        train_samples=64, train_steps=2, test_samples=400, test_steps=200, batch_size=32, eval_batch_size=16,
        data_dir_save=str(d), data_dir_load_train=str(d / "train_data_64_2.npy"),
        data_dir_load_val=str(d / "val_data_400_200.npy"))
    BAD = SOARM101DataGenerator.BAD_FLAGS
    # tables WITHOUT contact parameters: the kernels only flag table contacts (with them they are simulated and a flag
    # is the exception: an edge of the table, the self-collision box), which gives this test plenty to replace
    tables_v = T_.tables_from_dict(T_.tables_to_dict(tables_v))
    tables_v.con_enabled = 0
    keep = SOARM101DataGenerator(args, tables=tables_v, on_contact="keep")
    raw = keep.generate_physics_based_data(400, 200, "chirp", seed=5)          # long chirp runs do leave the safe region
    flags = keep.last_flags.copy()
    nbad = int(np.count_nonzero(flags & BAD))
    assert flags.shape == (400,) and nbad > 0, "test premise: some chirp trajectories are flagged"
    gen = SOARM101DataGenerator(args, tables=tables_v)                         # default: regenerate
    data = gen.generate_physics_based_data(400, 200, "chirp", seed=5)
    assert data.shape == raw.shape and gen.last_replaced == nbad
    good = (flags & BAD) == 0
    np.testing.assert_array_equal(data[good], raw[good])                       # untouched where nothing was flagged
    assert not np.array_equal(data[~good], raw[~good])
    # every replacement is a flag-free trajectory of a fresh env index, in order
    env_flags = []
    cand, f = gen.generate_device(nbad + nbad // 4 + 32, 200, "chirp", seed=5, env_offset=400)
    okc = cand[(f & BAD) == 0][:nbad].cpu().numpy()
    if len(okc) == nbad:
        np.testing.assert_array_equal(data[~good], okc)
    np.testing.assert_array_equal(gen.generate_physics_based_data(400, 200, "chirp", seed=5), data)   # deterministic
    with pytest.raises(ContactError):
        SOARM101DataGenerator(args, tables=tables_v, on_contact="raise").generate_physics_based_data(400, 200, "chirp", seed=5)
    keep.generate_and_save_data()
    masks = sorted(p.name for p in d.iterdir() if p.name.endswith(".flags.npy"))
    assert masks, "on_contact='keep' must write the mask next to a dataset that contains flagged trajectories"
    for name in masks:
        m = np.load(d / name)
        assert m.shape[0] == np.load(d / name.replace(".flags.npy", ".npy")).shape[0]


def test_env_step_reports_abnormal_flags_in_info(tables_v):
    """SOARM101Env.step returns {} like the reference while the env is clean, and the status word once it is not."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    import ctypes as C
    from lerobot_mujoco_sim2real_b200 import _lib
    env = SOARM101VecEnv(tables=tables_v, num_envs=2)
    q = np.zeros((2, 6)); q[1, :5] = [0.0, 1.2, 1.2, 0.0, 0.0]                  # env 1 far outside the safe joint box
    env.set_state(q, np.zeros((2, 6)), np.zeros((2, 6)))
    obs = np.zeros((8, 2), dtype=np.float32); fl = np.zeros(2, dtype=np.uint32); u = np.zeros((6, 2))
    _lib.check(_lib.lib().so101_batch_step_host(env._h, u.ctypes.data, 6, 1, obs.ctypes.data, fl.ctypes.data, env._stream()))
    assert fl[0] & T_.FLAG_ABNORMAL == 0 and fl[1] & T_.FLAG_ABNORMAL != 0
    assert np.array_equal(fl.astype(np.int32), env.flags().cpu().numpy())


def test_fp32_free_running_scene_b(contact_free, tables_p):
    """fp32 mode, stated tolerance: 1000 physics steps on the contractive scene within 5e-6 rad / 1e-4 rad/s
    (measured 3.5e-7 / 4.1e-6); contact-free pipeline on both sides."""
    O = contact_free
    from lerobot_mujoco_sim2real_b200 import tables as T_
    n, T = 512, 100
    rng = np.random.default_rng(6)
    q0 = np.zeros((n, 6)); q0[:, :5] = rng.uniform(-0.3, 0.3, (n, 5))
    U = np.cumsum(rng.uniform(-0.05, 0.05, (T + 1, 5, n)), axis=0) + q0[:, :5].T[None]
    _, fin, _ = O.rollout(tables_p, O.make_spec(kind=3, u=np.ascontiguousarray(U)), n, T, 10, qpos0=q0, want_rows=False)
    env = _vec(tables_p, n, dtype="float32", hulls=None)
    env.set_state(q0, np.zeros((n, 6)), np.zeros((n, 6)))
    env.rollout(T, "tensor", u=torch.as_tensor(U, dtype=torch.float32).cuda().contiguous(), flags=T_.ROLL_NO_RESET)
    q, v, _ = env.get_state()
    eq = np.abs(q.cpu().numpy() - fin[:, :6]).max()
    ev = np.abs(v.cpu().numpy() - fin[:, 6:12]).max()
    print(f"fp32 scene B after 1000 physics steps: |dq| {eq:.2e} |dqvel| {ev:.2e}")
    assert eq < 5e-6 and ev < 1e-4


def test_fma_peak_is_plausible():
    from lerobot_mujoco_sim2real_b200.vec_env import fma_peak_tflops
    p64, p32 = fma_peak_tflops("float64"), fma_peak_tflops("float32")
    assert 20 < p64 < 45 and 40 < p32 < 90


@pytest.mark.parametrize("family", ["onewarp", "team"])
def test_contact_tripwire_flags(tables_v, family):
    """Without hull data (no contact path): TRIP_TABLE <=> some collision box below the table plane (numpy
    restatement); TRIP_SELF <=> q outside the joint box of the fast accept AND the oriented boxes of two colliding geoms on
    non-adjacent links overlap (tripwire.self_overlap_numpy).  Evaluated at the pose the step starts from.  With
    hull data the same boxes select the hulls that get the exact test, and TRIP_TABLE is left for contacts the kernels
    cannot represent (tests/test_contact.py)."""
    from lerobot_mujoco_sim2real_b200 import tables as T_, tripwire
    t = tables_v
    n = 4096
    rng = np.random.default_rng(8)
    q = rng.uniform(-1.0, 1.0, (n, 6)); q[:, 5] = np.clip(q[:, 5], -0.17, None)
    q[: n // 4, :5] = rng.uniform(-0.3, 0.3, (n // 4, 5)); q[: n // 4, 5] = 0     # the reset box
    lo_r = np.array([t.jnt_range[k][0] for k in range(6)]); hi_r = np.array([t.jnt_range[k][1] for k in range(6)])
    q[n // 2:] = rng.uniform(np.maximum(lo_r, -1.6), np.minimum(hi_r, 1.6), (n - n // 2, 6))   # folded poses: boxes do overlap
    env = _vec(t, n, hulls=None)
    env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_ONEWARP if family == "onewarp" else T_.FAMILY_TEAM)
    env.set_state(q, np.zeros((n, 6)), np.zeros((n, 6)))
    env.clear_flags()
    env.step_soa(torch.zeros((5, n), dtype=torch.float64, device="cuda"), 1)
    f = env.flags().cpu().numpy()
    clear = np.array([tripwire.table_clearance_numpy(t, q[i]) for i in range(n)])
    sure = np.abs(clear) > 1e-9
    assert np.array_equal(((f & T_.FLAG_TRIP_TABLE) != 0)[sure], (clear < 0)[sure])
    lo = np.array([t.trip_qbox[k][0] for k in range(6)]); hi = np.array([t.trip_qbox[k][1] for k in range(6)])
    outside = ((q < lo) | (q > hi)).any(axis=1)
    overlap = np.array([outside[i] and tripwire.self_overlap_numpy(t, q[i]) for i in range(n)])
    assert np.array_equal((f & T_.FLAG_TRIP_SELF) != 0, overlap)
    assert 20 < overlap.sum() < 0.2 * n and outside.mean() > 0.5        # the flag is rare outside the box, and exercised
    assert (f[: n // 4] & T_.FLAG_TRIP).sum() == 0            # nothing trips inside the reset box
    assert 0.1 < (clear[: n // 2] < 0).mean() < 0.4           # ~19 % of |q| <= 1 poses touch the table


def test_self_collision_flag_is_sticky_and_float32_agrees(tables_v):
    """The flag stays once raised (an env that folded onto itself and came back is still not what MuJoCo would have produced),
    and the float32 kernels decide like the float64 ones except on poses within rounding of the boxes touching."""
    from lerobot_mujoco_sim2real_b200 import tables as T_, tripwire
    t = tables_v
    n = 2048
    rng = np.random.default_rng(9)
    lo_r = np.array([t.jnt_range[k][0] for k in range(6)]); hi_r = np.array([t.jnt_range[k][1] for k in range(6)])
    q = rng.uniform(np.maximum(lo_r, -1.6), np.minimum(hi_r, 1.6), (n, 6))
    res = {}
    for dtype in ("float64", "float32"):
        env = _vec(t, n, dtype=dtype, hulls=None)
        env.set_state(q, np.zeros((n, 6)), np.zeros((n, 6)))
        env.clear_flags()
        env.step_soa(torch.zeros((5, n), dtype=getattr(torch, dtype), device="cuda"), 1)
        res[dtype] = (env.flags().cpu().numpy() & T_.FLAG_TRIP_SELF) != 0
    assert res["float64"].sum() > 20
    assert (res["float64"] != res["float32"]).mean() < 0.003
    env = _vec(t, n, hulls=None)
    env.set_state(q, np.zeros((n, 6)), np.zeros((n, 6)))
    env.clear_flags()
    env.step_soa(torch.zeros((5, n), dtype=torch.float64, device="cuda"), 1)
    env.set_state(np.zeros((n, 6)), np.zeros((n, 6)), np.zeros((n, 6)))      # back to the rest pose, flags kept
    env.step_soa(torch.zeros((5, n), dtype=torch.float64, device="cuda"), 1)
    assert np.array_equal((env.flags().cpu().numpy() & T_.FLAG_TRIP_SELF) != 0, res["float64"])
    # SO101_OPT_SELF_TEST = 1: no box-box test, every pose outside the fast-accept joint box is flagged (round-1 behaviour)
    env = _vec(t, n, hulls=None)
    env.set_option(T_.OPT_SELF_TEST, 1)
    env.set_state(q, np.zeros((n, 6)), np.zeros((n, 6)))
    env.clear_flags()
    env.step_soa(torch.zeros((5, n), dtype=torch.float64, device="cuda"), 1)
    lo = np.array([t.trip_qbox[k][0] for k in range(6)]); hi = np.array([t.trip_qbox[k][1] for k in range(6)])
    assert np.array_equal((env.flags().cpu().numpy() & T_.FLAG_TRIP_SELF) != 0, ((q < lo) | (q > hi)).any(axis=1))


@pytest.mark.parametrize("family", ["onewarp", "team"])
def test_self_collision_flags_of_a_free_running_rollout(tables_v, family):
    """The test runs only when an env has used up its separation budget (joint travel since the last test); skipping it must not
    lose a flag.  Constant joint velocities that fold half of the arms onto themselves; the numpy box-box test at every recorded
    pose of every env: an env whose boxes overlap at a recorded pose is flagged, and flags without a recorded overlap (a graze
    between two rows) stay rare."""
    from lerobot_mujoco_sim2real_b200 import tables as T_, tripwire
    t = tables_v
    n, Tn = 512, 150
    g = torch.Generator().manual_seed(4)
    rate = (torch.rand((5, n), generator=g, dtype=torch.float64) - 0.3) * 0.7        # mostly towards the folded poses
    rate[:, : n // 8] *= 0.2                                                         # some arms barely leave the joint box
    U = rate[None].repeat(Tn + 1, 1, 1).cuda().contiguous()
    env = _vec(t, n, hulls=None)
    env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_ONEWARP if family == "onewarp" else T_.FAMILY_TEAM)
    env.reset(options={"initial_state": torch.zeros((n, 10), dtype=torch.float64)})
    rows = env.rollout(Tn, "tensor", u=U, flags=T_.ROLL_NO_RESET).cpu().numpy()
    fl = (env.flags().cpu().numpy() & T_.FLAG_TRIP_SELF) != 0
    lo = np.array([t.trip_qbox[k][0] for k in range(5)]); hi = np.array([t.trip_qbox[k][1] for k in range(5)])
    qs = rows[:, :, 8:13]
    out = ((qs < lo) | (qs > hi)).any(axis=2)
    hit = np.zeros(n, bool)
    for i in np.nonzero(out.any(axis=1))[0]:
        for k in np.nonzero(out[i])[0][::2]:
            if tripwire.self_overlap_numpy(t, np.append(qs[i, k], 0.0)):
                hit[i] = True
                break
    assert out.any(axis=1).mean() > 0.5 and hit.sum() > 30   # many envs leave the fast-accept box, some fold onto themselves
    assert not (hit & ~fl).any()                             # none that overlaps at a recorded pose is missed
    assert (fl & ~hit).sum() <= 0.1 * hit.sum() + 3          # a flag without a recorded overlap: a graze between two rows


def test_rollout_host_single_abi_call(tables_v):
    """so101_batch_rollout_host (host controls + initial angles in, rows out) == the device-tensor path."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    n, T = 777, 4                              # not a multiple of the warp / block size
    g = torch.Generator().manual_seed(3)
    U = (torch.rand((T + 1, 5, n), generator=g, dtype=torch.float64) - 0.5).contiguous()
    q0 = torch.zeros((6, n), dtype=torch.float64); q0[:5] = (torch.rand((5, n), generator=g, dtype=torch.float64) - 0.5) * 0.6
    env = _vec(tables_v, n)
    rows_h = env.rollout_host(T, "tensor", u_host=U.pin_memory(), qpos0_host=q0.pin_memory())
    assert rows_h.device.type == "cpu" and rows_h.shape == (n, T + 1, 13)
    env2 = _vec(tables_v, n)
    env2.set_state(q0.t(), torch.zeros((n, 6)), torch.zeros((n, 6)))
    rows_d = env2.rollout(T, "tensor", u=U.cuda(), flags=T_.ROLL_NO_RESET)
    assert torch.equal(rows_h, rows_d.cpu())
    # generator-driven variant: random reset + Philox controls, pageable host memory
    out = torch.empty((n, T + 1, 13), dtype=torch.float64)
    env.rollout_host(T, "random", seed=5, out_host=out)
    assert torch.equal(out, env2.rollout(T, "random", seed=5).cpu())


@pytest.mark.parametrize("direct", [2, 1])
@pytest.mark.parametrize("kind", ["tensor", "random", "chirp"])
def test_rollout_host_pipeline_chunks_are_invisible(tables_v, monkeypatch, kind, direct):
    """rollout_host cuts long rollouts into time chunks that overlap upload / compute / download; the rows and the
    final state must not depend on the number of chunks (continuation launches regenerate u_t0 and skip row t0), nor on
    whether the rows are staged on the device and downloaded (direct = 2) or stored by the kernels straight into the pinned
    host buffer (direct = 1, the default for datasets below 128 MB)."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    n, T = 300, 23                             # 23 control steps: uneven chunks
    g = torch.Generator().manual_seed(9)
    U = (torch.rand((T + 1, 5, n), generator=g, dtype=torch.float64) - 0.5).contiguous().pin_memory()
    res = {}
    for chunks in ("1", "2", "4", "8", "12"):
        env = _vec(tables_v, n)
        env.set_option(T_.OPT_HOST_CHUNKS, int(chunks))
        env.set_option(T_.OPT_HOST_DIRECT, direct)
        out = torch.full((n, T + 1, 13), float("nan"), dtype=torch.float64).pin_memory()
        env.rollout_host(T, kind, seed=21, u_host=U if kind == "tensor" else None, out_host=out,
                         flags=T_.ROLL_GRAVCOMP_HOLD)
        q, v, w = env.get_state()
        res[chunks] = (out.clone(), q.clone(), v.clone(), w.clone())
        assert env.stats()["physics_steps"] == n * T * 10
        assert not torch.isnan(out).any()
    for chunks in ("2", "4", "8", "12"):
        for a, b in zip(res["1"], res[chunks]):
            assert torch.equal(a, b)
    dev = _vec(tables_v, n).rollout(T, kind, seed=21, u=U.cuda() if kind == "tensor" else None, flags=T_.ROLL_GRAVCOMP_HOLD)
    assert torch.equal(res["1"][0], dev.cpu())


@pytest.mark.parametrize("n", [1, 33, 1000])
def test_odd_batch_sizes_and_f32_rows(oracle_mod, tables_v, n):
    from lerobot_mujoco_sim2real_b200 import tables as T_
    O = oracle_mod
    env = _vec(tables_v, n)
    rows = env.rollout(3, "random", seed=77).cpu().numpy()
    ref, _, _ = O.rollout(tables_v, O.make_spec(kind=0, seed=77), n, 3, 10)
    np.testing.assert_array_equal(rows[:, :, :5], ref[:, :, :5])
    assert np.abs(rows - ref).max() <= 6e-8
    r32 = env.rollout(3, "random", seed=77, flags=T_.ROLL_ROWS_F32)
    assert r32.dtype == torch.float32
    np.testing.assert_array_equal(r32.cpu().numpy(), rows.astype(np.float32))


def test_gravcomp_hold_rollout_matches_oracle(oracle_mod, tables_v):
    """qfrc_applied <- qfrc_bias at the start of each control step, held over the sub-steps
    [REF Koopman_MPC.py:119]."""
    O = oracle_mod
    from lerobot_mujoco_sim2real_b200 import tables as T_
    n, T = 256, 4
    env = _vec(tables_v, n)
    rows = env.rollout(T, "random", seed=21, flags=T_.ROLL_GRAVCOMP_HOLD).cpu().numpy()
    ref, final, _ = O.rollout(tables_v, O.make_spec(kind=0, seed=21), n, T, 10, flags=T_.ROLL_GRAVCOMP_HOLD)
    assert np.abs(rows - ref).max() <= 6e-8
    np.testing.assert_allclose(env.get_state()[0].cpu().numpy(), final[:, :6], atol=1e-11)
    plain, _, _ = O.rollout(tables_v, O.make_spec(kind=0, seed=21), n, T, 10)
    assert np.abs(ref - plain).max() > 1e-4


def test_two_models_coexist_and_bad_arguments(tables_v, tables_p):
    import ctypes as C
    from lerobot_mujoco_sim2real_b200 import _lib
    a, b = _vec(tables_v, 64), _vec(tables_p, 64)
    ra1 = a.rollout(3, "random", seed=1); rb1 = b.rollout(3, "random", seed=1)
    ra2 = a.rollout(3, "random", seed=1); rb2 = b.rollout(3, "random", seed=1)
    assert torch.equal(ra1, ra2) and torch.equal(rb1, rb2) and not torch.equal(ra1, rb1)
    L = _lib.lib()
    spec = a.make_spec("tensor")            # tensor controls without a pointer
    assert L.so101_batch_rollout(a._h, C.byref(spec), 3, 10, None, 0, None) == -1
    assert b"spec->u" in L.so101_last_error()
    assert L.so101_batch_step(a._h, None, 7, 10, None, None) == -1
    with pytest.raises(_lib.So101Error):
        _lib.check(L.so101_batch_rollout(a._h, C.byref(a.make_spec("random")), -1, 10, None, 0, None))
    h = C.c_void_p()
    assert L.so101_batch_create(a.model._h, 0, 0, 0, None, C.byref(h)) == -1
    assert L.so101_batch_create(a.model._h, 8, 5, 0, None, C.byref(h)) == -1


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_reference_published_eval_metric_reproduced_on_gpu(tables_v, dtype):
    """The reference's only published number for data from this path: open-loop 200-step MAE of its
    trained Koopman model = 6.84e-3 [REF results/SOARM101/11_27/DKUC/best_scores.json:7].  The CUDA
    rollout (2000 x 200 'random' validation trajectories, gravity-compensation line active, see
    tests/test_oracle.py::test_koopman_long_horizon_fingerprint) reproduces it."""
    import os
    from lerobot_mujoco_sim2real_b200 import tables as T_
    W = {k: torch.as_tensor(v, dtype=torch.float64, device="cuda")
         for k, v in np.load(os.path.join(os.path.dirname(__file__), "golden", "koopman_dkuc.npz")).items()}

    def predict(x, u):
        h = x
        for i in range(5):
            h = h @ W[f"x_encode_net.linear_{i}.weight"].T + W[f"x_encode_net.linear_{i}.bias"]
            if i != 4:
                h = torch.relu(h)
        return (torch.cat([x, h], -1) @ W["lA.weight"].T + u @ W["lB.weight"].T) @ W["lC.weight"].T

    env = _vec(tables_v, 2000, dtype=dtype)
    rows = env.rollout(200, "random", seed=2024, flags=T_.ROLL_GRAVCOMP_HOLD)
    x, u = rows[:, :, 5:], rows[:, :, :5]
    x0, mae = x[:, 0], 0.0
    for i in range(200):
        x0 = predict(x0, u[:, i])
        mae += float((x0 - x[:, i + 1]).abs().mean())
    mae /= 200
    print(f"{dtype}: open-loop 200-step MAE of the reference's trained model on GPU data = {mae:.3e} (published 6.84e-3)")
    assert 5.5e-3 < mae < 8.5e-3


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_reference_training_loss_fingerprints_on_gpu(tables_v, dtype):
    """koopman_loss / pred_loss / recon_loss / train_total_loss of the reference's best epoch
    [REF results/SOARM101/11_27/DKUC/best_scores.json:3-8] on the reference's train-shaped job (50000 x 20 'random')
    generated by the CUDA rollout: four more statistics that have seen real MuJoCo output of this path."""
    import os
    from lerobot_mujoco_sim2real_b200 import tables as T_
    from oracle.koopman_oracle import check_training_loss_fingerprints
    W = {k: v.astype(np.float64) for k, v in np.load(os.path.join(os.path.dirname(__file__), "golden", "koopman_dkuc.npz")).items()}
    env = _vec(tables_v, 50000, dtype=dtype)
    rows = env.rollout(20, "random", seed=99, flags=T_.ROLL_GRAVCOMP_HOLD).cpu().numpy()
    check_training_loss_fingerprints(W, rows[:, :, 5:], rows[:, :, :5], f"CUDA {dtype}:")


# ------------------------------------------------------------------------------------------------
# small batches run the three-warp "team" kernels (so101_physics.cuh, SplitXch): same bits as the one-warp kernels,
# so the batch size (and with it the shard a rank holds) never changes a trajectory
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("n", [1, 300, 4096])
def test_split_team_is_bitwise_identical(tables_v, monkeypatch, dtype, n):
    out = {}
    from lerobot_mujoco_sim2real_b200 import tables as T_
    for split in ("0", "1"):
        env = _vec(tables_v, n, dtype)
        env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_TEAM if split == "1" else T_.FAMILY_ONEWARP)
        rows = env.rollout(100, "random", seed=7, env_offset=5)          # 1000 chaotic physics steps
        q, v, w = env.get_state()
        obs = env.step(torch.full((n, 5), 0.3, dtype=env.torch_dtype, device=env.device))   # k_step path
        U = torch.rand((7, 5, n), dtype=env.torch_dtype, device=env.device, generator=torch.Generator(env.device).manual_seed(3)) - 0.5
        X = env.shoot(np.concatenate([tables_v.key_qpos[:], np.zeros(12)]), U)               # k_shoot path
        out[split] = [t.clone() for t in (rows, q, v, w, obs if torch.is_tensor(obs) else obs[0], X, env.flags())]
        assert env.stats()["physics_steps"] == n * (100 + 1 + 7) * 10
    for a, b in zip(out["0"], out["1"]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n", [5, 4096, 20000])     # team kernels (one partial team / one team per SM), one-warp kernels
def test_zero_length_calls(tables_v, n):
    """Degenerate sizes the reference API allows: step with 0 sub-steps, rollout of 0 control steps, shoot with H = 0."""
    env = _vec(tables_v, n)
    env.reset(seed=3)
    q0, v0, _ = [t.clone() for t in env.get_state()]
    obs_f, _ = env.forward()
    obs_f = obs_f.clone()
    u = torch.full((5, n), 0.2, dtype=env.torch_dtype, device=env.device)
    obs0 = env.step_soa(u, 0).t().clone()
    assert torch.equal(obs0, obs_f)                                    # no physics: observation of the current state
    q1, v1, _ = env.get_state()
    assert torch.equal(q0, q1) and torch.equal(v0, v1)
    rows = env.rollout(0, "random", seed=9)                            # reset row only
    assert rows.shape == (n, 1, 13) and torch.isfinite(rows).all()
    rows_h = env.rollout_host(0, "random", seed=9)
    assert torch.equal(rows.cpu(), rows_h)
    s0 = np.concatenate([tables_v.key_qpos[:], np.zeros(12)])
    X = env.shoot(s0, torch.empty((0, 5, n), dtype=env.torch_dtype, device=env.device))
    assert X.shape == (n, 1, 8) and bool((X == X[0]).all())
    assert env.stats()["physics_steps"] == 0


@pytest.mark.parametrize("dtype,kind", [("float64", "random"), ("float64", "chirp"), ("float32", "random")])
def test_time_sliced_rollout_is_invisible(tables_v, dtype, kind):
    """Large batches whose env groups do not fill whole waves of resident blocks run the rollout time-sliced over a
    persistent grid (k_rollout_sliced: units = (group, time chunk), chunk c of a group possibly on another SM than chunk
    c-1).  Rows, final state, flags and solver statistics equal the plain launch bit for bit - also for a continuation
    (t0 > 0 through rollout_host's chunks), ragged sizes and with table contact."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    n, Tn = 40_003, 12
    res = []
    for mode in (2, 1):                                     # never / always
        env = _vec(tables_v, n, dtype=dtype)
        env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_ONEWARP)
        env.set_option(T_.OPT_SLICED, mode)
        rows = env.rollout(Tn, kind, seed=11)
        q, v, w = env.get_state()
        st = env.stats()
        res.append((rows.clone(), q.clone(), v.clone(), w.clone(), env.flags().clone(), st))
    a, b = res
    for x, y in zip(a[:5], b[:5]):
        assert torch.equal(x, y)
    assert a[5] == b[5] and a[5]["physics_steps"] == n * Tn * 10
    assert int((a[4] & T_.FLAG_CONTACT).ne(0).sum()) >= 0
    # automatic choice: this size (157 groups of 256 on 148 SMs in f64) is worth slicing; the result is the same again
    env = _vec(tables_v, n, dtype=dtype)
    rows = env.rollout(Tn, kind, seed=11)
    assert torch.equal(rows, a[0])


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_regrouped_rollout_is_invisible(tables_v, dtype):
    """Long rollouts of large batches are launched per chunk of control steps with the envs regrouped in between (the ones that
    touched the table share blocks afterwards: a block pays for a contact of any of its lanes).  Which lane or block an env sits
    in changes nothing: rows, final state, flags and statistics equal the plain launch bit for bit."""
    from lerobot_mujoco_sim2real_b200 import tables as T_
    n, Tn = 20_011, 60
    res = []
    for regroup in (2, 1):
        env = _vec(tables_v, n, dtype=dtype)
        env.set_option(T_.OPT_KERNEL_FAMILY, T_.FAMILY_ONEWARP)
        env.set_option(T_.OPT_REGROUP, regroup)
        if regroup == 2:
            env.set_option(T_.OPT_SLICED, 2)
        rows = env.rollout(Tn, "chirp", seed=7)
        q, v, w = env.get_state()
        res.append((rows.clone(), q.clone(), v.clone(), w.clone(), env.flags().clone(), env.stats()))
    a, b = res
    assert int((a[4] & T_.FLAG_CONTACT).ne(0).sum()) > 10          # the case has contacts to regroup
    for x, y in zip(a[:5], b[:5]):
        assert torch.equal(x, y)
    assert a[5] == b[5]
