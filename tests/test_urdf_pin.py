"""Known answers from the reference's own URDF of the arm (SOARM101/SO101/so101_new_calib.urdf) against the MJCF-compiled
tables, the CPU oracle and the CUDA kernels.

The hot path loads the MJCF; the URDF is an independent description of the same mechanism in another convention that
the reference ships alongside.  tools/gen_urdf_golden.py derives, from the URDF alone, the end-effector position, the
link centres of mass and the joint-space mass matrix (no armature) at 25 joint vectors -> tests/golden/
urdf_kinematics.npz.  Both files print ~6 significant digits, so the stated tolerance is 5e-6 m / 2e-5 relative.
This pins the inputs and outputs of mj_kinematics / mj_comPos / mj_crb (geometry, masses, inertias) to a
reference-owned source; it does not pin the dynamics (actuators, friction rows, solver), which only MuJoCo could."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "urdf_kinematics.npz")
LINK_TO_BODY = {"base_link": "base", "shoulder_link": "shoulder", "upper_arm_link": "upper_arm",
                "lower_arm_link": "lower_arm", "wrist_link": "wrist", "gripper_link": "gripper",
                "moving_jaw_so101_v1_link": "moving_jaw_so101_v1"}
POS_TOL, M_TOL = 5e-6, 2e-5


def _bodies():
    from lerobot_mujoco_sim2real_b200 import tables as T
    doc = json.load(open(os.path.join(T.ASSET_DIR, T.BUILTIN_SCENES["scene_with_table_v.xml"])))
    return doc["meta"]["bodies"]


def test_compiled_tables_match_the_urdf(tables_v, tables_p):
    from lerobot_mujoco_sim2real_b200 import mjcf
    g = np.load(GOLD)
    bodies = _bodies()
    assert abs(sum(g["link_mass"]) - 0.632006) < 1e-9
    for t in (tables_v, tables_p):
        # joint ranges (the limit rows' geometry): MJCF and URDF print them to 6 digits
        rng_mjcf = np.array([list(t.jnt_range[k]) for k in range(6)])
        assert all(t.jnt_limited[k] for k in range(6)) and np.abs(rng_mjcf - g["joint_limits"]).max() < 5e-6
        # servo constants as the reference's stand-alone SO101/joints_properties.xml states them for class sts3215
        # (damping 0.60, frictionloss 0.052, armature 0.028): a second reference-owned source for the dof tables
        for k in range(6):
            assert (t.dof_damping[k], t.dof_frictionloss[k], t.dof_armature[k]) == (0.6, 0.052, 0.028)
        for n, m in zip(g["link_names"], g["link_mass"]):
            assert t.body_mass[bodies.index(LINK_TO_BODY[str(n)])] == pytest.approx(m, abs=1e-12)
        for i, q in enumerate(g["q"]):
            assert np.abs(mjcf.site_numpy(t, q) - g["ee"][i]).max() < POS_TOL
            xpos, xmat, _, _ = mjcf.fk_numpy(t, q)
            for k, n in enumerate(g["link_names"]):
                b = bodies.index(LINK_TO_BODY[str(n)])
                com = xpos[b] + xmat[b] @ np.array(t.body_ipos[b][:])
                assert np.abs(com - g["com"][i][k]).max() < POS_TOL
            M = mjcf.mass_matrix_numpy(t, q) - np.diag(np.array(t.dof_armature[:]))
            assert np.abs(M - g["M"][i]).max() < M_TOL * np.abs(g["M"][i]).max()


def test_oracle_kinematics_and_crb_match_the_urdf(tables_v, oracle_mod):
    """The C restatement of mj_kinematics / mj_comPos / mj_crb (quaternion FK, composite inertias, sparse qM)."""
    g = np.load(GOLD)
    o = oracle_mod.Oracle(tables_v)
    arm = np.diag(np.array(tables_v.dof_armature[:]))
    for i, q in enumerate(g["q"]):
        o.reset()
        o.set("qpos", q)
        o.forward()
        assert np.abs(np.array(o.arr("site_xpos")) - g["ee"][i]).max() < POS_TOL
        assert np.abs(o.full_M() - arm - g["M"][i]).max() < M_TOL * np.abs(g["M"][i]).max()
        # whole-arm centre of mass (subtree_com of the world) against the URDF's mass-weighted link COMs
        com = (g["com"][i] * g["link_mass"][:, None]).sum(0) / g["link_mass"].sum()
        assert np.abs(np.array(o.arr("subtree_com")[0]) - com).max() < POS_TOL


@pytest.mark.gpu
def test_cuda_kinematics_and_ik_match_the_urdf(tables_v):
    """The kernels' link-local FK (observation site) and the IK built on it, against the URDF's end-effector positions."""
    import torch
    from lerobot_mujoco_sim2real_b200.TrajectoryGenerator import CartesianTrajectoryGenerator
    from lerobot_mujoco_sim2real_b200.vec_env import SOARM101VecEnv
    g = np.load(GOLD)
    n = len(g["q"])
    env = SOARM101VecEnv(tables=tables_v, num_envs=n, dtype="float64")
    q = torch.as_tensor(g["q"]).cuda()
    env.set_state(q, torch.zeros_like(q), torch.zeros_like(q))
    obs = env.forward()[0].cpu().numpy().astype(np.float64)
    assert np.abs(obs[:, :3] - g["ee"]).max() < POS_TOL + 1e-7           # float32 observation
    # IK to the URDF's end-effector positions, started nearby: converges, and the URDF's own FK of the kernel's answer
    # is what the kernel was asked for (re-evaluated here through the MJCF tables, which the test above ties to the URDF)
    gen = CartesianTrajectoryGenerator(tables=tables_v)
    q0 = g["q"].copy(); q0[:, :5] += 0.05
    qs, st, err = gen.solve_tracks(g["ee"][:, None, :], q0=q0, return_err=True)
    assert (st[:, 0] & 1).all() and float(err.max()) < 1e-6
    env.set_state(qs[:, 0].contiguous(), torch.zeros_like(q), torch.zeros_like(q))
    obs2 = env.forward()[0].cpu().numpy().astype(np.float64)
    assert np.abs(obs2[:, :3] - g["ee"]).max() < 2e-6
