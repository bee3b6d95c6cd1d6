"""Host MJCF compiler: anchors of SURVEY.md Appendix B, default-class semantics, committed tables."""
import os
import textwrap

import numpy as np
import pytest

from lerobot_mujoco_sim2real_b200 import builtin_tables, mjcf, tables as T

REF_SO101 = "/root/reference/SOARM101/SO101"


def test_builtin_tables_anchor_values(tables_v):
    t = tables_v
    assert (t.nbody, t.nv, t.nu) == (8, 6, 6)
    assert t.timestep == 0.002 and t.iterations == 100 and t.ls_iterations == 50
    assert t.tolerance == 1e-8 and t.ls_tolerance == 0.01
    assert list(t.gravity) == [0.0, 0.0, -9.81]
    assert abs(sum(t.body_mass[:]) - 0.632006) < 1e-9
    # [PROBE] anchors of SURVEY.md Appendix B
    np.testing.assert_allclose(mjcf.site_numpy(t, np.zeros(6)), [0.391362, -0.000011, 0.226469], atol=1e-6)
    np.testing.assert_allclose(mjcf.site_numpy(t, np.array(t.key_qpos[:])), [0.224218, 0.007517, 0.051111], atol=1e-6)
    M0 = mjcf.mass_matrix_numpy(t, np.zeros(6))
    np.testing.assert_allclose(np.diag(M0), [0.043286, 0.044310, 0.036609, 0.028997, 0.028041, 0.028016], atol=1e-6)
    np.testing.assert_allclose(t.dof_invweight0[:], [23.1025, 24.1765, 29.2483, 34.8568, 35.6627, 35.6943], atol=1e-4)
    np.testing.assert_allclose(mjcf.gravity_bias_numpy(t, np.zeros(6)),
                               [0, -0.538595, -0.452666, -0.117051, -0.000037, 0.003530], atol=1e-6)
    np.testing.assert_allclose(mjcf.gravity_bias_numpy(t, np.array(t.key_qpos[:])),
                               [0, -0.029734, -0.439150, -0.103442, -0.000071, 0.002996], atol=1e-6)
    assert abs(t.meaninertia - np.mean(np.diag(M0))) < 1e-15
    np.testing.assert_allclose(t.key_qpos[:], [-0.04137, -1.68611, 1.69453, 0.5, 0.01833, 0])


def test_scene_a_velocity_servo_inherits_kp(tables_v):
    """F2: <velocity> without kv inherits gainprm[0]=50 from the class's <position kp=50>."""
    t = tables_v
    for i in range(6):
        assert t.act_gain[i] == 50.0
        assert list(t.act_bias[i]) == [0.0, 0.0, -50.0]
        assert list(t.act_ctrlrange[i]) == [-2.0, 2.0] and list(t.act_forcerange[i]) == [-3.5, 3.5]
        assert t.act_ctrllimited[i] == 1 and t.act_forcelimited[i] == 1
        assert (t.dof_damping[i], t.dof_frictionloss[i], t.dof_armature[i]) == (0.60, 0.052, 0.028)
        assert t.jnt_limited[i] == 1


def test_scene_b_position_servo_dampratio(tables_p, tables_v):
    t = tables_p
    kv = [-t.act_bias[i][2] for i in range(6)]
    expect = [2.0 * np.sqrt(50.0 * t.dof_M0[i]) for i in range(6)]
    np.testing.assert_allclose(kv, expect, rtol=1e-14)
    np.testing.assert_allclose(kv, [2.94, 2.98, 2.71, 2.41, 2.37, 2.37], atol=0.01)
    for i in range(6):
        assert t.act_gain[i] == 50.0 and t.act_bias[i][1] == -50.0
        assert list(t.act_forcerange[i]) == [-33.5, 33.5]
    assert list(t.act_ctrlrange[0]) == [-1.91986, 1.91986]
    np.testing.assert_allclose(t.key_ctrl[:], t.key_qpos[:])
    # kinematics / inertia identical to scene A
    assert bytes(bytearray(t.body_pos)) == bytes(bytearray(tables_v.body_pos))
    assert bytes(bytearray(t.body_inertia)) == bytes(bytearray(tables_v.body_inertia))


def test_json_round_trip_is_bit_exact(tables_v, tmp_path):
    p = tmp_path / "t.json"
    T.save_tables(tables_v, str(p))
    assert bytes(T.load_tables(str(p))) == bytes(tables_v)


@pytest.mark.reference
@pytest.mark.parametrize("scene", sorted(T.BUILTIN_SCENES))
def test_committed_tables_reproduce_from_reference_xml(scene):
    cm = mjcf.compile_mjcf(os.path.join(REF_SO101, scene))
    try:
        from lerobot_mujoco_sim2real_b200 import tripwire
        tripwire.fill_tripwire(cm)
    except ImportError:
        pass
    assert bytes(cm.tables) == bytes(builtin_tables(scene))
    assert cm.joint_names[:5] == ["shoulder_pan", "shoulder_lift", "elbow_flex", "wrist_flex", "wrist_roll"]
    assert cm.site_names == ["baseframe", "gripperframe"] and cm.tables.site_body == 6
    ncoll = sum(1 for g in cm.geoms if g.contype or g.conaffinity)
    assert ncoll == 15   # 13 collision meshes + table box + floor plane


CHAIN = """
<mujoco>
  <compiler angle="{angle}" autolimits="true"/>
  <option timestep="0.001" gravity="0 0 -9.81"/>
  <default>
    <default class="a">
      <joint damping="0.5" armature="0.01"/>
      <position kp="20" kv="2"/>
      <default class="b"><joint frictionloss="0.1"/></default>
    </default>
  </default>
  <worldbody>
    <body name="base" pos="0 0 0.1" childclass="a">
      <inertial pos="0 0 0" mass="1" diaginertia="0.01 0.01 0.01"/>
      {links}
    </body>
  </worldbody>
  <actuator>{acts}</actuator>
</mujoco>
"""


def _chain_xml(angle="radian", rng="-1 1", act="position"):
    links, close = "", ""
    for k in range(6):
        cls = ' class="b"' if k % 2 else ""
        links += (f'<body name="l{k}" pos="0.1 0 0" quat="0.7071068 0.7071068 0 0">'
                  f'<joint name="j{k}" axis="0 0 1" range="{rng}"{cls}/>'
                  f'<inertial pos="0.05 0 0" mass="0.2" diaginertia="0.001 0.002 0.003"/>')
        if k == 5:
            links += '<site name="gripperframe" pos="0.1 0 0"/>'
        close += "</body>"
    acts = "".join(f'<{act} name="a{k}" joint="j{k}" class="a" ctrlrange="-1 1"/>' for k in range(6))
    return textwrap.dedent(CHAIN.format(angle=angle, links=links + close, acts=acts))


def test_default_class_nesting_and_autolimits(tmp_path):
    p = tmp_path / "chain.xml"
    p.write_text(_chain_xml())
    t = mjcf.compile_mjcf(str(p)).tables
    assert t.timestep == 0.001
    assert [t.dof_frictionloss[k] for k in range(6)] == [0, 0.1, 0, 0.1, 0, 0.1]   # class b only on odd joints
    assert all(t.dof_damping[k] == 0.5 and t.dof_armature[k] == 0.01 for k in range(6))
    assert all(t.jnt_limited[k] == 1 for k in range(6))
    assert all(t.act_gain[k] == 20 and list(t.act_bias[k]) == [0, -20, -2] for k in range(6))
    assert all(t.act_ctrllimited[k] == 1 and t.act_forcelimited[k] == 0 for k in range(6))


def test_degree_angles_and_velocity_inherits_kp(tmp_path):
    p = tmp_path / "deg.xml"
    p.write_text(_chain_xml(angle="degree", rng="-90 90", act="velocity"))
    t = mjcf.compile_mjcf(str(p)).tables
    np.testing.assert_allclose(t.jnt_range[0][:], [-np.pi / 2, np.pi / 2])
    assert all(t.act_gain[k] == 20 and list(t.act_bias[k]) == [0, 0, -20] for k in range(6))


def test_include_and_unsupported_features(tmp_path):
    inner = tmp_path / "inner.xml"
    inner.write_text(_chain_xml())
    outer = tmp_path / "outer.xml"
    outer.write_text('<mujoco><include file="inner.xml"/><option timestep="0.004"/></mujoco>')
    assert mjcf.compile_mjcf(str(outer)).tables.timestep == 0.004   # later <option> wins
    bad = tmp_path / "bad.xml"
    bad.write_text(_chain_xml().replace('<option timestep="0.001"', '<option integrator="RK4" timestep="0.001"'))
    with pytest.raises(mjcf.MjcfError):
        mjcf.compile_mjcf(str(bad))
    bad.write_text(_chain_xml().replace('axis="0 0 1"', 'type="slide" axis="0 0 1"', 1))
    with pytest.raises(mjcf.MjcfError):
        mjcf.compile_mjcf(str(bad))
    with pytest.raises(FileNotFoundError):
        mjcf.compile_mjcf(str(tmp_path / "missing.xml"))


def test_fullinertia_principal_axes():
    mom, q = mjcf._principal(np.array([8.3759e-05, 8.10403e-05, 2.39783e-05, 7.55525e-08, -1.16342e-06, 1.54663e-07]))
    assert mom[0] >= mom[1] >= mom[2] > 0
    R = mjcf.q_mat(q)
    full = R @ np.diag(mom) @ R.T
    np.testing.assert_allclose([full[0, 0], full[1, 1], full[2, 2], full[0, 1], full[0, 2], full[1, 2]],
                               [8.3759e-05, 8.10403e-05, 2.39783e-05, 7.55525e-08, -1.16342e-06, 1.54663e-07],
                               atol=1e-18)
    assert abs(np.linalg.det(R) - 1) < 1e-12


def test_tripwire_tables(tables_v):
    """Committed tripwire boxes: clear at rest, tripped at the survey's first-contact pose (F5)."""
    from lerobot_mujoco_sim2real_b200 import tripwire
    t = tables_v
    assert t.ntrip == 10 and abs(t.trip_plane_z + 0.0009) < 1e-12       # table top
    assert tripwire.table_clearance_numpy(t, np.zeros(6)) > 0.05
    rng = np.random.default_rng(0)
    for _ in range(200):                                                 # the reset box never trips
        q = np.zeros(6); q[:5] = rng.uniform(-0.3, 0.3, 5)
        assert tripwire.table_clearance_numpy(t, q) > 0
    q = np.zeros(6); q[1] = q[2] = q[3] = 0.33                           # first contact at 0.315
    assert tripwire.table_clearance_numpy(t, q) < 0
    for k in range(6):
        assert t.trip_qbox[k][0] < -0.15 and t.trip_qbox[k][1] > 0.5
