"""CPU restatement of the reference's Cartesian-track inverse kinematics.

TEST INFRASTRUCTURE ONLY: the product (`lerobot-mujoco-sim2real_b200/`) never imports this module.

**Parity unpinned.**  The arithmetic of this path lives in two third-party packages that are neither vendored nor
version-pinned by the reference and are absent from this image: `dm_control` (`dm_control.utils.inverse_kinematics.
qpos_from_site_pose` and `nullspace_method`) and `mujoco` (`mj_fwdPosition`, `mj_jacSite`, `mju_mat2Quat`,
`mju_negQuat`, `mju_mulQuat`, `mju_quat2Vel`, `mj_integratePos`).  The reference holds no tests or golden vectors
for it.  This file restates their published algorithm and anchors on the reference's own call site:

    qpos_from_site_pose(physics, site_name, target_pos, target_quat, joint_names=<5 arm joints>, inplace=True,
                        max_steps=100, tol=1e-6, rot_weight=0.5, regularization_strength=1e-2)
                                                                       [REF control/TrajectoryGenerator.py:96-107]
    (dm_control defaults for the rest: regularization_threshold=0.1, max_update_norm=2.0, progress_thresh=20.0)

and the way-point loop around it [REF control/TrajectoryGenerator.py:180-210]: solves are chained (each starts from
the joint vector the previous one left), a failed solve repeats the previous answer and restores the joint vector it
started from, a failure at the first way-point raises.

Kinematics come from the C oracle's restatement of mj_kinematics / mj_comPos (so101_oracle.c) in MuJoCo's own
formulation (quaternion FK; Jacobian from `cdof` and `subtree_com` exactly as `mj_jac` forms it), which shares no
structure with the CUDA kernel's link-local chain.
"""
from __future__ import annotations

from typing import NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import oracle as O

MJ_MINVAL = 1e-15

ARM_DOFS = (0, 1, 2, 3, 4)   # joint_names of the reference call: the five arm joints [REF TrajectoryGenerator.py:56]


class IKResult(NamedTuple):
    qpos: np.ndarray
    err_norm: float
    steps: int
    success: bool
    cut_noise: bool = False     # diagnostic, not part of dm_control's IKResult: see `lstsq_kept_noise`


def lstsq_kept_noise(jac_joints: np.ndarray, delta: np.ndarray, update: np.ndarray) -> bool:
    """Did `np.linalg.lstsq(J'J, J'delta, rcond=-1)` keep a rounding-noise singular value?

    With a position-only target J is 3 x 5, so J'J (5 x 5) has rank 3 and two singular values that are pure rounding
    noise, ~1e-16 * s_max — right AT the rcond=-1 cut (DBL_EPSILON * s_max).  Usually LAPACK drops them and the update
    is the minimum-norm one, pinv(J) delta.  Every few hundred calls one lands above the cut (measured here: 4 of 601
    calls on the reference's Circle/idx=0 curve, s_4/s_1 = 1.1e-16 .. 1.5e-16) and the update acquires a null-space
    component of the size of the update itself (|x - pinv(J) delta| up to 7e-3 rad).  It does not move the site, so the
    solve still converges, but the joint angles from there on depend on rounding inside LAPACK — the reference's own
    answer is not reproducible across BLAS builds at those way-points.  The oracle reports the event so parity tests
    can say where bitwise-level agreement with ANY other implementation stops being meaningful."""
    clean = np.linalg.pinv(jac_joints) @ delta
    return bool(np.linalg.norm(update - clean) > 1e-10)


# ---- mju_* helpers (engine_util_spatial.c) ---------------------------------------------------------------------
def mju_mat2Quat(mat: np.ndarray) -> np.ndarray:
    m = np.asarray(mat, dtype=np.float64).reshape(9)
    q = np.zeros(4)
    if m[0] + m[4] + m[8] > 0:
        q[0] = 0.5 * np.sqrt(1 + m[0] + m[4] + m[8])
        q[1] = 0.25 * (m[7] - m[5]) / q[0]
        q[2] = 0.25 * (m[2] - m[6]) / q[0]
        q[3] = 0.25 * (m[3] - m[1]) / q[0]
    elif m[0] > m[4] and m[0] > m[8]:
        q[1] = 0.5 * np.sqrt(1 + m[0] - m[4] - m[8])
        q[0] = 0.25 * (m[7] - m[5]) / q[1]
        q[2] = 0.25 * (m[1] + m[3]) / q[1]
        q[3] = 0.25 * (m[2] + m[6]) / q[1]
    elif m[4] > m[8]:
        q[2] = 0.5 * np.sqrt(1 - m[0] + m[4] - m[8])
        q[0] = 0.25 * (m[2] - m[6]) / q[2]
        q[1] = 0.25 * (m[1] + m[3]) / q[2]
        q[3] = 0.25 * (m[5] + m[7]) / q[2]
    else:
        q[3] = 0.5 * np.sqrt(1 - m[0] - m[4] + m[8])
        q[0] = 0.25 * (m[3] - m[1]) / q[3]
        q[1] = 0.25 * (m[2] + m[6]) / q[3]
        q[2] = 0.25 * (m[5] + m[7]) / q[3]
    n = np.sqrt(np.dot(q, q))
    if n < MJ_MINVAL:
        return np.array([1.0, 0.0, 0.0, 0.0])
    return q / n


def mju_negQuat(q: np.ndarray) -> np.ndarray:
    return np.array([q[0], -q[1], -q[2], -q[3]])


def mju_mulQuat(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    return np.array([
        a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
        a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
        a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
        a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def mju_quat2Vel(q: np.ndarray, dt: float) -> np.ndarray:
    axis = np.array(q[1:4], dtype=np.float64)
    sin_a_2 = np.sqrt(np.dot(axis, axis))
    if sin_a_2 < MJ_MINVAL:
        axis = np.array([1.0, 0.0, 0.0])
    else:
        axis = axis / sin_a_2
    speed = 2 * np.arctan2(sin_a_2, q[0])
    if speed > np.pi:
        speed -= 2 * np.pi
    return axis * (speed / dt)


def _quat2mat(q: np.ndarray) -> np.ndarray:
    """mju_quat2Mat."""
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def nullspace_method(jac_joints: np.ndarray, delta: np.ndarray, regularization_strength: float = 0.0) -> np.ndarray:
    """dm_control.utils.inverse_kinematics.nullspace_method."""
    hess_approx = jac_joints.T.dot(jac_joints)
    joint_delta = jac_joints.T.dot(delta)
    if regularization_strength > 0:
        hess_approx += np.eye(hess_approx.shape[0]) * regularization_strength
        return np.linalg.solve(hess_approx, joint_delta)
    return np.linalg.lstsq(hess_approx, joint_delta, rcond=-1)[0]


class Physics:
    """The slice of dm_control's Physics the IK touches: qpos, mj_fwdPosition, the site pose, mj_jacSite."""

    def __init__(self, tables):
        self.t = tables
        self.o = O.Oracle(tables)
        self.o.reset()
        self.qpos = np.array(self.o.arr("qpos"), dtype=np.float64)
        sq = np.array(tables.site_quat[:], dtype=np.float64)
        self._site_local = _quat2mat(sq / np.sqrt(np.dot(sq, sq)))
        self.fwd_position()

    def fwd_position(self) -> None:
        """mj_fwdPosition (kinematics + comPos are what the IK reads; the oracle's forward computes them)."""
        self.o.set("qpos", self.qpos)
        self.o.forward()
        b = self.t.site_body
        self.site_xpos = np.array(self.o.arr("site_xpos"), dtype=np.float64)
        self.site_xmat = np.array(self.o.arr("xmat")[b]).reshape(3, 3) @ self._site_local

    def jac_site(self) -> Tuple[np.ndarray, np.ndarray]:
        """mj_jacSite -> mj_jac(point = site_xpos, body = site body) (engine_core_smooth.c)."""
        nv = self.t.nv
        jacp, jacr = np.zeros((3, nv)), np.zeros((3, nv))
        body = self.t.site_body
        root = int(self.o.arr("body_root")[body])
        offset = self.site_xpos - np.array(self.o.arr("subtree_com")[root])
        cdof = np.array(self.o.arr("cdof"))
        # dofs on the path from the site's body to the world
        a = body
        while a > 0:
            j = self.t.body_jnt[a]
            if j >= 0:
                jacr[:, j] = cdof[j][0:3]
                jacp[:, j] = cdof[j][3:6] + np.cross(cdof[j][0:3], offset)
            a = self.t.body_parent[a]
        return jacp, jacr


def qpos_from_site_pose(physics: Physics, target_pos: Optional[np.ndarray] = None,
                        target_quat: Optional[np.ndarray] = None, dof_indices: Sequence[int] = ARM_DOFS,
                        tol: float = 1e-14, rot_weight: float = 1.0, regularization_threshold: float = 0.1,
                        regularization_strength: float = 3e-2, max_update_norm: float = 2.0,
                        progress_thresh: float = 20.0, max_steps: int = 100) -> IKResult:
    """dm_control.utils.inverse_kinematics.qpos_from_site_pose with inplace=True (defaults are dm_control's)."""
    if target_pos is None and target_quat is None:
        raise ValueError("At least one of `target_pos` or `target_quat` must be specified.")
    nv = physics.t.nv
    dof_indices = list(dof_indices)
    update_nv = np.zeros(nv)
    physics.fwd_position()
    steps = 0
    success = False
    err_norm = 0.0
    cut_noise = False
    for steps in range(max_steps):
        err_norm = 0.0
        parts = []
        if target_pos is not None:
            err_pos = np.asarray(target_pos, dtype=np.float64) - physics.site_xpos
            err_norm += np.linalg.norm(err_pos)
            parts.append(err_pos)
        if target_quat is not None:
            site_xquat = mju_mat2Quat(physics.site_xmat)
            neg_site_xquat = mju_negQuat(site_xquat)
            err_rot_quat = mju_mulQuat(np.asarray(target_quat, dtype=np.float64), neg_site_xquat)
            err_rot = mju_quat2Vel(err_rot_quat, 1)
            err_norm += np.linalg.norm(err_rot) * rot_weight
            parts.append(err_rot)
        err = np.concatenate(parts)
        if err_norm < tol:
            success = True
            break
        jacp, jacr = physics.jac_site()
        rows = ([jacp] if target_pos is not None else []) + ([jacr] if target_quat is not None else [])
        jac_joints = np.concatenate(rows, axis=0)[:, dof_indices]
        reg_strength = regularization_strength if err_norm > regularization_threshold else 0.0
        update_joints = nullspace_method(jac_joints, err, regularization_strength=reg_strength)
        if reg_strength == 0.0:
            cut_noise = cut_noise or lstsq_kept_noise(jac_joints, err, update_joints)
        update_norm = np.linalg.norm(update_joints)
        with np.errstate(divide="ignore", invalid="ignore"):
            progress_criterion = err_norm / update_norm
        if progress_criterion > progress_thresh:
            break
        if update_norm > max_update_norm:
            update_joints = update_joints * (max_update_norm / update_norm)
        update_nv[:] = 0.0
        update_nv[dof_indices] = update_joints
        physics.qpos = physics.qpos + update_nv     # mj_integratePos on hinges
        physics.fwd_position()
    return IKResult(qpos=physics.qpos, err_norm=float(err_norm), steps=int(steps), success=bool(success),
                    cut_noise=cut_noise)


def track(tables, xyz: np.ndarray, target_quat: Optional[np.ndarray] = None, q0: Optional[np.ndarray] = None,
          max_steps: int = 100, tol: float = 1e-6, rot_weight: float = 0.5, regularization_strength: float = 1e-2,
          dof_indices: Sequence[int] = ARM_DOFS):
    """The way-point loop of CartesianTrajectoryGenerator.generate [REF control/TrajectoryGenerator.py:180-210] for one
    track `xyz [P,3]`.  Returns (q [P,nv], status [P] int32 = success | aborted<<1 | cut_noise<<2 | steps<<8, err_norm
    [P]).  A failure at the first way-point (where the reference raises RuntimeError) marks the whole track aborted;
    cut_noise (bit 2, oracle-only diagnostic) marks way-points where numpy's lstsq kept a rounding-noise singular value
    (`lstsq_kept_noise`)."""
    ph = Physics(tables)
    if q0 is not None:
        ph.qpos = np.array(q0, dtype=np.float64)
    P = len(xyz)
    nv = tables.nv
    q_out = np.zeros((P, nv))
    status = np.zeros(P, dtype=np.int32)
    errs = np.zeros(P)
    have_any = False
    aborted = False
    for i in range(P):
        last_successful_qpos = ph.qpos.copy()
        if aborted:
            q_out[i], status[i] = ph.qpos, 2
            continue
        res = qpos_from_site_pose(ph, target_pos=xyz[i], target_quat=target_quat, dof_indices=dof_indices,
                                  max_steps=max_steps, tol=tol, rot_weight=rot_weight,
                                  regularization_strength=regularization_strength)
        errs[i] = res.err_norm
        if res.success:
            have_any = True
        else:
            ph.qpos = last_successful_qpos
            if not have_any:
                aborted = True
        q_out[i] = ph.qpos
        status[i] = (1 if res.success else 0) | (2 if aborted else 0) | (4 if res.cut_noise else 0) | (res.steps << 8)
    return q_out, status, errs


def reference_curve(traj_name: str = "Fig8", idx: int = 1, time_horizon: float = 60, time_steps_per_sec: int = 5,
                    traj_scale: float = 0.5) -> Tuple[np.ndarray, np.ndarray]:
    """Way-points of CartesianTrajectoryGenerator.generate [REF control/TrajectoryGenerator.py:136-170].
    Returns (xyz [P,3], time_vector [P])."""
    steps = time_steps_per_sec * time_horizon
    time_vector = np.linspace(0, time_horizon, steps)
    t = 1.6 + 0.02 * np.linspace(0, time_horizon * 5, len(time_vector))
    one = np.ones(len(t))
    if traj_name == "Fig8":
        a = b = 0.2 * traj_scale
        u = 2 * a * np.sin(t) * np.cos(t) / (1 + np.sin(t) ** 2)
        y = b * np.cos(t) / (1 + np.sin(t) ** 2)
        xyz = np.stack([0.4 * one, y, 0.2 + u], 1) if idx == 1 else np.stack([0.3 + u, y, 0.2 * one], 1)
    elif traj_name == "Circle":
        r = 0.1
        if idx == 1:
            xyz = np.stack([0.4 * one, 0.0 + r * np.cos(t), 0.2 + r * np.sin(t)], 1)
        else:
            xyz = np.stack([0.3 + r * np.cos(t), 0.0 + r * np.sin(t), 0.2 * one], 1)
    else:
        raise ValueError(f"未知的轨迹名称: {traj_name}")
    return xyz, time_vector
