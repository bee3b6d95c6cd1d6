#!/usr/bin/env python
"""Golden vectors for the Cartesian-track inverse kinematics: inputs and the CPU oracle's outputs.

    python tools/gen_ik_golden.py        ->  tests/golden/ik_tracks.npz

These are outputs of oracle/ik_oracle.py (the restatement of dm_control's qpos_from_site_pose on the C oracle's
kinematics), NOT of dm_control/MuJoCo themselves — neither is installable here (parity unpinned, see the oracle's
header).  They pin the oracle against edits and let the GPU box compare the CUDA kernel without re-running it.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from lerobot_mujoco_sim2real_b200 import builtin_tables  # noqa: E402
from oracle import ik_oracle as IK  # noqa: E402


def pose_of(ph: "IK.Physics", q: np.ndarray):
    ph.qpos = np.array(q, dtype=np.float64)
    ph.fwd_position()
    return ph.site_xpos.copy(), IK.mju_mat2Quat(ph.site_xmat)


def make_cases(tables, seed: int = 7):
    rng = np.random.default_rng(seed)
    ph = IK.Physics(tables)
    cases = []
    # (1) the reference's four curves, position only, from qpos = 0: what Koopman_MPC.py runs (target_quat=None)
    for name in ("Fig8", "Circle"):
        for idx in (1, 0):
            xyz, _ = IK.reference_curve(name, idx)
            cases.append(dict(kind=f"{name}/idx{idx}/pos", xyz=xyz, quat=None, q0=np.zeros(6)))
    # (2) the generator's default target_orientation = identity: unreachable for the 5-dof arm -> first point fails
    cases.append(dict(kind="Fig8/idx1/identity-quat", xyz=IK.reference_curve("Fig8", 1)[0][:20],
                      quat=np.array([1.0, 0, 0, 0]), q0=np.zeros(6)))
    # (3) reachable pose tracks: targets = site pose along a joint-space line, start offset from the first target
    for k in range(10):
        qa = rng.uniform(-0.8, 0.8, 6)
        qa[5] = 0.0
        # joints 1-3 have parallel axes: moving them with a constant sum keeps the site orientation, so the pose
        # targets of the even tracks stay on the arm's 5-dimensional pose manifold; the odd tracks also move the
        # pan joint, their orientation target becomes unreachable and way-points fail (repeat-previous / restore)
        a, b = rng.uniform(-0.5, 0.5, 2)
        d = np.array([0.3 * (k % 2), a, b, -a - b, 0.0, 0.0])
        P = 16
        pts = np.zeros((P, 3))
        quat = pose_of(ph, qa)[1]
        for i in range(P):
            pts[i] = pose_of(ph, qa + d * (i / (P - 1)))[0]
        q0 = qa + np.concatenate([rng.uniform(-0.25, 0.25, 5), [0.0]])
        cases.append(dict(kind=f"pose/{k}", xyz=pts, quat=quat, q0=q0))
    # (4) position-only tracks far from the start (regularised phase, update clipping), some out of reach
    for k in range(10):
        P = 16
        c = np.array([0.25, 0.0, 0.15]) + rng.uniform(-0.1, 0.1, 3)
        r = 0.05 + 0.25 * (k / 9.0) ** 2          # the last ones leave the workspace: failures after a good start
        ang = np.linspace(0, 2 * np.pi, P)
        pts = c + np.stack([r * np.cos(ang) * (k % 3 != 0), r * np.sin(ang), r * np.cos(ang) * (k % 3 == 0)], 1)
        cases.append(dict(kind=f"pos/{k}", xyz=pts, quat=None, q0=np.concatenate([rng.uniform(-1.0, 1.0, 5), [0.0]])))
    return cases


def main() -> None:
    tables = builtin_tables("scene_with_table_v.xml")
    out = {}
    cases = make_cases(tables)
    kinds = []
    for i, c in enumerate(cases):
        q, status, err = IK.track(tables, c["xyz"], c["quat"], c["q0"])
        out[f"xyz_{i}"] = c["xyz"]
        out[f"quat_{i}"] = c["quat"] if c["quat"] is not None else np.zeros(0)
        out[f"q0_{i}"] = c["q0"]
        out[f"q_{i}"], out[f"status_{i}"], out[f"err_{i}"] = q, status, err
        kinds.append(c["kind"])
        print(f"{c['kind']:28s} P={len(q):3d} success {np.mean(status & 1):.2f} aborted {bool((status & 2).any())} "
              f"steps mean {np.mean(status >> 8):.2f} max {np.max(status >> 8)}")
    out["kinds"] = np.array(kinds)
    path = os.path.join(ROOT, "tests", "golden", "ik_tracks.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
