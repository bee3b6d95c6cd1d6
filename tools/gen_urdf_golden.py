#!/usr/bin/env python
"""Kinematic / inertial known answers from the reference's URDF of the same arm -> tests/golden/urdf_kinematics.npz

The reference ships the SO-ARM101 twice: as MJCF (`SOARM101/SO101/so101_new_calib_v.xml`, what the hot path loads) and
as URDF (`SOARM101/SO101/so101_new_calib.urdf`, not loaded by the hot path).  The URDF is an independent description
of the same mechanism in another convention (joint frames as xyz/rpy, full inertia tensors about the link COM), so
quantities derived from it pin the MJCF compiler and the model tables against a reference-owned source:
end-effector position, link centres of mass, total mass and the joint-space mass matrix (without armature) at random
joint vectors.  This script needs /root/reference (build container only); the .npz it writes travels with the repo.

    python tools/gen_urdf_golden.py [--urdf /root/reference/SOARM101/SO101/so101_new_calib.urdf]
"""
import argparse
import os
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JOINTS = ["shoulder_pan", "shoulder_lift", "elbow_flex", "wrist_flex", "wrist_roll", "gripper"]


def rpy(r, p, y):
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    return Rz @ Ry @ Rx


def origin(elem):
    o = elem.find("origin") if elem is not None else None
    xyz = np.array([float(v) for v in (o.get("xyz", "0 0 0") if o is not None else "0 0 0").split()])
    ang = [float(v) for v in (o.get("rpy", "0 0 0") if o is not None else "0 0 0").split()]
    return xyz, rpy(*ang)


def axis_rot(a, q):
    a = a / np.linalg.norm(a)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(q) * K + (1 - np.cos(q)) * K @ K


class Urdf:
    def __init__(self, path):
        root = ET.parse(path).getroot()
        self.links = {}
        for L in root.findall("link"):
            I = L.find("inertial")
            if I is None:
                self.links[L.get("name")] = None
                continue
            c, Ri = origin(I)
            m = float(I.find("mass").get("value"))
            t = I.find("inertia")
            g = lambda k: float(t.get(k))
            Ic = np.array([[g("ixx"), g("ixy"), g("ixz")], [g("ixy"), g("iyy"), g("iyz")], [g("ixz"), g("iyz"), g("izz")]])
            self.links[L.get("name")] = (m, c, Ri @ Ic @ Ri.T)       # mass, COM in link frame, inertia in link axes
        self.joints = {}
        for J in root.findall("joint"):
            if J.find("parent") is None:
                continue
            xyz, R = origin(J)
            ax = J.find("axis")
            a = np.array([float(v) for v in ax.get("xyz").split()]) if ax is not None else np.zeros(3)
            lim = J.find("limit")
            self.joints[J.get("name")] = dict(type=J.get("type"), parent=J.find("parent").get("link"),
                                              child=J.find("child").get("link"), xyz=xyz, R=R, axis=a,
                                              limit=(float(lim.get("lower")), float(lim.get("upper"))) if lim is not None else None)
        children = {j["child"] for j in self.joints.values()}
        self.root = [n for n in self.links if n not in children][0]

    def fk(self, q):
        """World pose (R, p) of every link at joint vector q (ordered as JOINTS)."""
        pose = {self.root: (np.eye(3), np.zeros(3))}
        todo = list(self.joints.items())
        while todo:
            rest = []
            for name, j in todo:
                if j["parent"] not in pose:
                    rest.append((name, j)); continue
                Rp, pp = pose[j["parent"]]
                R = Rp @ j["R"]
                p = pp + Rp @ j["xyz"]
                if j["type"] in ("revolute", "continuous"):
                    R = R @ axis_rot(j["axis"], q[JOINTS.index(name)])
                pose[j["child"]] = (R, p)
            todo = rest
        return pose

    def quantities(self, q):
        pose = self.fk(q)
        ee = pose["gripper_frame_link"][1]
        coms, masses = {}, {}
        for n, v in self.links.items():
            if v is None or v[0] < 1e-6:
                continue
            R, p = pose[n]
            coms[n], masses[n] = p + R @ v[1], v[0]
        # joint-space mass matrix by the Jacobian sum  M = sum_b m Jv'Jv + Jw' I_world Jw
        axes, anchors, below = [], [], []
        for jn in JOINTS:
            j = self.joints[jn]
            R, p = pose[j["child"]]
            axes.append(R @ (j["axis"] / np.linalg.norm(j["axis"]))); anchors.append(p)
        parent_of = {j["child"]: (jn, j["parent"]) for jn, j in self.joints.items()}
        M = np.zeros((6, 6))
        for n in coms:
            chain = []
            a = n
            while a in parent_of:
                jn, a2 = parent_of[a]
                if jn in JOINTS:
                    chain.append(JOINTS.index(jn))
                a = a2
            Jv, Jw = np.zeros((3, 6)), np.zeros((3, 6))
            for k in chain:
                Jw[:, k] = axes[k]
                Jv[:, k] = np.cross(axes[k], coms[n] - anchors[k])
            R = pose[n][0]
            Iw = R @ self.links[n][2] @ R.T
            M += masses[n] * Jv.T @ Jv + Jw.T @ Iw @ Jw
        return ee, coms, masses, M


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--urdf", default="/root/reference/SOARM101/SO101/so101_new_calib.urdf")
    args = ap.parse_args()
    u = Urdf(args.urdf)
    rng = np.random.default_rng(2024)
    Q = np.concatenate([np.zeros((1, 6)), rng.uniform(-1.0, 1.0, (24, 6))])
    Q[:, 5] = np.clip(Q[:, 5], -0.17, 1.0)
    names = None
    ee, com, Ms = [], [], []
    for q in Q:
        e, coms, masses, M = u.quantities(q)
        names = sorted(coms)
        ee.append(e); com.append([coms[n] for n in names]); Ms.append(M)
    out = dict(q=Q, ee=np.array(ee), com=np.array(com), M=np.array(Ms), link_names=np.array(names),
               link_mass=np.array([masses[n] for n in names]),
               joint_limits=np.array([u.joints[j]["limit"] for j in JOINTS]), source=np.array("SOARM101/SO101/so101_new_calib.urdf"))
    path = os.path.join(ROOT, "tests", "golden", "urdf_kinematics.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "links", names, "total mass", sum(masses.values()))
    print("ee(q=0) =", ee[0])


if __name__ == "__main__":
    main()
