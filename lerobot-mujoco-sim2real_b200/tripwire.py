"""Contact tripwire tables (host, offline): conservative detection of "this env may be in contact".

The hot-path scene has contacts enabled (13 collision meshes vs a table box and a floor plane,
[REF SOARM101/SO101/so101_new_calib_v.xml:53-117, scene_with_table_v.xml:28,31])
but the CUDA kernels do not simulate contact.  Instead every env carries two flag bits:

  TRIP_TABLE  an oriented bounding box of a link's collision geometry dipped below the highest
              static support plane (the table top).  Box ⊇ convex hull ⊇ mesh, so a table contact
              can never be missed (no false negatives); it may fire slightly early.
  TRIP_SELF   the joint vector left a box |q_i| <= qbox_i inside which no pair of non-adjacent
              link hulls intersects (found by sampling hull-hull LP feasibility and shrunk by a
              safety factor: a heuristic fast accept, documented as such) AND the oriented boxes of
              two geoms on non-adjacent links overlap (separating-axis test, exact for the boxes;
              box ⊇ hull, so outside the joint box no self-collision is missed and the flag fires only
              slightly early: `self_overlap_numpy` is the restatement the kernels are tested against).

Parity and throughput claims are made over flag-free envs.  This module needs the STL meshes of the
reference tree, so it only runs in the build container (tools/gen_tables.py); the resulting numbers
travel inside the committed tables.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

from . import mjcf
from . import tables as T


def load_stl_vertices(path: str) -> np.ndarray:
    """Unique vertices of a binary STL."""
    with open(path, "rb") as f:
        f.seek(80)
        ntri = int(np.frombuffer(f.read(4), dtype="<u4")[0])
        rec = np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")])
        data = np.frombuffer(f.read(ntri * rec.itemsize), dtype=rec, count=ntri)
    v = data["v"].reshape(-1, 3).astype(np.float64)
    return np.unique(v, axis=0)


def geom_hulls(cm: mjcf.CompiledModel) -> List[Tuple[int, str, np.ndarray]]:
    """(body, mesh name, convex-hull vertices in the body frame) of every colliding geom on a moving body."""
    from scipy.spatial import ConvexHull
    out = []
    for g in cm.geoms:
        if g.body == 0 or not (g.contype or g.conaffinity):
            continue
        if g.type != "mesh" or g.mesh is None:
            raise mjcf.MjcfError(f"tripwire: colliding geom type {g.type!r} on a moving body is not handled")
        fn = cm.mesh_files.get(g.mesh)
        if fn is None:
            raise mjcf.MjcfError(f"tripwire: mesh {g.mesh!r} has no file")
        v = load_stl_vertices(os.path.join(cm.meshdir, fn))
        v = v[ConvexHull(v).vertices]
        # the body-frame placement of the mesh vertices is geom_pos + R(geom_quat) v (MuJoCo's
        # recentring of the mesh at its COM is folded back into the geom frame)
        out.append((g.body, g.mesh, g.pos + v @ mjcf.q_mat(g.quat).T))
    return out


def body_hulls(cm: mjcf.CompiledModel) -> Dict[int, np.ndarray]:
    """Convex-hull vertices (body frame) of all colliding geometry of every moving body."""
    from scipy.spatial import ConvexHull
    pts: Dict[int, List[np.ndarray]] = {}
    for b, _, v in geom_hulls(cm):
        pts.setdefault(b, []).append(v)
    out = {}
    for b, lst in pts.items():
        allv = np.concatenate(lst)
        out[b] = allv[ConvexHull(allv).vertices]
    return out


def obb(points: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Bounding box of a point set -> (center, axes rows, half sizes): the smaller of the PCA-aligned
    and the frame-aligned box."""
    c = points.mean(axis=0)
    _, _, vt = np.linalg.svd(points - c, full_matrices=False)
    if np.linalg.det(vt) < 0:
        vt[2] = -vt[2]
    proj = (points - c) @ vt.T
    lo, hi = proj.min(axis=0), proj.max(axis=0)
    pca = (c + ((lo + hi) / 2) @ vt, vt, (hi - lo) / 2)
    lo, hi = points.min(axis=0), points.max(axis=0)
    aligned = ((lo + hi) / 2, np.eye(3), (hi - lo) / 2)
    return min((pca, aligned), key=lambda b: float(np.prod(b[2])))


def support_plane_z(cm: mjcf.CompiledModel) -> float:
    """Highest top face among the static colliding world geoms (table box, floor plane)."""
    z = -np.inf
    for g in cm.geoms:
        if g.body != 0 or not (g.contype or g.conaffinity):
            continue
        if g.type == "box":
            z = max(z, g.pos[2] + g.size[2])
        elif g.type == "plane":
            z = max(z, g.pos[2])
    return float(z)


def _hulls_intersect(A: np.ndarray, B: np.ndarray) -> bool:
    """Do conv(A) and conv(B) intersect?  LP feasibility: sum a_i A_i = sum b_j B_j, convex weights."""
    from scipy.optimize import linprog
    na, nb = len(A), len(B)
    Aeq = np.zeros((5, na + nb))
    Aeq[:3, :na], Aeq[:3, na:] = A.T, -B.T
    Aeq[3, :na], Aeq[4, na:] = 1, 1
    beq = np.array([0, 0, 0, 1, 1.0])
    res = linprog(np.zeros(na + nb), A_eq=Aeq, b_eq=beq, bounds=(0, None), method="highs")
    return res.status == 0


def self_collision_box(cm: mjcf.CompiledModel, samples: int = 3000, seed: int = 0, safety: float = 0.9) -> np.ndarray:
    """Fast-accept joint box of the self-collision flag: the largest (of a fixed ladder) symmetric box in which no sampled
    pose has overlapping boxes of non-adjacent links (`self_overlap_numpy`, itself conservative for the hulls), x safety.
    Joints up to and including the first link that carries a colliding geom do not change the relative pose of any two
    such geoms: they are left free.  A heuristic (sampling), which is why poses outside it get the box-box test instead of
    a flag.  Needs the trip_* boxes of the tables (fill_tripwire calls it last)."""
    t = cm.tables
    rng = np.random.default_rng(seed)
    lo = np.array([t.jnt_range[k][0] for k in range(t.nv)])
    hi = np.array([t.jnt_range[k][1] for k in range(t.nv)])
    nb = t.ntrip + t.nself
    first_link = min(t.body_jnt[t.trip_body[i]] for i in range(nb)) if nb else t.nv
    free = np.arange(t.nv) <= first_link
    best = 0.0
    for scale in (0.3, 0.45, 0.6, 0.7, 0.8, 0.85, 0.9, 0.95, 1.0, 1.1, 1.2, 1.35, 1.5):
        blo, bhi = np.where(free, lo, np.maximum(lo, -scale)), np.where(free, hi, np.minimum(hi, scale))
        if any(self_overlap_numpy(t, rng.uniform(blo, bhi)) for _ in range(samples)):
            break
        best = scale
    box = np.zeros((t.nv, 2))
    box[:, 0] = np.where(free, -1e30, np.maximum(lo, -best * safety))
    box[:, 1] = np.where(free, 1e30, np.minimum(hi, best * safety))
    return box


def fill_tripwire(cm: mjcf.CompiledModel, with_self_box: bool = True) -> None:
    """Fill `cm.tables.trip_*` in place from the collision meshes (needs the reference assets).

    One box per colliding geom.  Geoms of a link whose height above the plane cannot change (first
    link of the chain turning about a vertical axis) are checked once here and get no box."""
    t = cm.tables
    plane = support_plane_z(cm)
    t.ntrip = 0
    t.nself = 0
    if not np.isfinite(plane):
        return
    xpos0, xmat0, _, axis0 = mjcf.fk_numpy(t, np.array(t.qpos0[:]))
    first = t.jnt_body[0]
    vertical_first = abs(abs(axis0[0][2]) - 1.0) < 1e-9
    for b, name, pts in geom_hulls(cm):
        if b == first and vertical_first:
            zmin = float((xpos0[b] + pts @ xmat0[b].T)[:, 2].min())
            if zmin <= plane:
                raise mjcf.MjcfError(f"geom {name!r} of the first link is in permanent contact with the support plane")
    boxes = []
    for b, name, pts in _contact_geoms(cm):
        c, ax, half = obb(pts)
        boxes.append((b, c, ax, half))
    # colliding geoms without a tripwire box (they cannot reach the plane) still take part in the self-collision test:
    # their boxes follow the tripwire boxes in the same arrays, [ntrip, ntrip + nself)
    have = {(b, name) for b, name, _ in _contact_geoms(cm)}
    extra = [(b,) + obb(pts) for b, name, pts in geom_hulls(cm) if (b, name) not in have and t.body_jnt[b] >= 0]
    if len(boxes) + len(extra) > T.MAXTRIP:
        raise mjcf.MjcfError("too many tripwire boxes")
    t.ntrip = len(boxes)
    t.nself = len(extra)
    for i, (b, c, ax, half) in enumerate(boxes + extra):
        t.trip_body[i] = b
        t.trip_center[i][:] = list(c)
        t.trip_axes[i][:] = list(ax.reshape(-1))
        t.trip_half[i][:] = list(half)
    t.trip_plane_z = plane
    fill_contact_params(cm)
    if with_self_box:
        box = self_collision_box(cm)
        for k in range(t.nv):
            t.trip_qbox[k][0], t.trip_qbox[k][1] = float(box[k, 0]), float(box[k, 1])


CUBE_RES = 16


def _contact_geoms(cm: mjcf.CompiledModel):
    """The colliding geoms that get a tripwire box / a hull, in table order (fill_tripwire's order), with hull
    vertices in the body frame.  Geoms of a first link that turns about a vertical axis are left out (their height
    above the plane never changes; fill_tripwire checks them once)."""
    t = cm.tables
    _, _, _, axis0 = mjcf.fk_numpy(t, np.array(t.qpos0[:]))
    first = t.jnt_body[0]
    vertical_first = abs(abs(axis0[0][2]) - 1.0) < 1e-9
    return [(b, name, pts) for b, name, pts in geom_hulls(cm)
            if t.body_jnt[b] >= 0 and not (b == first and vertical_first)]


def fill_contact_params(cm: mjcf.CompiledModel) -> None:
    """Contact parameters of (static support box, colliding geom) pairs as mj_contactParam mixes them; the CUDA
    contact path handles ONE parameter set, condim 3, and the top face of ONE axis-aligned box."""
    t = cm.tables
    t.con_enabled = 0
    boxes = [g for g in cm.geoms if g.body == 0 and (g.contype or g.conaffinity) and g.type == "box"]
    planes = [g for g in cm.geoms if g.body == 0 and (g.contype or g.conaffinity) and g.type == "plane"]
    if len(boxes) != 1 or t.ntrip == 0:
        return
    box = boxes[0]
    if abs(abs(mjcf.q_norm(box.quat)[0]) - 1.0) > 1e-12:
        raise mjcf.MjcfError("contact: the support box must be axis aligned")
    if any(p.pos[2] >= box.pos[2] + box.size[2] for p in planes):
        raise mjcf.MjcfError("contact: a static plane above the table top is not handled")
    movers = [g for g in cm.geoms if g.body != 0 and (g.contype or g.conaffinity)]
    params = set()
    for g in movers:
        if not ((g.contype & box.conaffinity) or (box.contype & g.conaffinity)):
            raise mjcf.MjcfError("contact: a colliding geom that does not collide with the table is not handled")
        if g.priority != box.priority:
            src = g if g.priority > box.priority else box
            fr, sr, si = src.friction, src.solref, src.solimp
        else:
            w = box.solmix / (box.solmix + g.solmix) if box.solmix + g.solmix > 0 else 0.5
            fr = np.maximum(box.friction, g.friction)
            sr, si = w * box.solref + (1 - w) * g.solref, w * box.solimp + (1 - w) * g.solimp
        params.add((tuple(fr), tuple(sr), tuple(si), max(box.margin, g.margin) - max(box.gap, g.gap),
                    max(box.condim, g.condim)))
    if len(params) != 1:
        raise mjcf.MjcfError("contact: the colliding geoms must share one contact parameter set")
    fr, sr, si, margin, condim = next(iter(params))
    if condim != 3:
        raise mjcf.MjcfError("contact: only condim 3 is handled")
    t.con_friction[:] = list(fr)
    t.con_solref[:] = list(sr)
    t.con_solimp[:] = list(si)
    t.con_margin = float(margin)
    t.con_condim = int(condim)
    t.con_box[:] = [box.pos[0] - box.size[0], box.pos[0] + box.size[0], box.pos[1] - box.size[1], box.pos[1] + box.size[1]]
    t.con_enabled = 1


def build_hulls(cm: mjcf.CompiledModel, res: int = CUBE_RES) -> Dict[str, np.ndarray]:
    """Hull vertices (body frame), their edge graph (CSR) and a direction -> start-vertex cube map per colliding geom,
    in tripwire-box order: what `so101_model_set_hulls` takes.  Convex hulls by qhull (scipy), as MuJoCo builds them
    for mesh collision.  The edge graph contains every polytope edge (plus diagonals of triangulated flat faces), so a
    steepest-ascent walk on it ends on the support vertex of any direction."""
    from scipy.spatial import ConvexHull
    t = cm.tables
    verts, vstart, adj_rows = [], [0], []
    cube = []
    for b, name, pts in _contact_geoms(cm):
        hull = ConvexHull(pts)
        keep = np.sort(hull.vertices)
        remap = -np.ones(len(pts), dtype=np.int64)
        remap[keep] = np.arange(len(keep))
        v = pts[keep]
        nbr = [set() for _ in range(len(v))]
        for tri in remap[hull.simplices]:
            for a_, b_ in ((0, 1), (1, 2), (0, 2)):
                nbr[tri[a_]].add(int(tri[b_])); nbr[tri[b_]].add(int(tri[a_]))
        base = vstart[-1]
        for n in nbr:
            adj_rows.append(sorted(base + x for x in n))
        # cube map: support vertex of the direction through the centre of every cell
        c = (np.arange(res) + 0.5) / res * 2 - 1
        uu, vv = np.meshgrid(c, c, indexing="ij")
        faces = []
        for face in range(6):
            ax, sgn = face // 2, (1.0 if face % 2 == 0 else -1.0)
            d = np.zeros((res, res, 3))
            d[..., ax] = sgn
            d[..., (ax + 1) % 3] = uu * sgn if False else uu
            d[..., (ax + 2) % 3] = vv
            faces.append(base + np.argmax(d.reshape(-1, 3) @ v.T, axis=1).reshape(res, res))
        cube.append(np.stack(faces))
        verts.append(v)
        vstart.append(base + len(v))
    adj_start = np.zeros(vstart[-1] + 1, dtype=np.int32)
    adj_start[1:] = np.cumsum([len(r) for r in adj_rows])
    return {
        "vert_start": np.asarray(vstart, dtype=np.int32), "vert": np.ascontiguousarray(np.concatenate(verts), dtype=np.float64),
        "adj_start": adj_start, "adj": np.asarray([x for r in adj_rows for x in r], dtype=np.int32),
        "cube": np.ascontiguousarray(np.stack(cube), dtype=np.int32), "cube_res": np.int32(res),
        "trip_center": np.ctypeslib.as_array(t.trip_center)[:t.ntrip].copy(),
        "trip_body": np.ctypeslib.as_array(t.trip_body)[:t.ntrip].copy(),
    }


def support_numpy(h: Dict[str, np.ndarray], g: int, d: np.ndarray) -> Tuple[int, float]:
    """Brute-force support vertex (global id, value) of geom g for direction d (body frame): the check for the walk."""
    lo, hi = h["vert_start"][g], h["vert_start"][g + 1]
    vals = h["vert"][lo:hi] @ d
    k = int(np.argmax(vals))
    return lo + k, float(vals[k])


def table_clearance_numpy(t: T.So101Tables, q: np.ndarray) -> float:
    """min over boxes of (lowest box point z - plane z): negative => TRIP_TABLE.  (numpy check)"""
    xpos, xmat, _, _ = mjcf.fk_numpy(t, q)
    best = np.inf
    for i in range(t.ntrip):
        b = t.trip_body[i]
        c = xpos[b] + xmat[b] @ np.array(t.trip_center[i][:])
        ax = np.array(t.trip_axes[i][:]).reshape(3, 3) @ xmat[b].T      # world axes (rows)
        ext = np.sum(np.abs(ax[:, 2]) * np.array(t.trip_half[i][:]))
        best = min(best, c[2] - ext - t.trip_plane_z)
    return float(best)


def obb_overlap(c1: np.ndarray, A1: np.ndarray, h1: np.ndarray, c2: np.ndarray, A2: np.ndarray, h2: np.ndarray) -> bool:
    """Separating-axis test of two oriented boxes (centre, rows = axes, half sizes; all in one frame): the 15 candidate
    axes (3 + 3 face normals, 9 edge cross products), with the customary 1e-12 guard on the absolute rotation for
    near-parallel edges (it can only make the answer 'overlap', i.e. keeps the test conservative)."""
    R = A1 @ A2.T
    tv = A1 @ (c2 - c1)
    AR = np.abs(R) + 1e-12
    for i in range(3):
        if abs(tv[i]) > h1[i] + AR[i] @ h2:
            return False
    for j in range(3):
        if abs(tv @ R[:, j]) > h1 @ AR[:, j] + h2[j]:
            return False
    for i in range(3):
        i1, i2 = (i + 1) % 3, (i + 2) % 3
        for j in range(3):
            j1, j2 = (j + 1) % 3, (j + 2) % 3
            ra = h1[i1] * AR[i2, j] + h1[i2] * AR[i1, j]
            rb = h2[j1] * AR[i, j2] + h2[j2] * AR[i, j1]
            if abs(tv[i2] * R[i1, j] - tv[i1] * R[i2, j]) > ra + rb:
                return False
    return True


def self_overlap_numpy(t: T.So101Tables, q: np.ndarray, margin: bool = False):
    """Do the oriented boxes of two colliding geoms on non-adjacent links overlap at pose q?  All `ntrip + nself` boxes of
    the tables; links a, b are non-adjacent when neither is the other's parent.  With `margin` returns the smallest
    slack over all axes of all pairs as well (how close the decision is), for tests."""
    xpos, xmat, _, _ = mjcf.fk_numpy(t, q)
    nb = t.ntrip + t.nself
    W = []
    for i in range(nb):
        b = t.trip_body[i]
        ax = np.array(t.trip_axes[i][:]).reshape(3, 3)
        W.append((b, xpos[b] + xmat[b] @ np.array(t.trip_center[i][:]), ax @ xmat[b].T, np.array(t.trip_half[i][:])))
    hit = False
    for i in range(nb):
        for j in range(i + 1, nb):
            a, b = W[i][0], W[j][0]
            if a == b or t.body_parent[a] == b or t.body_parent[b] == a:
                continue
            hit |= obb_overlap(*W[i][1:], *W[j][1:])
    return hit
