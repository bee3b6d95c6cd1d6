"""Contact tripwire tables (host, offline): conservative detection of "this env may be in contact".

The hot-path scene has contacts enabled (13 collision meshes vs a table box and a floor plane,
[REF SOARM101/SO101/so101_new_calib_v.xml:53-117, scene_with_table_v.xml:28,31])
but the CUDA kernels do not simulate contact.  Instead every env carries two flag bits:

  TRIP_TABLE  an oriented bounding box of a link's collision geometry dipped below the highest
              static support plane (the table top).  Box ⊇ convex hull ⊇ mesh, so a table contact
              can never be missed (no false negatives); it may fire slightly early.
  TRIP_SELF   the joint vector left a box |q_i| <= qbox_i inside which no pair of non-adjacent
              link hulls intersects.  The box is found by sampling (hull-hull LP feasibility) and
              shrunk by a safety factor: a heuristic, documented as such.

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


def self_collision_box(cm: mjcf.CompiledModel, hulls: Dict[int, np.ndarray], samples: int = 120,
                       seed: int = 0, safety: float = 0.8) -> np.ndarray:
    """Largest (of a fixed ladder) symmetric joint box without sampled self-intersection, x safety."""
    t = cm.tables
    rng = np.random.default_rng(seed)
    bodies = sorted(hulls)
    pairs = [(a, b) for i, a in enumerate(bodies) for b in bodies[i + 1:]
             if t.body_parent[b] != a and t.body_parent[a] != b]
    small = {b: hulls[b][np.linspace(0, len(hulls[b]) - 1, min(len(hulls[b]), 60)).astype(int)] for b in bodies}
    lo = np.array([t.jnt_range[k][0] for k in range(t.nv)])
    hi = np.array([t.jnt_range[k][1] for k in range(t.nv)])
    best = 0.0
    for scale in (0.3, 0.45, 0.6, 0.8, 1.0, 1.2, 1.5):
        ok = True
        for _ in range(samples):
            q = rng.uniform(np.maximum(lo, -scale), np.minimum(hi, scale))
            xpos, xmat, _, _ = mjcf.fk_numpy(t, q)
            world = {b: xpos[b] + small[b] @ xmat[b].T for b in bodies}
            if any(_hulls_intersect(world[a], world[b]) for a, b in pairs):
                ok = False
                break
        if not ok:
            break
        best = scale
    box = np.zeros((t.nv, 2))
    box[:, 0], box[:, 1] = np.maximum(lo, -best * safety), np.minimum(hi, best * safety)
    return box


def fill_tripwire(cm: mjcf.CompiledModel, with_self_box: bool = True) -> None:
    """Fill `cm.tables.trip_*` in place from the collision meshes (needs the reference assets).

    One box per colliding geom.  Geoms of a link whose height above the plane cannot change (first
    link of the chain turning about a vertical axis) are checked once here and get no box."""
    t = cm.tables
    plane = support_plane_z(cm)
    t.ntrip = 0
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
    if len(boxes) > T.MAXTRIP:
        raise mjcf.MjcfError("too many tripwire boxes")
    t.ntrip = len(boxes)
    for i, (b, c, ax, half) in enumerate(boxes):
        t.trip_body[i] = b
        t.trip_center[i][:] = list(c)
        t.trip_axes[i][:] = list(ax.reshape(-1))
        t.trip_half[i][:] = list(half)
    t.trip_plane_z = plane
    fill_contact_params(cm)
    if with_self_box:
        box = self_collision_box(cm, body_hulls(cm))
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
