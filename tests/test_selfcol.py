"""Self-collision flag (TRIP_SELF): the host side - the boxes in the tables, the numpy restatement of the box-box test the
kernels are checked against (tests/test_gpu_api.py), and, where the reference meshes are present, that the test never
misses an intersection of the real hulls."""
import numpy as np
import pytest

from lerobot_mujoco_sim2real_b200 import builtin_tables, mjcf, tripwire

REF_SO101 = "/root/reference/SOARM101/SO101"


def _ranges(t):
    return (np.array([t.jnt_range[k][0] for k in range(6)]), np.array([t.jnt_range[k][1] for k in range(6)]))


def test_tables_carry_every_colliding_geom():
    t = builtin_tables()
    assert t.ntrip == 10 and t.nself == 3                      # 13 collision meshes [REF so101_new_calib_v.xml:53-117]
    bodies = [t.trip_body[i] for i in range(t.ntrip + t.nself)]
    assert bodies[t.ntrip:] == [t.jnt_body[0]] * 3             # the first link's geoms: no table box, self test only
    for i in range(t.ntrip + t.nself):
        ax = np.array(t.trip_axes[i][:]).reshape(3, 3)
        assert np.allclose(ax @ ax.T, np.eye(3), atol=1e-12) and min(t.trip_half[i][:]) > 0


def test_obb_overlap_known_cases():
    I, h = np.eye(3), np.array([1.0, 0.5, 0.25])
    z = np.zeros(3)
    assert tripwire.obb_overlap(z, I, h, np.array([1.9, 0, 0]), I, h)
    assert not tripwire.obb_overlap(z, I, h, np.array([2.1, 0, 0]), I, h)
    c, s = np.cos(np.pi / 4), np.sin(np.pi / 4)
    Rz = np.array([[c, s, 0], [-s, c, 0], [0, 0, 1.0]])
    cube = np.ones(3)
    # two unit cubes, one turned by 45 degrees about z: face axes alone see them apart up to a centre distance of 1 + sqrt 2
    assert tripwire.obb_overlap(z, I, cube, np.array([2.3, 0, 0]), Rz, cube)
    assert not tripwire.obb_overlap(z, I, cube, np.array([2.5, 0, 0]), Rz, cube)
    # an edge-edge case: only a cross-product axis separates them
    Rx = np.array([[1, 0, 0], [0, c, s], [0, -s, c]])
    A, B = Rz, Rx @ Rz.T
    d = np.array([1.0, 1.0, 1.0]) / np.sqrt(3)
    far = next(r for r in np.arange(1.0, 4.0, 0.01) if not tripwire.obb_overlap(z, A, cube, r * d, B, cube))
    pts = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=float)
    from scipy.optimize import linprog

    def hulls_meet(P, Q):
        na, nb = len(P), len(Q)
        Aeq = np.zeros((5, na + nb)); Aeq[:3, :na], Aeq[:3, na:] = P.T, -Q.T; Aeq[3, :na], Aeq[4, na:] = 1, 1
        return linprog(np.zeros(na + nb), A_eq=Aeq, b_eq=[0, 0, 0, 1, 1.0], bounds=(0, None), method="highs").status == 0
    assert hulls_meet(pts @ A, (far - 0.03) * d + pts @ B) and not hulls_meet(pts @ A, (far + 0.01) * d + pts @ B)


def test_no_overlap_inside_the_fast_accept_box():
    t = builtin_tables()
    lo_r, hi_r = _ranges(t)
    lo = np.maximum(lo_r, [t.trip_qbox[k][0] for k in range(6)]); hi = np.minimum(hi_r, [t.trip_qbox[k][1] for k in range(6)])
    assert lo[0] == lo_r[0] and hi[0] == hi_r[0] and hi[2] > 0.6       # the first joint moves no colliding geom against another
    rng = np.random.default_rng(0)
    q = rng.uniform(lo, hi, (600, 6))
    corners = np.array([[lo[k] if (c >> k) & 1 else hi[k] for k in range(6)] for c in range(64)])
    assert not any(tripwire.self_overlap_numpy(t, x) for x in np.concatenate([q, corners]))
    far = rng.uniform(np.maximum(lo_r, -1.6), np.minimum(hi_r, 1.6), (300, 6))
    assert 0.02 < np.mean([tripwire.self_overlap_numpy(t, x) for x in far]) < 0.4


@pytest.mark.reference
def test_box_test_never_misses_a_hull_intersection():
    """Box contains hull: whenever two non-adjacent link hulls intersect (LP feasibility on the real meshes) the box test
    fires too, and it fires rarely without (tight boxes): 400 poses over most of the joint ranges."""
    cm = mjcf.compile_mjcf(f"{REF_SO101}/scene_with_table_v.xml")
    tripwire.fill_tripwire(cm, with_self_box=False)
    t = cm.tables
    hulls = tripwire.body_hulls(cm)
    bodies = sorted(hulls)
    small = {b: hulls[b][np.linspace(0, len(hulls[b]) - 1, min(len(hulls[b]), 100)).astype(int)] for b in bodies}
    pairs = [(a, b) for i, a in enumerate(bodies) for b in bodies[i + 1:] if t.body_parent[b] != a and t.body_parent[a] != b]
    lo_r, hi_r = _ranges(t)
    rng = np.random.default_rng(5)
    n_box = n_hull = 0
    for _ in range(400):
        q = rng.uniform(np.maximum(lo_r, -1.5), np.minimum(hi_r, 1.5))
        box = tripwire.self_overlap_numpy(t, q)
        n_box += box
        xpos, xmat, _, _ = mjcf.fk_numpy(t, q)
        world = {b: xpos[b] + small[b] @ xmat[b].T for b in bodies}
        hull = any(tripwire._hulls_intersect(world[a], world[b]) for a, b in pairs)
        n_hull += hull
        assert box or not hull
    assert n_hull > 10 and n_box < 2 * n_hull + 10
