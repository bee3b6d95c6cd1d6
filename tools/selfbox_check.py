#!/usr/bin/env python
"""Evidence for the fast-accept joint box of the self-collision flag: N uniformly sampled poses inside it (free joints over their
whole range), the numpy box-box test of the non-adjacent links at each.  usage: tools/selfbox_check.py [N] [procs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from multiprocessing import Pool
from lerobot_mujoco_sim2real_b200 import builtin_tables, tripwire


def work(args):
    seed, n = args
    t = builtin_tables()
    lo_r = np.array([t.jnt_range[k][0] for k in range(6)]); hi_r = np.array([t.jnt_range[k][1] for k in range(6)])
    lo = np.maximum(lo_r, [t.trip_qbox[k][0] for k in range(6)]); hi = np.minimum(hi_r, [t.trip_qbox[k][1] for k in range(6)])
    rng = np.random.default_rng(seed)
    hits = 0
    for _ in range(n):
        # half of the samples on the faces of the box, where a collision would appear first
        q = rng.uniform(lo, hi)
        if rng.random() < 0.5:
            k = rng.integers(1, 6)
            q[k] = hi[k] if rng.random() < 0.5 else lo[k]
        hits += bool(tripwire.self_overlap_numpy(t, q))
    return hits


if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    P = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
    with Pool(P) as pool:
        res = pool.map(work, [(1000 + i, N // P) for i in range(P)])
    t = builtin_tables()
    print(f"box {[[round(float(x), 3) for x in t.trip_qbox[k][:]] for k in range(1, 6)]} (joint 0 free): {sum(res)} of {N // P * P} sampled poses "
          f"(half of them on a face of the box) have overlapping link boxes")
