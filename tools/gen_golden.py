#!/usr/bin/env python
"""Generate the small fixtures under tests/golden/ (run in the build container).

  koopman_dkuc.npz   weights of the reference's trained DKUC Koopman model
                     [REF results/SOARM101/11_27/DKUC/best_model.pt], the only in-container
                     fingerprint of the true MuJoCo dynamics.  Extracted
                     with torch.load; architecture [REF models/KoopmanBase.py:12-60, args.py:103].
  oracle_v.npz /     regression pins of the CPU oracle on scene A / scene B: stage outputs of one
  oracle_p.npz       mj_forward at fixed states (M, qfrc_bias, site, qacc) and a short rollout.
                     These pin the ORACLE against accidental edits; they are NOT MuJoCo outputs
                     (MuJoCo is not installable here: "parity unpinned").
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def koopman():
    import torch
    sd = torch.load(os.path.join(REF, "results/SOARM101/11_27/DKUC/best_model.pt"), map_location="cpu",
                    weights_only=False)
    if not isinstance(sd, dict):
        sd = sd.state_dict()
    np.savez_compressed(os.path.join(GOLD, "koopman_dkuc.npz"), **{k: v.numpy() for k, v in sd.items()})


def oracle_pins():
    from lerobot_mujoco_sim2real_b200 import builtin_tables
    from oracle import oracle as O
    for tag, scene in (("v", "scene_with_table_v.xml"), ("p", "scene_with_table.xml")):
        t = builtin_tables(scene)
        rng = np.random.default_rng(2024)
        q = rng.uniform(-1.0, 1.0, (4, 6))
        q[3] = [1.93, -1.75, 1.70, 0.0, 2.85, -0.18]     # beyond several joint limits
        v = rng.uniform(-2.0, 2.0, (4, 6))
        u = rng.uniform(-2.5, 2.5, (4, 6))
        w = rng.uniform(-50, 50, (4, 6))
        out = {k: [] for k in ("M", "qfrc_bias", "site_xpos", "qacc_smooth", "qacc", "nefc", "efc_aref", "efc_R")}
        for i in range(4):
            o = O.Oracle(t)
            o.reset(); o.set("qpos", q[i]); o.set("qvel", v[i]); o.set("ctrl", u[i]); o.set("qacc_warmstart", w[i])
            o.forward()
            out["M"].append(o.full_M()); out["nefc"].append(o.d.nefc)
            for k in ("qfrc_bias", "site_xpos", "qacc_smooth", "qacc", "efc_aref", "efc_R"):
                out[k].append(o.arr(k).copy())
        rows, final, iters = O.rollout(t, O.make_spec(kind=0, seed=7), 8, 20, 10)
        rows_sin, _, _ = O.rollout(t, O.make_spec(kind=1, seed=7), 4, 20, 10)
        rows_chirp, _, _ = O.rollout(t, O.make_spec(kind=2, seed=7), 4, 20, 10)
        np.savez_compressed(os.path.join(GOLD, f"oracle_{tag}.npz"), q=q, v=v, u=u, w=w, rows=rows, final=final,
                            rows_sin=rows_sin, rows_chirp=rows_chirp, newton_iters=iters,
                            **{k: np.array(x) for k, x in out.items()})


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    koopman()
    oracle_pins()
    print(sorted(os.listdir(GOLD)))
