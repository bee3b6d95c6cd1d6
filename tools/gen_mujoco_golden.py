#!/usr/bin/env python
"""Golden vectors of the REAL `mujoco` engine for the SO-ARM101 stepping path -> tests/golden/mujoco_{v,p}.npz.

This is the recipe SURVEY.md 7 step 0 / 8c item 4 ask for: run it once wherever the `mujoco` wheel and the reference's
MJCF files are available (neither exists in the build container nor on the GPU box: "parity unpinned"),

    pip install mujoco
    python tools/gen_mujoco_golden.py --ref /path/to/Lerobot-mujoco-sim2real        # writes both scenes

commit the two .npz files, and `tests/test_mujoco_pin.py` pins the C oracle (CPU suite) and the CUDA kernels
(`-m gpu`) against MuJoCo itself: compiled constants, every stage of mj_forward, teacher-forced mj_step (P1),
the free-running curve against MuJoCo's own 1-ulp self-divergence (P2), the contractive scene free-running (P3)
and the contact rows of table-plane contacts.

What is recorded (all float64; `S` = physics steps, `E` = envs, `K` = forward-stage states):
  meta_*      mujoco version, scene file, numpy seed, source ("mujoco")
  m_*         the compiled mjModel constants of SURVEY Appendix B that mjcf.py reproduces
  fw_*        K states (qpos, qvel, ctrl, qacc_warmstart in / stage outputs of mj_forward out), contacts disabled
  tf_*        E x S teacher-forcing trajectory (state BEFORE every step incl. qacc_warmstart, ctrl, qacc), contacts
              disabled (the contact-free pipeline of SURVEY Appendix A); reset U(-0.3,0.3), ctrl U(-0.5,0.5) redrawn every
              `frame_skip` steps (BASELINE config 2's distribution)
  tfc_*       the same initial states and controls with contacts ENABLED (the reference's actual scene) + ncon per step
  sd_*        self-divergence: the tf run repeated from qpos + 1 ulp, max |dqpos|, |dqvel| per step (F3)
  ct_*        K2 poses in |q| <= 1.0 with contacts enabled: every contact's geoms, dist, pos, frame, and the
              constraint rows / qacc of that forward (pins the table-plane contact path)

`--backend oracle` writes the SAME schema from oracle/so101_oracle.c (meta_source = "oracle-selftest"): it only lets
the harness of tests/test_mujoco_pin.py be exercised without MuJoCo and is never a pin (the tests refuse to treat such
a file as one).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
SCENES = {"v": "scene_with_table_v.xml", "p": "scene_with_table.xml"}
NV, MAXEFC, MAXCON = 6, 40, 8
FRAME_SKIP = 10


# --------------------------------------------------------------------------------------------------------------------
# inputs shared by both backends (functions of the seed only)
# --------------------------------------------------------------------------------------------------------------------
def make_inputs(tag: str, seed: int, E: int, S: int, K: int, K2: int, jnt_range: np.ndarray):
    rng = np.random.default_rng(seed)
    fw = dict(qpos=rng.uniform(-1.0, 1.0, (K, NV)), qvel=rng.uniform(-2.0, 2.0, (K, NV)),
              ctrl=rng.uniform(-2.5, 2.5, (K, NV)), warm=rng.uniform(-50.0, 50.0, (K, NV)))
    # a quarter of the forward states sit on / beyond joint limits (limit rows active)
    for i in range(0, K, 4):
        j = rng.integers(0, NV, 2)
        side = rng.integers(0, 2, 2)
        fw["qpos"][i, j] = jnt_range[j, side] + rng.uniform(-0.02, 0.05, 2) * np.where(side == 1, 1.0, -1.0)
    fw["qvel"][:4] = 0.0                                   # static cases (gravity bias only)
    uscale = 0.5 if tag == "v" else 1.5                    # scene B: position targets, wider range
    q0 = np.zeros((E, NV))
    q0[:, :5] = rng.uniform(-0.3, 0.3, (E, 5))
    ctrl = np.zeros((E, S, NV))
    for t in range(0, S, FRAME_SKIP):
        ctrl[:, t:t + FRAME_SKIP, :5] = rng.uniform(-uscale, uscale, (E, 1, 5))
    ct_q = rng.uniform(-1.0, 1.0, (K2, NV))
    ct_q[:, 5] = rng.uniform(-0.17, 1.0, K2)
    ct_v = rng.uniform(-1.0, 1.0, (K2, NV))
    return fw, q0, ctrl, ct_q, ct_v


# --------------------------------------------------------------------------------------------------------------------
# MuJoCo backend
# --------------------------------------------------------------------------------------------------------------------
def run_mujoco(xml: str, tag: str, seed: int, E: int, S: int, K: int, K2: int) -> dict:
    import mujoco
    m = mujoco.MjModel.from_xml_path(xml)
    assert m.nq == NV and m.nv == NV and m.nu == NV, "not the SO-ARM101 hinge chain"
    out = {"meta_version": np.array(mujoco.__version__), "meta_scene": np.array(os.path.basename(xml)),
           "meta_seed": np.array(seed), "meta_source": np.array("mujoco"), "meta_frame_skip": np.array(FRAME_SKIP)}
    site = mujoco.mj_name2id(m, mujoco.mjtObj.mjOBJ_SITE, "gripperframe")
    out.update({
        "m_timestep": m.opt.timestep, "m_gravity": np.array(m.opt.gravity), "m_tolerance": m.opt.tolerance,
        "m_ls_tolerance": m.opt.ls_tolerance, "m_iterations": m.opt.iterations, "m_ls_iterations": m.opt.ls_iterations,
        "m_integrator": m.opt.integrator, "m_solver": m.opt.solver, "m_cone": m.opt.cone, "m_impratio": m.opt.impratio,
        "m_disableflags": m.opt.disableflags, "m_meaninertia": m.stat.meaninertia, "m_nbody": m.nbody,
        "m_body_parentid": np.array(m.body_parentid), "m_body_pos": np.array(m.body_pos),
        "m_body_quat": np.array(m.body_quat), "m_body_ipos": np.array(m.body_ipos),
        "m_body_iquat": np.array(m.body_iquat), "m_body_inertia": np.array(m.body_inertia),
        "m_body_mass": np.array(m.body_mass), "m_body_invweight0": np.array(m.body_invweight0),
        "m_jnt_bodyid": np.array(m.jnt_bodyid), "m_jnt_pos": np.array(m.jnt_pos), "m_jnt_axis": np.array(m.jnt_axis),
        "m_jnt_range": np.array(m.jnt_range), "m_jnt_limited": np.array(m.jnt_limited),
        "m_jnt_margin": np.array(m.jnt_margin), "m_jnt_solref": np.array(m.jnt_solref),
        "m_jnt_solimp": np.array(m.jnt_solimp), "m_jnt_stiffness": np.array(m.jnt_stiffness),
        "m_qpos0": np.array(m.qpos0), "m_qpos_spring": np.array(m.qpos_spring),
        "m_dof_armature": np.array(m.dof_armature), "m_dof_damping": np.array(m.dof_damping),
        "m_dof_frictionloss": np.array(m.dof_frictionloss), "m_dof_solref": np.array(m.dof_solref),
        "m_dof_solimp": np.array(m.dof_solimp), "m_dof_invweight0": np.array(m.dof_invweight0),
        "m_dof_M0": np.array(m.dof_M0),
        "m_actuator_gainprm": np.array(m.actuator_gainprm[:, :3]), "m_actuator_biasprm": np.array(m.actuator_biasprm[:, :3]),
        "m_actuator_ctrlrange": np.array(m.actuator_ctrlrange), "m_actuator_forcerange": np.array(m.actuator_forcerange),
        "m_actuator_ctrllimited": np.array(m.actuator_ctrllimited),
        "m_actuator_forcelimited": np.array(m.actuator_forcelimited), "m_actuator_gear": np.array(m.actuator_gear[:, 0]),
        "m_site_pos": np.array(m.site_pos[site]), "m_site_quat": np.array(m.site_quat[site]),
        "m_site_bodyid": m.site_bodyid[site],
        "m_key_qpos": np.array(m.key_qpos[0]) if m.nkey else np.zeros(NV),
        "m_key_ctrl": np.array(m.key_ctrl[0]) if m.nkey else np.zeros(NV),
        # collision side (N1): geoms that can touch the table / floor
        "m_geom_bodyid": np.array(m.geom_bodyid), "m_geom_type": np.array(m.geom_type),
        "m_geom_contype": np.array(m.geom_contype), "m_geom_conaffinity": np.array(m.geom_conaffinity),
        "m_geom_pos": np.array(m.geom_pos), "m_geom_quat": np.array(m.geom_quat), "m_geom_size": np.array(m.geom_size),
        "m_geom_friction": np.array(m.geom_friction), "m_geom_solref": np.array(m.geom_solref),
        "m_geom_solimp": np.array(m.geom_solimp), "m_geom_margin": np.array(m.geom_margin),
        "m_geom_condim": np.array(m.geom_condim), "m_geom_dataid": np.array(m.geom_dataid),
    })
    fw, q0, ctrl, ct_q, ct_v = make_inputs(tag, seed, E, S, K, K2, np.array(m.jnt_range))
    CONTACT = int(mujoco.mjtDisableBit.mjDSBL_CONTACT)

    def niter(d):
        v = np.atleast_1d(np.asarray(d.solver_niter))
        return int(v[0])

    def efc(d, key):
        a = np.zeros((MAXEFC,) + ((NV,) if key == "efc_J" else ()))
        n = min(int(d.nefc), MAXEFC)
        src = np.asarray(getattr(d, key)).reshape(-1)
        if key == "efc_J":
            a[:n] = src[:int(d.nefc) * NV].reshape(int(d.nefc), NV)[:n]
        else:
            a[:n] = src[:n]
        return a

    # ---- forward stages, contacts disabled --------------------------------------------------------------------------
    m.opt.disableflags |= CONTACT
    d = mujoco.MjData(m)
    keys = ["xpos", "xquat", "xipos", "site_xpos", "subtree_com", "cinert", "cdof", "qfrc_bias", "qfrc_passive",
            "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "qacc", "qfrc_constraint"]
    rec = {k: [] for k in keys + ["qM", "nefc", "niter", "efc_type", "efc_pos", "efc_D", "efc_R", "efc_aref", "efc_J",
                                  "efc_force"]}
    for i in range(K):
        mujoco.mj_resetData(m, d)
        d.qpos[:] = fw["qpos"][i]; d.qvel[:] = fw["qvel"][i]; d.ctrl[:] = fw["ctrl"][i]
        d.qacc_warmstart[:] = fw["warm"][i]
        mujoco.mj_forward(m, d)
        for k in keys:
            a = np.array(getattr(d, k))
            rec[k].append(a[site] if k == "site_xpos" else a)
        M = np.zeros((NV, NV)); mujoco.mj_fullM(m, M, d.qM)
        rec["qM"].append(M); rec["nefc"].append(int(d.nefc)); rec["niter"].append(niter(d))
        for k in ("efc_type", "efc_pos", "efc_D", "efc_R", "efc_aref", "efc_J", "efc_force"):
            rec[k].append(efc(d, k))
    out.update({f"fw_in_{k}": v for k, v in fw.items()})
    out.update({f"fw_{k}": np.array(v) for k, v in rec.items()})

    # ---- teacher-forcing trajectories -------------------------------------------------------------------------------
    def rollout(q_init, contacts: bool):
        if contacts:
            m.opt.disableflags &= ~CONTACT
        else:
            m.opt.disableflags |= CONTACT
        Q = np.zeros((E, S + 1, NV)); V = np.zeros((E, S + 1, NV)); W = np.zeros((E, S + 1, NV))
        A = np.zeros((E, S, NV)); N = np.zeros((E, S), dtype=np.int32); SITE = np.zeros((E, S + 1, 3))
        NI = np.zeros((E, S), dtype=np.int32)
        for e in range(E):
            mujoco.mj_resetData(m, d)
            d.qpos[:] = q_init[e]
            mujoco.mj_forward(m, d)                      # as SOARM101Env.reset does [REF SOARM101_Env.py:102]
            d.qacc_warmstart[:] = 0.0                    # mj_forward leaves it untouched; state it
            for t in range(S):
                Q[e, t] = d.qpos; V[e, t] = d.qvel; W[e, t] = d.qacc_warmstart; SITE[e, t] = d.site_xpos[site]
                d.ctrl[:] = ctrl[e, t]
                mujoco.mj_step(m, d)
                A[e, t] = d.qacc; N[e, t] = d.ncon; NI[e, t] = niter(d)
            Q[e, S] = d.qpos; V[e, S] = d.qvel; W[e, S] = d.qacc_warmstart; SITE[e, S] = d.site_xpos[site]
        return Q, V, W, A, N, SITE, NI

    Q, V, W, A, N, SITE, NI = rollout(q0, False)
    out.update(tf_qpos=Q, tf_qvel=V, tf_warm=W, tf_ctrl=ctrl, tf_qacc=A, tf_site=SITE, tf_niter=NI)
    q1 = q0.copy()
    q1[:, :5] = np.nextafter(q1[:, :5], np.inf)
    Q1, V1, *_ = rollout(q1, False)
    out.update(sd_qpos=np.abs(Q1 - Q).max(axis=(0, 2)), sd_qvel=np.abs(V1 - V).max(axis=(0, 2)))
    Qc, Vc, Wc, Ac, Nc, SITEc, NIc = rollout(q0, True)
    out.update(tfc_qpos=Qc, tfc_qvel=Vc, tfc_warm=Wc, tfc_qacc=Ac, tfc_ncon=Nc, tfc_site=SITEc, tfc_niter=NIc)

    # ---- contact poses ----------------------------------------------------------------------------------------------
    m.opt.disableflags &= ~CONTACT
    ct = {k: [] for k in ("ncon", "geom1", "geom2", "dist", "pos", "frame", "friction", "dim", "includemargin",
                          "nefc", "efc_type", "efc_J", "efc_D", "efc_R", "efc_aref", "efc_pos", "efc_force", "qacc",
                          "qfrc_constraint")}
    for i in range(K2):
        mujoco.mj_resetData(m, d)
        d.qpos[:] = ct_q[i]; d.qvel[:] = ct_v[i]
        mujoco.mj_forward(m, d)
        nc = min(int(d.ncon), MAXCON)
        g1 = np.full(MAXCON, -1); g2 = np.full(MAXCON, -1); dist = np.zeros(MAXCON); pos = np.zeros((MAXCON, 3))
        frame = np.zeros((MAXCON, 9)); fr = np.zeros((MAXCON, 5)); dim = np.zeros(MAXCON, dtype=np.int32)
        inc = np.zeros(MAXCON)
        for c in range(nc):
            con = d.contact[c]
            g1[c], g2[c], dist[c], dim[c], inc[c] = con.geom1, con.geom2, con.dist, con.dim, con.includemargin
            pos[c] = con.pos; frame[c] = con.frame; fr[c] = con.friction
        ct["ncon"].append(int(d.ncon)); ct["geom1"].append(g1); ct["geom2"].append(g2); ct["dist"].append(dist)
        ct["pos"].append(pos); ct["frame"].append(frame); ct["friction"].append(fr); ct["dim"].append(dim)
        ct["includemargin"].append(inc); ct["nefc"].append(int(d.nefc))
        for k in ("efc_type", "efc_J", "efc_D", "efc_R", "efc_aref", "efc_pos", "efc_force"):
            ct[k].append(efc(d, k))
        ct["qacc"].append(np.array(d.qacc)); ct["qfrc_constraint"].append(np.array(d.qfrc_constraint))
    out.update(ct_in_qpos=ct_q, ct_in_qvel=ct_v)
    out.update({f"ct_{k}": np.array(v) for k, v in ct.items()})
    return out


# --------------------------------------------------------------------------------------------------------------------
# oracle backend: same schema, for exercising the test harness only
# --------------------------------------------------------------------------------------------------------------------
def run_oracle(tag: str, seed: int, E: int, S: int, K: int, K2: int) -> dict:
    from lerobot_mujoco_sim2real_b200 import builtin_tables, tables as T_
    from oracle import oracle as O
    t = builtin_tables(SCENES[tag])
    O.set_hulls(None)            # fw_* / tf_* / sd_* are the contact-free pipeline (MuJoCo backend: mjDSBL_CONTACT)
    try:
        return _run_oracle(O, t, tag, seed, E, S, K, K2)
    finally:
        O.set_hulls(T_.builtin_hulls())


def _run_oracle(O, t, tag: str, seed: int, E: int, S: int, K: int, K2: int) -> dict:
    out = {"meta_version": np.array("oracle"), "meta_scene": np.array(SCENES[tag]), "meta_seed": np.array(seed),
           "meta_source": np.array("oracle-selftest"), "meta_frame_skip": np.array(FRAME_SKIP)}
    arr = lambda x: np.ctypeslib.as_array(x).copy()
    nb = t.nbody
    out.update({
        "m_timestep": t.timestep, "m_gravity": arr(t.gravity), "m_tolerance": t.tolerance,
        "m_ls_tolerance": t.ls_tolerance, "m_iterations": t.iterations, "m_ls_iterations": t.ls_iterations,
        "m_meaninertia": t.meaninertia, "m_nbody": nb, "m_body_parentid": arr(t.body_parent)[:nb],
        "m_body_pos": arr(t.body_pos)[:nb], "m_body_quat": arr(t.body_quat)[:nb], "m_body_ipos": arr(t.body_ipos)[:nb],
        "m_body_iquat": arr(t.body_iquat)[:nb], "m_body_inertia": arr(t.body_inertia)[:nb],
        "m_body_mass": arr(t.body_mass)[:nb], "m_jnt_bodyid": arr(t.jnt_body), "m_jnt_pos": arr(t.jnt_pos),
        "m_jnt_axis": arr(t.jnt_axis), "m_jnt_range": arr(t.jnt_range), "m_jnt_limited": arr(t.jnt_limited),
        "m_jnt_margin": arr(t.jnt_margin), "m_jnt_solref": arr(t.jnt_solref), "m_jnt_solimp": arr(t.jnt_solimp),
        "m_jnt_stiffness": arr(t.jnt_stiffness), "m_qpos0": arr(t.qpos0), "m_qpos_spring": arr(t.qpos_spring),
        "m_dof_armature": arr(t.dof_armature), "m_dof_damping": arr(t.dof_damping),
        "m_dof_frictionloss": arr(t.dof_frictionloss), "m_dof_solref": arr(t.dof_solref),
        "m_dof_solimp": arr(t.dof_solimp), "m_dof_invweight0": arr(t.dof_invweight0), "m_dof_M0": arr(t.dof_M0),
        "m_actuator_gainprm": np.stack([arr(t.act_gain), np.zeros(NV), np.zeros(NV)], 1),
        "m_actuator_biasprm": arr(t.act_bias), "m_actuator_ctrlrange": arr(t.act_ctrlrange),
        "m_actuator_forcerange": arr(t.act_forcerange), "m_actuator_ctrllimited": arr(t.act_ctrllimited),
        "m_actuator_forcelimited": arr(t.act_forcelimited), "m_actuator_gear": arr(t.act_gear),
        "m_site_pos": arr(t.site_pos), "m_site_quat": arr(t.site_quat), "m_site_bodyid": t.site_body,
        "m_key_qpos": arr(t.key_qpos), "m_key_ctrl": arr(t.key_ctrl),
    })
    fw, q0, ctrl, ct_q, ct_v = make_inputs(tag, seed, E, S, K, K2, arr(t.jnt_range))
    o = O.Oracle(t)
    keys = ["xpos", "xquat", "xipos", "site_xpos", "subtree_com", "cinert", "cdof", "qfrc_bias", "qfrc_passive",
            "qfrc_actuator", "qfrc_smooth", "qacc_smooth", "qacc", "qfrc_constraint"]
    ekeys = ("efc_type", "efc_pos", "efc_D", "efc_R", "efc_aref", "efc_J", "efc_force")
    rec = {k: [] for k in keys + ["qM", "nefc", "niter"] + list(ekeys)}

    def efc(key):
        a = np.zeros((MAXEFC,) + ((NV,) if key == "efc_J" else ()))
        src = o.arr(key)
        n = min(o.d.nefc, MAXEFC, src.shape[0])
        a[:n] = src[:n]
        return a

    for i in range(K):
        o.reset(); o.set("qpos", fw["qpos"][i]); o.set("qvel", fw["qvel"][i]); o.set("ctrl", fw["ctrl"][i])
        o.set("qacc_warmstart", fw["warm"][i])
        o.forward()
        for k in keys:
            a = o.arr(k).copy()
            rec[k].append(a[:nb] if a.ndim == 2 and a.shape[0] == O.NB else a)
        rec["qM"].append(o.full_M()); rec["nefc"].append(o.d.nefc); rec["niter"].append(o.d.solver_niter)
        for k in ekeys:
            rec[k].append(efc(k))
    out.update({f"fw_in_{k}": v for k, v in fw.items()})
    out.update({f"fw_{k}": np.array(v) for k, v in rec.items()})

    def rollout(q_init):
        Q = np.zeros((E, S + 1, NV)); V = np.zeros((E, S + 1, NV)); W = np.zeros((E, S + 1, NV))
        A = np.zeros((E, S, NV)); SITE = np.zeros((E, S + 1, 3)); NI = np.zeros((E, S), dtype=np.int32)
        for e in range(E):
            o.reset(); o.set("qpos", q_init[e]); o.forward()
            for s in range(S):
                Q[e, s] = o.arr("qpos"); V[e, s] = o.arr("qvel"); W[e, s] = o.arr("qacc_warmstart")
                SITE[e, s] = o.arr("site_xpos")
                o.set("ctrl", ctrl[e, s]); o.step()
                A[e, s] = o.arr("qacc"); NI[e, s] = o.d.solver_niter
            Q[e, S] = o.arr("qpos"); V[e, S] = o.arr("qvel"); W[e, S] = o.arr("qacc_warmstart")
            SITE[e, S] = o.arr("site_xpos")
        return Q, V, W, A, SITE, NI

    Q, V, W, A, SITE, NI = rollout(q0)
    out.update(tf_qpos=Q, tf_qvel=V, tf_warm=W, tf_ctrl=ctrl, tf_qacc=A, tf_site=SITE, tf_niter=NI)
    q1 = q0.copy()
    q1[:, :5] = np.nextafter(q1[:, :5], np.inf)
    Q1, V1, *_ = rollout(q1)
    out.update(sd_qpos=np.abs(Q1 - Q).max(axis=(0, 2)), sd_qvel=np.abs(V1 - V).max(axis=(0, 2)))
    return out


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ref", default="/root/reference", help="checkout of the reference repository")
    ap.add_argument("--out", default=GOLD)
    ap.add_argument("--backend", choices=["mujoco", "oracle"], default="mujoco")
    ap.add_argument("--scenes", default="v,p")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--envs", type=int, default=4)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--states", type=int, default=64)
    ap.add_argument("--contact-states", type=int, default=256)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for tag in a.scenes.split(","):
        if a.backend == "mujoco":
            xml = os.path.join(a.ref, "SOARM101", "SO101", SCENES[tag])
            data = run_mujoco(xml, tag, a.seed, a.envs, a.steps, a.states, a.contact_states)
        else:
            data = run_oracle(tag, a.seed, a.envs, a.steps, a.states, a.contact_states)
        path = os.path.join(a.out, f"mujoco_{tag}.npz")
        np.savez_compressed(path, **data)
        print(f"{path}: {os.path.getsize(path) / 1e6:.2f} MB, source={data['meta_source']}, "
              f"version={data['meta_version']}")


if __name__ == "__main__":
    main()
