"""ctypes binding of oracle/so101_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  PARITY UNPINNED: see the header of so101_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from lerobot_mujoco_sim2real_b200.tables import So101CtrlSpec, So101Hulls, So101Tables  # noqa: E402

NV, NB, NM, MAXEFC, MAXCON = 6, 8, 21, 64, 10
_d, _i = C.c_double, C.c_int32


class OracleData(C.Structure):
    _fields_ = [
        ("qpos", _d * NV), ("qvel", _d * NV), ("qacc_warmstart", _d * NV), ("ctrl", _d * NV),
        ("qfrc_applied", _d * NV), ("time", _d),
        ("xpos", (_d * 3) * NB), ("xquat", (_d * 4) * NB), ("xmat", (_d * 9) * NB),
        ("xipos", (_d * 3) * NB), ("ximat", (_d * 9) * NB),
        ("xanchor", (_d * 3) * NV), ("xaxis", (_d * 3) * NV), ("site_xpos", _d * 3),
        ("subtree_com", (_d * 3) * NB), ("cinert", (_d * 10) * NB), ("crb", (_d * 10) * NB),
        ("cdof", (_d * 6) * NV),
        ("qM", _d * NM), ("qLD", _d * NM), ("qLDiagInv", _d * NV),
        ("cvel", (_d * 6) * NB), ("cdof_dot", (_d * 6) * NV),
        ("qfrc_bias", _d * NV), ("qfrc_passive", _d * NV), ("qfrc_actuator", _d * NV),
        ("actuator_force", _d * NV),
        ("qfrc_smooth", _d * NV), ("qacc_smooth", _d * NV), ("qfrc_constraint", _d * NV), ("qacc", _d * NV),
        ("nefc", _i), ("nf", _i),
        ("efc_type", _i * MAXEFC), ("efc_id", _i * MAXEFC), ("efc_state", _i * MAXEFC),
        ("efc_J", (_d * NV) * MAXEFC), ("efc_pos", _d * MAXEFC), ("efc_margin", _d * MAXEFC),
        ("efc_frictionloss", _d * MAXEFC),
        ("efc_diagApprox", _d * MAXEFC), ("efc_R", _d * MAXEFC), ("efc_D", _d * MAXEFC),
        ("efc_KBIP", (_d * 4) * MAXEFC),
        ("efc_vel", _d * MAXEFC), ("efc_aref", _d * MAXEFC), ("efc_b", _d * MAXEFC), ("efc_force", _d * MAXEFC),
        ("solver_niter", _i), ("solver_nls", _i), ("warning_bad", _i), ("used_warmstart", _i),
        ("solver_cost", _d),
        ("dof_parent", _i * NV), ("dof_Madr", _i * NV), ("dof_body", _i * NV), ("body_root", _i * NB),
        ("nM", _i),
        ("body_subtreemass", _d * NB),
        ("ls_evals_iter", _i * 8),
        ("ncon", _i), ("con_unsupported", _i), ("con_geom", _i * MAXCON), ("con_vert", _i * MAXCON),
        ("con_dist", _d * MAXCON), ("con_pos", (_d * 3) * MAXCON), ("con_frame", (_d * 9) * MAXCON),
        ("con_gap", _d * MAXCON),
    ]


_LIB: Optional[C.CDLL] = None
LIB_PATH = os.path.join(_HERE, "_build", "libso101_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "so101_oracle.c")
    hdr = os.path.join(_ROOT, "include", "so101_b200.h")
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    gcc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [gcc, "-O2", "-fPIC", "-fopenmp", "-ffp-contract=off", "-Wall", "-std=c11", "-D_GNU_SOURCE",
           "-shared", "-o", LIB_PATH, src, "-lm"]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return LIB_PATH


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.so101o_sizeof_data.restype = C.c_size_t
        if L.so101o_sizeof_data() != C.sizeof(OracleData):
            raise RuntimeError("OracleData mirror out of sync with so101_oracle.c")
        P = C.POINTER
        L.so101o_init.argtypes = [P(So101Tables), P(OracleData)]
        L.so101o_init.restype = C.c_int
        L.so101o_reset.argtypes = [P(So101Tables), P(OracleData)]
        L.so101o_forward.argtypes = [P(So101Tables), P(OracleData)]
        L.so101o_step.argtypes = [P(So101Tables), P(OracleData)]
        L.so101o_fullM.argtypes = [P(OracleData), P(_d)]
        L.so101o_uniform8.argtypes = [C.c_uint64, C.c_int64, C.c_uint32, C.c_uint32, P(_d)]
        L.so101o_rollout.argtypes = [P(So101Tables), P(So101CtrlSpec), C.c_int64, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int,
                                     P(C.c_int64)]
        L.so101o_rollout.restype = C.c_int64
        L.so101o_step_batch.argtypes = [P(So101Tables), C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.so101o_shoot.argtypes = [P(So101Tables), C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int,
                                   C.c_void_p, C.c_uint32, C.c_int]
        L.so101o_set_hulls.argtypes = [P(So101Hulls)]
        L.so101o_contact_probe.argtypes = [P(So101Tables), C.c_int64, C.c_void_p, C.c_void_p, C.c_int]
        L.so101o_num_threads.restype = C.c_int
        L.so101o_set_line_search.argtypes = [C.c_int]
        L.so101o_set_solver_start.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One-env MuJoCo-like `model`/`data` pair for stage-by-stage checks."""

    def __init__(self, tables: So101Tables):
        self.m = tables
        self.d = OracleData()
        rc = lib().so101o_init(C.byref(self.m), C.byref(self.d))
        if rc != 0:
            raise ValueError(f"so101o_init failed: {rc}")

    def reset(self):
        lib().so101o_reset(C.byref(self.m), C.byref(self.d))

    def forward(self):
        lib().so101o_forward(C.byref(self.m), C.byref(self.d))

    def step(self, n: int = 1):
        for _ in range(n):
            lib().so101o_step(C.byref(self.m), C.byref(self.d))

    def arr(self, name: str) -> np.ndarray:
        return np.ctypeslib.as_array(getattr(self.d, name))

    def set(self, name: str, v) -> None:
        a = self.arr(name)
        a[...] = v

    def full_M(self) -> np.ndarray:
        out = np.zeros((NV, NV))
        lib().so101o_fullM(C.byref(self.d), out.ctypes.data_as(C.POINTER(_d)))
        return out


def hulls_struct(h):
    """dict of hull arrays (tripwire.build_hulls / tables.builtin_hulls) -> (So101Hulls, keep-alive list)"""
    arrs = [np.ascontiguousarray(h["vert_start"], dtype=np.int32), np.ascontiguousarray(h["vert"], dtype=np.float64),
            np.ascontiguousarray(h["adj_start"], dtype=np.int32), np.ascontiguousarray(h["adj"], dtype=np.int32),
            np.ascontiguousarray(h["cube"], dtype=np.int32)]
    s = So101Hulls()
    s.ngeom, s.nvert, s.nadj, s.cube_res = len(arrs[0]) - 1, arrs[1].shape[0], arrs[3].shape[0], int(h["cube_res"])
    s.vert_start, s.vert, s.adj_start, s.adj, s.cube = [a.ctypes.data for a in arrs]
    return s, arrs


def set_hulls(h) -> None:
    """Load (or, with None, unload) the convex hulls: with them the oracle simulates table-plane contact."""
    if h is None:
        lib().so101o_set_hulls(None)
        return
    s, keep = hulls_struct(h)
    lib().so101o_set_hulls(C.byref(s))


def uniform8(seed: int, env: int, step: int, stream: int) -> np.ndarray:
    out = np.zeros(8)
    lib().so101o_uniform8(seed, env, step, stream, out.ctypes.data_as(C.POINTER(_d)))
    return out


def philox4x32_10(ctr, key) -> np.ndarray:
    c = np.asarray(ctr, dtype=np.uint32).copy()
    k = np.asarray(key, dtype=np.uint32).copy()
    out = np.zeros(4, dtype=np.uint32)
    lib().so101o_philox4x32_10(c.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


def make_spec(kind: int = 0, seed: int = 42, env_offset: int = 0, amp: float = 0.5, t_total: int = 200,
              freq_lo: float = 0.0025, freq_hi: float = 0.05, reset_lo: float = -0.3, reset_hi: float = 0.3,
              u: Optional[np.ndarray] = None) -> So101CtrlSpec:
    s = So101CtrlSpec()
    s.kind, s.t_total, s.seed, s.env_offset = kind, t_total, seed, env_offset
    s.amp, s.freq_lo, s.freq_hi, s.reset_lo, s.reset_hi = amp, freq_lo, freq_hi, reset_lo, reset_hi
    s.u = None if u is None else u.ctypes.data
    return s


def rollout(tables: So101Tables, spec: So101CtrlSpec, n: int, T: int, frame_skip: int = 10,
            qpos0: Optional[np.ndarray] = None, qvel0: Optional[np.ndarray] = None, flags: int = 0,
            nthreads: int = 0, want_rows: bool = True) -> Tuple[Optional[np.ndarray], np.ndarray, int]:
    """-> (rows [n,T+1,13] float64, final_state [n,18], total Newton iterations)."""
    rows = np.empty((n, T + 1, 13)) if want_rows else None
    final = np.empty((n, 18))
    if qpos0 is not None:
        qpos0 = np.ascontiguousarray(qpos0, dtype=np.float64)
        assert qpos0.shape == (n, NV)
    if qvel0 is not None:
        qvel0 = np.ascontiguousarray(qvel0, dtype=np.float64)
        assert qvel0.shape == (n, NV)
    it = C.c_int64(0)
    lib().so101o_rollout(C.byref(tables), C.byref(spec), n, T, frame_skip, _ptr(qpos0), _ptr(qvel0),
                         _ptr(rows), _ptr(final), flags, nthreads, C.byref(it))
    return rows, final, it.value


def step_batch(tables: So101Tables, state: np.ndarray, ctrl: np.ndarray, nsub: int = 1,
               qfrc_applied: Optional[np.ndarray] = None, nthreads: int = 0):
    """Teacher-forced steps: state [n,18], ctrl [n,6] -> (state_out [n,18], obs [n,8] f64, aux [n,4])."""
    state = np.ascontiguousarray(state, dtype=np.float64)
    ctrl = np.ascontiguousarray(ctrl, dtype=np.float64)
    n = state.shape[0]
    assert state.shape == (n, 18) and ctrl.shape == (n, NV)
    if qfrc_applied is not None:
        qfrc_applied = np.ascontiguousarray(qfrc_applied, dtype=np.float64)
    out = np.empty((n, 18))
    obs = np.empty((n, 8))
    aux = np.empty((n, 4))
    lib().so101o_step_batch(C.byref(tables), n, _ptr(state), _ptr(ctrl), _ptr(qfrc_applied), nsub,
                            _ptr(out), _ptr(obs), _ptr(aux), nthreads)
    return out, obs, aux


def contact_probe(tables: So101Tables, state: np.ndarray, nthreads: int = 0) -> np.ndarray:
    """state [n, 18] -> [n, 4] = (ncon, unsupported, smallest runner-up gap of the witness vertices, deepest dist)."""
    state = np.ascontiguousarray(state, dtype=np.float64)
    out = np.empty((state.shape[0], 4))
    lib().so101o_contact_probe(C.byref(tables), state.shape[0], _ptr(state), _ptr(out), nthreads)
    return out


def shoot(tables: So101Tables, state0: np.ndarray, U: np.ndarray, frame_skip: int = 10, flags: int = 0,
          nthreads: int = 0) -> np.ndarray:
    """U [H,5,B] float64 -> X [B,H+1,8] float32."""
    state0 = np.ascontiguousarray(state0, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    H, nu, B = U.shape
    assert nu == 5 and state0.shape == (18,)
    X = np.empty((B, H + 1, 8), dtype=np.float32)
    lib().so101o_shoot(C.byref(tables), _ptr(state0), _ptr(U), B, H, frame_skip, _ptr(X), flags, nthreads)
    return X


def set_line_search(mode: int) -> None:
    """0: MuJoCo's PrimalSearch (default).  1: the exact piece-walking search the CUDA kernels use."""
    lib().so101o_set_line_search(int(mode))


def set_solver_start(mode: int) -> None:
    """0: MuJoCo's warm start (default).  1: the per-dof prox start the CUDA kernels use."""
    lib().so101o_set_solver_start(int(mode))


def num_threads() -> int:
    return lib().so101o_num_threads()
