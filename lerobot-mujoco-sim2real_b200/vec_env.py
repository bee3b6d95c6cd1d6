"""SOARM101VecEnv — N independent SO-ARM101 environments stepped in lockstep on one B200.

Batched counterpart of `SOARM101Env` [REF SOARM101/SOARM101_Env.py:23-153]: same reset / step /
observation semantics per environment (frame_skip = round(dt / timestep) sub-steps of
mj_step per step, observation = float32 [ee_pos(3), qpos[0:5]] with the reference's one
sub-step ee lag), state resident in HBM as structure-of-arrays, all arithmetic in the CUDA
library behind include/so101_b200.h.  torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple, Union

import numpy as np
import torch

from . import _lib
from . import tables as T
from .tables import So101CtrlSpec, So101Tables

JOINT_NAMES = ["shoulder_pan", "shoulder_lift", "elbow_flex", "wrist_flex", "wrist_roll"]


def _dtype_code(dtype: Union[str, torch.dtype]) -> Tuple[int, torch.dtype]:
    if dtype in ("float64", "fp64", "f64", torch.float64, np.float64):
        return T.F64, torch.float64
    if dtype in ("float32", "fp32", "f32", torch.float32, np.float32):
        return T.F32, torch.float32
    raise ValueError(f"dtype must be float64 or float32, got {dtype!r}")


class Model:
    """Owns a `So101Model*` (compiled tables uploaded as kernel-parameter constants)."""

    def __init__(self, tables: So101Tables, hulls="auto"):
        """hulls: dict of hull arrays (tripwire.build_hulls), "auto" = the built-in ones when the tables carry the
        built-in scenes' contact geometry, None = none.  With hulls the batch SIMULATES table-plane contact; without,
        envs that would touch the table are only flagged (SO101_FLAG_TRIP_TABLE)."""
        self.tables = tables
        self._h = C.c_void_p()
        _lib.check(_lib.lib().so101_model_create(C.byref(tables), C.byref(self._h)))
        if isinstance(hulls, str):
            hulls = T.hulls_for(tables)
        self.has_contact = False
        if hulls is not None and tables.con_enabled:
            arrs = [np.ascontiguousarray(hulls["vert_start"], dtype=np.int32), np.ascontiguousarray(hulls["vert"], dtype=np.float64),
                    np.ascontiguousarray(hulls["adj_start"], dtype=np.int32), np.ascontiguousarray(hulls["adj"], dtype=np.int32),
                    np.ascontiguousarray(hulls["cube"], dtype=np.int32)]
            h = T.So101Hulls()
            h.ngeom, h.nvert, h.nadj, h.cube_res = len(arrs[0]) - 1, arrs[1].shape[0], arrs[3].shape[0], int(hulls["cube_res"])
            h.vert_start, h.vert, h.adj_start, h.adj, h.cube = [a.ctypes.data for a in arrs]
            _lib.check(_lib.lib().so101_model_set_hulls(self._h, C.byref(h)))     # copied by the library
            self.has_contact = True

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None and _lib._LIB is not None:
            _lib._LIB.so101_model_destroy(h)
            self._h = None


class SOARM101VecEnv:
    def __init__(self, xml_path: Optional[str] = None, num_envs: int = 1, dt: float = 0.02,
                 dtype: Union[str, torch.dtype] = "float64", device: Union[int, str, torch.device] = 0,
                 tables: Optional[So101Tables] = None, seed: int = 42, gravity_compensation: bool = False,
                 hulls="auto"):
        """xml_path: MJCF scene (compiled on the host once) — or pass pre-compiled `tables`.
        gravity_compensation=True: `step` first sets qfrc_applied = qfrc_bias of the current state, as the reference's
        control loops do before every env step [REF Koopman_MPC.py:119; SOARM101_Env.py:120 (commented out)]."""
        _lib.require_device()
        if tables is None:
            if xml_path is None:
                raise ValueError("give xml_path or tables")
            from .mjcf import attach_tripwire, compile_mjcf
            self.compiled = compile_mjcf(xml_path)
            attach_tripwire(self.compiled, xml_path)
            tables = self.compiled.tables
            if isinstance(hulls, str) and self.compiled.hulls is not None:
                hulls = self.compiled.hulls
        else:
            self.compiled = None
        self.tables = tables
        self.num_envs = int(num_envs)
        self.dtype_code, self.torch_dtype = _dtype_code(dtype)
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.So101Error("SOARM101VecEnv runs on CUDA devices only (no CPU fallback)")
        self.device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        # [REF SOARM101_Env.py:39-40]
        self.frame_skip = max(1, int(np.round(dt / tables.timestep)))
        self.dt = tables.timestep * self.frame_skip
        self.joint_names = list(JOINT_NAMES)
        self.udim, self.xdim, self.max_speed = T.NU_ENV, T.NOBS, 0.5
        self.seed = int(seed)
        self._episode = 0
        self.gravity_compensation = bool(gravity_compensation)
        self.model = Model(tables, hulls)
        L = _lib.lib()
        nbytes = L.so101_batch_state_bytes(self.num_envs, self.dtype_code)
        # caller-owned state: one torch allocation, bound to the batch for its lifetime
        self._state = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self._h = C.c_void_p()
        _lib.check(L.so101_batch_create(self.model._h, self.num_envs, self.dtype_code, self.device_index,
                                         self._state.data_ptr(), C.byref(self._h)))
        self._obs = torch.empty((T.NOBS, self.num_envs), dtype=torch.float32, device=self.device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None and _lib._LIB is not None:
            _lib._LIB.so101_batch_destroy(h)
            self._h = None

    # ---- helpers ----------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _soa(self, x, rows: int) -> torch.Tensor:
        """[N, >=rows] array-like (or [>=rows] when N == 1) -> contiguous [rows, N] tensor, batch dtype."""
        t = torch.as_tensor(x)
        if t.dim() == 1:
            t = t.reshape(1, -1)
        if t.dim() != 2 or t.shape[0] != self.num_envs or t.shape[1] < rows:
            raise ValueError(f"expected shape ({self.num_envs}, >={rows}), got {tuple(t.shape)}")
        return t[:, :rows].t().to(device=self.device, dtype=self.torch_dtype).contiguous()

    # ---- reference API ----------------------------------------------------------------------------
    def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None):
        """-> (obs [N, 8] float32 device tensor, {}).  [REF SOARM101_Env.py:77-106]"""
        L = _lib.lib()
        if seed is not None:
            self.seed, self._episode = int(seed), 0
        if options and "initial_state" in options:
            init = torch.as_tensor(options["initial_state"]).to(self.device, self.torch_dtype).reshape(self.num_envs, -1)
            q = torch.as_tensor(list(self.tables.qpos0), dtype=self.torch_dtype, device=self.device)
            q = q.reshape(T.NV, 1).repeat(1, self.num_envs)      # mj_resetData: qpos = qpos0 (the gripper keeps it)
            v = torch.zeros_like(q)
            q[:5] = init[:, :5].t()
            v[:5] = init[:, 5:10].t()
            _lib.check(L.so101_batch_reset(self._h, q.data_ptr(), v.data_ptr(), self._obs.data_ptr(), self._stream()))
        else:
            # fresh draw per episode: Philox stream keyed (seed + episode, env)
            _lib.check(L.so101_batch_reset_random(self._h, self.seed + 1000003 * self._episode, 0, -0.3, 0.3,
                                                  self._obs.data_ptr(), self._stream()))
            self._episode += 1
        return self._obs.t(), {}

    def step(self, action):
        """action [N, 5] -> (obs [N, 8], 0.0, False, False, {}).  [REF SOARM101_Env.py:108-142]"""
        u = self._soa(action, T.NU_ENV)
        self.step_soa(u)
        return self._obs.t(), 0.0, False, False, {}

    def step_soa(self, ctrl_soa: torch.Tensor, n_substeps: Optional[int] = None) -> torch.Tensor:
        """Zero-copy step: ctrl_soa [5 or 6, N] contiguous, batch dtype.  Returns obs [8, N].  With
        gravity_compensation the launch itself sets qfrc_applied = qfrc_bias of the state it starts from (one fused
        launch instead of forward + copy + step)."""
        assert ctrl_soa.is_contiguous() and ctrl_soa.dtype == self.torch_dtype and ctrl_soa.shape[1] == self.num_envs
        ns = self.frame_skip if n_substeps is None else n_substeps
        _lib.check(_lib.lib().so101_batch_step_flags(self._h, ctrl_soa.data_ptr(), ctrl_soa.shape[0], ns,
                                                     self._obs.data_ptr(),
                                                     T.ROLL_GRAVCOMP_HOLD if self.gravity_compensation else 0,
                                                     self._stream()))
        return self._obs

    def forward(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """mj_forward at the current state -> (obs [N, 8], qfrc_bias [N, 6])."""
        bias = torch.empty((T.NV, self.num_envs), dtype=self.torch_dtype, device=self.device)
        _lib.check(_lib.lib().so101_batch_forward(self._h, self._obs.data_ptr(), bias.data_ptr(), self._stream()))
        return self._obs.t(), bias.t()

    # ---- fused paths --------------------------------------------------------------------------------
    def make_spec(self, input_type: str = "random", seed: Optional[int] = None, env_offset: int = 0,
                  u: Optional[torch.Tensor] = None, amp: float = 0.5, t_total: int = 200) -> So101CtrlSpec:
        s = So101CtrlSpec()
        s.kind = T.CTRL_KINDS[input_type]
        s.t_total = t_total
        s.seed = self.seed if seed is None else int(seed)
        s.env_offset = env_offset
        s.amp, s.freq_lo, s.freq_hi = amp, 0.0025, 0.05     # [REF SOARM101_DataCollection.py:97-103]
        s.reset_lo, s.reset_hi = -0.3, 0.3                  # [REF SOARM101_Env.py:95]
        s.u = u.data_ptr() if u is not None else None
        return s

    def rollout(self, steps: int, input_type: str = "random", seed: Optional[int] = None, env_offset: int = 0,
                u: Optional[torch.Tensor] = None, flags: int = 0, out: Optional[torch.Tensor] = None,
                frame_skip: Optional[int] = None, out_ptr: Optional[int] = None) -> Optional[torch.Tensor]:
        """One launch = `generate_physics_based_data(num_envs, steps, input_type)`
        [REF SOARM101_DataCollection.py:90-136].  -> rows [N, steps+1, 13] float64 on device.
        out_ptr: raw device address the rows are written to instead of a torch tensor - e.g. this rank's slice of the
        dataset buffer that lives on another GPU of the node (`sharding.SharedRows`); returns None then."""
        row_dtype = torch.float32 if flags & T.ROLL_ROWS_F32 else torch.float64
        if out_ptr is None:
            if out is None:
                out = torch.empty((self.num_envs, steps + 1, T.ROW), dtype=row_dtype, device=self.device)
            assert out.is_contiguous() and out.dtype == row_dtype and out.shape == (self.num_envs, steps + 1, T.ROW)
            out_ptr = out.data_ptr()
        else:
            assert out is None, "give out or out_ptr, not both"
        if u is not None:
            assert u.is_contiguous() and u.dtype == self.torch_dtype and u.shape == (steps + 1, T.NU_ENV, self.num_envs)
        spec = self.make_spec(input_type, seed, env_offset, u)
        _lib.check(_lib.lib().so101_batch_rollout(self._h, C.byref(spec), steps, frame_skip or self.frame_skip,
                                                  out_ptr, flags, self._stream()))
        return out

    def rollout_host(self, steps: int, input_type: str = "random", seed: Optional[int] = None, env_offset: int = 0,
                     u_host: Optional[torch.Tensor] = None, qpos0_host: Optional[torch.Tensor] = None,
                     out_host: Optional[torch.Tensor] = None, flags: int = 0) -> torch.Tensor:
        """`generate_physics_based_data` with HOST buffers through one C-ABI call
        (`so101_batch_rollout_host`): controls [steps+1, 5, N] / initial angles [6, N] in, rows out."""
        row_dtype = torch.float32 if flags & T.ROLL_ROWS_F32 else torch.float64
        if out_host is None:
            out_host = torch.empty((self.num_envs, steps + 1, T.ROW), dtype=row_dtype).pin_memory()
        assert out_host.device.type == "cpu" and out_host.is_contiguous() and out_host.dtype == row_dtype
        assert out_host.shape == (self.num_envs, steps + 1, T.ROW)
        spec = self.make_spec(input_type, seed, env_offset)
        if input_type == "tensor":
            assert u_host is not None and u_host.device.type == "cpu" and u_host.is_contiguous()
            assert u_host.dtype == self.torch_dtype and u_host.shape == (steps + 1, T.NU_ENV, self.num_envs)
            spec.u = u_host.data_ptr()
        q0 = None
        if qpos0_host is not None:
            assert qpos0_host.device.type == "cpu" and qpos0_host.is_contiguous()
            assert qpos0_host.dtype == self.torch_dtype and qpos0_host.shape == (T.NV, self.num_envs)
            q0 = qpos0_host.data_ptr()
        _lib.check(_lib.lib().so101_batch_rollout_host(self._h, C.byref(spec), q0, steps, self.frame_skip,
                                                       out_host.data_ptr(), flags, self._stream()))
        return out_host

    def rollout_discard(self, steps: int, input_type: str = "random", seed: Optional[int] = None,
                        flags: int = 0, frame_skip: Optional[int] = None) -> None:
        """Rollout without writing rows (throughput measurements, warm-up)."""
        spec = self.make_spec(input_type, seed)
        _lib.check(_lib.lib().so101_batch_rollout(self._h, C.byref(spec), steps, frame_skip or self.frame_skip,
                                                  None, flags, self._stream()))

    def shoot(self, state0, U: torch.Tensor, flags: int = 0, frame_skip: Optional[int] = None) -> torch.Tensor:
        """num_envs control sequences U [H, 5, N] from one shared state0 (18 doubles: qpos, qvel,
        qacc_warmstart) -> X [N, H+1, 8] float32."""
        s0 = np.ascontiguousarray(np.asarray(state0, dtype=np.float64).reshape(18))
        assert U.is_contiguous() and U.dtype == self.torch_dtype and U.shape[1:] == (T.NU_ENV, self.num_envs)
        H = U.shape[0]
        X = torch.empty((self.num_envs, H + 1, T.NOBS), dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().so101_batch_shoot(self._h, s0.ctypes.data, U.data_ptr(), H,
                                                frame_skip or self.frame_skip, X.data_ptr(), flags, self._stream()))
        return X

    # ---- state access ---------------------------------------------------------------------------------
    def get_state(self) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """-> qpos, qvel, qacc_warmstart, each [N, 6] (views of fresh [6, N] copies)."""
        out = [torch.empty((T.NV, self.num_envs), dtype=self.torch_dtype, device=self.device) for _ in range(3)]
        _lib.check(_lib.lib().so101_batch_get_state(self._h, out[0].data_ptr(), out[1].data_ptr(),
                                                    out[2].data_ptr(), self._stream()))
        return out[0].t(), out[1].t(), out[2].t()

    def set_state(self, qpos=None, qvel=None, qacc_warmstart=None) -> None:
        ts = [None if x is None else self._soa(x, T.NV) for x in (qpos, qvel, qacc_warmstart)]
        _lib.check(_lib.lib().so101_batch_set_state(self._h, *[None if t is None else t.data_ptr() for t in ts],
                                                    self._stream()))
        torch.cuda.current_stream(self.device).synchronize()  # keep `ts` alive until copied

    def set_qfrc_applied(self, qfrc) -> None:
        t = self._soa(qfrc, T.NV)
        _lib.check(_lib.lib().so101_batch_set_qfrc_applied(self._h, t.data_ptr(), self._stream()))
        torch.cuda.current_stream(self.device).synchronize()

    def flags(self) -> torch.Tensor:
        f = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib().so101_batch_get_flags(self._h, f.data_ptr(), self._stream()))
        return f

    def clear_flags(self) -> None:
        _lib.check(_lib.lib().so101_batch_clear_flags(self._h, self._stream()))

    def set_option(self, option: int, value: int) -> None:
        """Explicit experiment options (tables.OPT_*): kernel family, block size, host pipeline depth.  The library
        never reads environment variables."""
        _lib.check(_lib.lib().so101_batch_set_option(self._h, int(option), int(value)))

    def stats(self) -> Dict[str, int]:
        """Counters since the last call: physics steps, Newton iterations, line-search evals, limit steps."""
        buf = (C.c_uint64 * 4)()
        _lib.check(_lib.lib().so101_batch_stats(self._h, buf, self._stream()))
        return {"physics_steps": buf[0], "newton_iters": buf[1], "ls_evals": buf[2], "limit_steps": buf[3]}

    def close(self) -> None:
        pass


def fma_peak_tflops(dtype: str = "float64", device: int = 0) -> float:
    code, _ = _dtype_code(dtype)
    out = C.c_double(0)
    _lib.check(_lib.lib().so101_fma_peak(code, device, C.byref(out)))
    return out.value
