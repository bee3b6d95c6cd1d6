"""Drop-in `SOARM101Env` backed by the B200 stepper (one environment = a batch of 1).

Mirrors the reference class [REF SOARM101/SOARM101_Env.py:8-153] attribute for attribute:
constructor `(xml_path, dt=0.02, render_mode=False)`, `reset(seed, options) -> (obs, {})`,
`step(action) -> (obs, 0.0, False, False, {})`, `frame_skip`, `dt`, `joint_names`, `joint_ids`,
`ee_site_id`, `udim`, `xdim`, `max_speed`, `action_space`, `observation_space`, `model`, `data`,
`viewer`, `render()`, `close()`.  Observations are float32 `[ee_pos(3), qpos[0:5]]` with the
reference's one-sub-step lag of ee_pos behind qpos.

`env.model` / `env.data` are light views that cover what the reference scripts touch
[REF Koopman_MPC.py:65-71,89,119,169-186]: `model.opt.timestep`, `model.key_qpos/key_ctrl`,
`model.joint(name).id`, `model.nq/nv/nu`, `data.qpos`, `data.qvel`, `data.ctrl`,
`data.qfrc_applied`, `data.qfrc_bias`, `data.site_xpos`, `data.time`.  The interactive viewer is
CPU-MuJoCo only and out of scope: `render_mode=True` raises.

For throughput use `SOARM101VecEnv` (vec_env.py); this class exists so that existing callers
run unchanged.  Inputs/outputs are host numpy arrays: every `step` is one H2D copy, one kernel,
one D2H copy through `so101_batch_step_host`.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib
from . import tables as T
from .vec_env import JOINT_NAMES, SOARM101VecEnv

try:  # gymnasium is optional: the reference subclasses gym.Env but uses nothing else from it
    import gymnasium as gym
    from gymnasium import spaces
    _EnvBase = gym.Env
    _Box = spaces.Box
except Exception:  # pragma: no cover - depends on the image
    gym = None

    class _EnvBase:  # minimal stand-in: seeding as gymnasium.Env.reset(seed=...) does
        np_random: np.random.Generator = np.random.default_rng()

        def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None):
            if seed is not None:
                self.np_random = np.random.default_rng(seed)

    class _Box:
        def __init__(self, low, high, shape, dtype=np.float32):
            self.low = np.full(shape, low, dtype=dtype)
            self.high = np.full(shape, high, dtype=dtype)
            self.shape, self.dtype = tuple(shape), np.dtype(dtype)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return np.random.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class _Opt:
    def __init__(self, t: T.So101Tables):
        self.timestep = t.timestep
        self.gravity = np.array(t.gravity[:])
        self.tolerance, self.iterations = t.tolerance, t.iterations


class _NamedId:
    def __init__(self, name: str, idx: int):
        self.name, self.id = name, idx


class ModelView:
    """The slice of `mujoco.MjModel` the reference scripts read."""

    def __init__(self, t: T.So101Tables, joint_names, site_names, key_names):
        self._t = t
        self.opt = _Opt(t)
        self.nq = self.nv = t.nv
        self.nu = t.nu
        self.nbody = t.nbody
        self._joint_names, self._site_names, self._key_names = list(joint_names), list(site_names), list(key_names)
        self.key_qpos = np.array(t.key_qpos[:]).reshape(1, -1)
        self.key_ctrl = np.array(t.key_ctrl[:]).reshape(1, -1)
        self.nkey = 1
        self.jnt_range = T.as_np(t.jnt_range)
        self.actuator_ctrlrange = T.as_np(t.act_ctrlrange)

    def joint(self, name: str) -> _NamedId:
        return _NamedId(name, self._joint_names.index(name))

    def site(self, name: str) -> _NamedId:
        return _NamedId(name, self._site_names.index(name))

    def key(self, name: str) -> _NamedId:
        return _NamedId(name, self._key_names.index(name))


class _DeviceVector:
    """numpy-like view of one per-env vector held on the device: reads fetch, writes push."""

    def __init__(self, getter, setter):
        self._get, self._set = getter, setter

    def __array__(self, dtype=None, copy=None):
        a = self._get()
        return a.astype(dtype) if dtype is not None else a

    def __getitem__(self, idx):
        return self._get()[idx]

    def __setitem__(self, idx, value):
        a = self._get()
        a[idx] = value
        self._set(a)

    def __len__(self):
        return len(self._get())

    def copy(self):
        return self._get().copy()

    def __repr__(self):
        return repr(self._get())


class DataView:
    """The slice of `mujoco.MjData` the reference scripts touch, backed by device state."""

    def __init__(self, env: "SOARM101Env"):
        object.__setattr__(self, "_env", env)

    # --- state vectors (write-through) ---
    def _state(self, k: int) -> np.ndarray:
        return self._env._vec.get_state()[k].cpu().numpy().reshape(-1).astype(np.float64)

    def _set_state(self, k: int, a: np.ndarray) -> None:
        args = [None, None, None]
        args[k] = np.asarray(a, dtype=np.float64).reshape(1, -1)
        self._env._vec.set_state(*args)

    @property
    def qpos(self):
        return _DeviceVector(lambda: self._state(0), lambda a: self._set_state(0, a))

    @property
    def qvel(self):
        return _DeviceVector(lambda: self._state(1), lambda a: self._set_state(1, a))

    @property
    def qacc_warmstart(self):
        return _DeviceVector(lambda: self._state(2), lambda a: self._set_state(2, a))

    @property
    def ctrl(self):
        return _DeviceVector(lambda: self._env._ctrl.copy(), self._env._set_ctrl)

    @property
    def qfrc_applied(self):
        return _DeviceVector(lambda: self._env._qfrc_applied.copy(), self._env._set_qfrc_applied)

    @property
    def qfrc_bias(self) -> np.ndarray:
        _, bias = self._env._vec.forward()
        return bias.cpu().numpy().reshape(-1).astype(np.float64)

    @property
    def site_xpos(self) -> np.ndarray:
        """[nsite, 3]; only the observation site (`gripperframe`) is tracked, other rows are NaN."""
        out = np.full((len(self._env._site_names), 3), np.nan)
        out[self._env.ee_site_id] = self._env._last_ee
        return out

    @property
    def time(self) -> float:
        return self._env._time


class SOARM101Env(_EnvBase):
    metadata = {"render_modes": ["human"]}

    def __init__(self, xml_path: str, dt: float = 0.02, render_mode=False, dtype: str = "float64",
                 device: int = 0, gravity_compensation: bool = False):
        """gravity_compensation=True: qfrc_applied = qfrc_bias before every env step, with qfrc_bias evaluated AT THE STATE
        THE STEP STARTS FROM - the semantics of the reference's control loop [REF Koopman_MPC.py:119-126: the assignment
        follows an explicit mj_forward], which is what `data.qfrc_bias` returns here.  The line the reference keeps
        commented out inside step() (`self.data.qfrc_applied[:] = self.data.qfrc_bias[:]`, [REF SOARM101_Env.py:120]) would
        read the qfrc_bias left by the LAST SUB-STEP's forward pass, i.e. one sub-step (2 ms) behind, like ee_pos (F6); the
        difference is O(h) in a gravity torque and no test distinguishes the two on the oracle either (it recomputes, too).
        The reference's shipped Koopman model was trained on data generated with gravity compensation."""
        super().__init__()
        self.gravity_compensation = bool(gravity_compensation)
        if render_mode:
            raise NotImplementedError("the interactive MuJoCo viewer is out of scope of the B200 path "
                                      "(use the reference env for rendering)")
        from . import mjcf
        self._compiled = mjcf.compile_mjcf(xml_path)   # raises FileNotFoundError like the reference
        self.contact_tables = mjcf.attach_tripwire(self._compiled, xml_path)   # "none": contacts go unnoticed (stderr says why)
        t = self._compiled.tables
        self._vec = SOARM101VecEnv(tables=t, num_envs=1, dt=dt, dtype=dtype, device=device)
        self.frame_skip = self._vec.frame_skip
        self.dt = self._vec.dt
        print(f"环境控制步长(dt): {self.dt:.4f}s (执行 {self.frame_skip} 个物理步骤)")
        self._site_names = self._compiled.site_names
        self.model = ModelView(t, self._compiled.joint_names, self._site_names, self._compiled.key_names)
        self.data = DataView(self)
        self.joint_names = list(JOINT_NAMES)
        self.joint_ids = [self.model.joint(name).id for name in self.joint_names]
        if self.joint_ids != [0, 1, 2, 3, 4]:
            raise ValueError("the arm joints must be joints 0..4 of the model")
        self.ee_site_id = self._compiled.site_id("gripperframe")
        self.udim = 5
        self.max_speed = 0.5
        self.action_space = _Box(low=-self.max_speed, high=self.max_speed, shape=(self.udim,), dtype=np.float32)
        self.xdim = 8
        self.observation_space = _Box(low=-np.inf, high=np.inf, shape=(self.xdim,), dtype=np.float32)
        self.render_mode = render_mode
        self.viewer = None
        self._np_dtype = np.float64 if dtype in ("float64", "fp64", "f64") else np.float32
        self._ctrl = np.zeros(T.NV)
        self._qfrc_applied = np.zeros(T.NV)
        self._last_ee = np.zeros(3)
        self._time = 0.0
        self._obs_host = np.zeros((T.NOBS, 1), dtype=np.float32)
        self._flags_host = np.zeros(1, dtype=np.uint32)

    # ---- internals ------------------------------------------------------------------------------
    def _set_ctrl(self, a) -> None:
        self._ctrl = np.asarray(a, dtype=np.float64).reshape(T.NV).copy()

    def _set_qfrc_applied(self, a) -> None:
        self._qfrc_applied = np.asarray(a, dtype=np.float64).reshape(T.NV).copy()
        self._vec.set_qfrc_applied(self._qfrc_applied.reshape(1, -1))

    def _obs(self) -> np.ndarray:
        o = self._obs_host[:, 0].copy()
        self._last_ee = o[:3].astype(np.float64)
        return o

    def _get_state(self) -> np.ndarray:
        """[REF SOARM101_Env.py:69-75]: concat(ee_pos, qpos[0:5]) as float32 (ee first)."""
        return self._obs()

    # ---- reference API --------------------------------------------------------------------------
    def reset(self, seed: Optional[int] = None, options: Optional[Dict] = None) -> Tuple[np.ndarray, Dict]:
        super().reset(seed=seed)
        if options and "initial_state" in options:
            initial_qpos = np.asarray(options["initial_state"][:5], dtype=np.float64)
            initial_qvel = np.asarray(options["initial_state"][5:10], dtype=np.float64)
        else:
            initial_qpos = self.np_random.uniform(low=-0.3, high=0.3, size=self.udim)
            initial_qvel = np.zeros(self.udim)
        q = np.array(self._compiled.tables.qpos0[:], dtype=self._np_dtype).reshape(T.NV, 1)
        v = np.zeros((T.NV, 1), dtype=self._np_dtype)
        q[:5, 0] = initial_qpos
        v[:5, 0] = initial_qvel
        self._ctrl[:] = 0.0           # mj_resetData clears ctrl and qfrc_applied
        self._qfrc_applied[:] = 0.0
        self._time = 0.0
        _lib.check(_lib.lib().so101_batch_reset_host(self._vec._h, q.ctypes.data, v.ctypes.data,
                                                     self._obs_host.ctypes.data, self._vec._stream()))
        return self._obs(), {}

    def step(self, action: np.ndarray) -> Tuple[np.ndarray, float, bool, bool, Dict]:
        if self.gravity_compensation:
            self.data.qfrc_applied[:] = self.data.qfrc_bias[:]
        target_velocity = np.asarray(action, dtype=np.float64).reshape(-1)[: self.udim]
        self._ctrl[: self.udim] = target_velocity          # ctrl[5] keeps its value (0 after reset)
        u = np.ascontiguousarray(self._ctrl.reshape(T.NV, 1), dtype=self._np_dtype)
        _lib.check(_lib.lib().so101_batch_step_host(self._vec._h, u.ctypes.data, T.NV, self.frame_skip,
                                                    self._obs_host.ctypes.data, self._flags_host.ctypes.data,
                                                    self._vec._stream()))
        self._time += self.dt
        # the reference returns an empty info dict; here it carries the env's status word whenever it is not clean
        # (tables.FLAG_*: a contact the simulator does not model, a state blow-up that mj_step would have reset, ...)
        fl = int(self._flags_host[0])
        return self._obs(), 0.0, False, False, ({"flags": fl} if fl & T.FLAG_ABNORMAL else {})

    def forward(self) -> np.ndarray:
        """mujoco.mj_forward(model, data) for callers that invoke it explicitly
        [REF Koopman_MPC.py:90,126]: refreshes the observation at the current state."""
        obs, _ = self._vec.forward()
        o = obs.cpu().numpy().reshape(-1)
        self._last_ee = o[:3].astype(np.float64)
        return o

    def render(self):
        pass

    def close(self):
        self.viewer = None
