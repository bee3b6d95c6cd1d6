"""Cartesian reference curves -> joint-angle tracks on the device.

Host-side mirror of the reference's `CartesianTrajectoryGenerator` [REF control/TrajectoryGenerator.py:10-210]: same
constructor arguments, same `generate(traj_name, target_orientation)` return tuple `(xyz [P,3], joint_angles
[P,num_joints], time_vector [P])`, same failure behaviour (a failed way-point repeats the previous answer, a failure
at the first way-point raises RuntimeError).  The per-way-point `dm_control` inverse kinematics
[REF control/TrajectoryGenerator.py:81-116] runs in one CUDA launch through the C ABI (`so101_ik_track`), for any
number of tracks at once (`solve_tracks`, `generate_batch`) — that is what makes `TrajectoryGenerator`-driven dataset
generation (BASELINE.json config 4) a device-side job: tracks come out as `[n, P, 6]` joint targets ready to feed the
batched stepper's control tensor.

Only torch plumbing and the C ABI are used; there is no CPU path and nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from . import tables as T
from .tables import So101IkParams, So101Tables
from .vec_env import JOINT_NAMES, Model

IK_SUCCESS, IK_ABORTED = 1, 2


def reference_curve(traj_name: str = "Fig8", idx: int = 1, time_horizon: float = 60, time_steps_per_sec: int = 5,
                    traj_scale: float = 0.5, center: Optional[Sequence[float]] = None,
                    radius: float = 0.1) -> Tuple[np.ndarray, np.ndarray]:
    """Way-points of the reference's curves [REF control/TrajectoryGenerator.py:136-170]: `Fig8` (lemniscate of
    half-widths 0.2*traj_scale) or `Circle` (radius 0.1) in the y-z plane at x = 0.4 (idx == 1) or the x-y plane at
    z = 0.2 (otherwise), parameter 1.6 + 0.02*linspace(0, 5*time_horizon, P), P = time_steps_per_sec*time_horizon.
    `center` overrides the curve centre ((0.4, 0, 0.2) resp. (0.3, 0, 0.2)).  Returns (xyz [P,3], time_vector [P])."""
    steps = int(time_steps_per_sec * time_horizon)
    time_vector = np.linspace(0, time_horizon, steps)
    t = 1.6 + 0.02 * np.linspace(0, time_horizon * 5, len(time_vector))
    c = np.array(center if center is not None else ((0.4, 0.0, 0.2) if idx == 1 else (0.3, 0.0, 0.2)), dtype=np.float64)
    if traj_name == "Fig8":
        a = b = 0.2 * traj_scale
        den = 1 + np.sin(t) ** 2
        u, v = 2 * a * np.sin(t) * np.cos(t) / den, b * np.cos(t) / den    # (in-plane "long" axis, y)
        if idx == 1:
            xyz = np.stack([c[0] * np.ones_like(t), c[1] + v, c[2] + u], axis=1)
        else:
            xyz = np.stack([c[0] + u, c[1] + v, c[2] * np.ones_like(t)], axis=1)
    elif traj_name == "Circle":
        if idx == 1:
            xyz = np.stack([c[0] * np.ones_like(t), c[1] + radius * np.cos(t), c[2] + radius * np.sin(t)], axis=1)
        else:
            xyz = np.stack([c[0] + radius * np.cos(t), c[1] + radius * np.sin(t), c[2] * np.ones_like(t)], axis=1)
    else:
        raise ValueError(f"未知的轨迹名称: {traj_name}")   # reference text
    return xyz, time_vector


class CartesianTrajectoryGenerator:
    def __init__(self, model_path: Optional[str] = None, ee_site_name: str = "gripperframe", num_joints: int = 5,
                 idx: int = 1, time_horizon: float = 60, time_steps_per_sec: int = 5,
                 tables: Optional[So101Tables] = None, device: Union[int, str, torch.device] = 0):
        """model_path / ee_site_name / num_joints / idx / time_horizon / time_steps_per_sec: as the reference's
        constructor [REF control/TrajectoryGenerator.py:15-43].  `tables`: pre-compiled scene instead of `model_path`."""
        _lib.require_device()
        self.idx = idx
        self.time_horizon = time_horizon
        self.time_steps = time_steps_per_sec * time_horizon
        self.time_steps_per_sec = time_steps_per_sec
        self.time_vector = np.linspace(0, self.time_horizon, int(self.time_steps))
        self.model_path = model_path
        self.ee_site_name = ee_site_name
        self.num_joints = int(num_joints)
        self.traj_scale = 0.5
        if tables is None:
            if model_path is None:
                raise ValueError("give model_path or tables")
            from .mjcf import MjcfError, compile_mjcf
            try:
                tables = compile_mjcf(model_path, site_name=ee_site_name).tables
            except MjcfError as exc:
                raise ValueError(f"错误: 在模型中找不到名为 '{ee_site_name}' 的 site。请检查XML文件。") from exc
        self.tables = tables
        self.joint_names = list(JOINT_NAMES)[:5]      # [REF control/TrajectoryGenerator.py:56]
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.So101Error("CartesianTrajectoryGenerator runs on CUDA devices only (no CPU fallback)")
        self.device_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.model = Model(tables)
        # the call's arguments [REF control/TrajectoryGenerator.py:96-107]; the rest are dm_control's defaults
        self.ik_params = So101IkParams(tol=1e-6, rot_weight=0.5, reg_strength=1e-2, max_steps=100, dof_mask=0x1F)
        self.qpos = np.zeros(T.NV)                    # physics.data.qpos of the reference's IK scratch model

    # ---- the batched solver -------------------------------------------------------------------------------------
    def solve_tracks(self, xyz, target_quat=None, q0=None, params: Optional[So101IkParams] = None,
                     return_err: bool = False):
        """xyz: [n, P, 3] way-points; target_quat: None (position only), [4] or [n, 4] (w, x, y, z) held along each track;
        q0: None (the model's qpos0), [6] or [n, 6] start joint vectors.
        Returns (q [n, P, 6] float64, status [n, P] int32) on the device (+ err_norm [n, P] if return_err);
        status bit 0 = success, bit 1 = track aborted (first way-point unsolvable), bits 8.. = iterations."""
        x = torch.as_tensor(xyz, dtype=torch.float64, device=self.device)
        if x.dim() == 2:
            x = x.unsqueeze(0)
        if x.dim() != 3 or x.shape[2] != 3:
            raise ValueError(f"xyz must be [n, P, 3], got {tuple(x.shape)}")
        n, P = int(x.shape[0]), int(x.shape[1])
        x_soa = x.permute(1, 2, 0).contiguous()                        # [P][3][n]

        def per_track(v, width, name):
            if v is None:
                return None
            t = torch.as_tensor(v, dtype=torch.float64, device=self.device)
            if t.dim() == 1:
                t = t.unsqueeze(0).expand(n, -1)
            if tuple(t.shape) != (n, width):
                raise ValueError(f"{name} must be [{width}] or [{n}, {width}], got {tuple(t.shape)}")
            return t.t().contiguous()                                  # [width][n]

        quat_soa = per_track(target_quat, 4, "target_quat")
        q0_soa = per_track(q0, T.NV, "q0")
        q_out = torch.empty((P, T.NV, n), dtype=torch.float64, device=self.device)
        status = torch.empty((P, n), dtype=torch.int32, device=self.device)
        err = torch.empty((P, n), dtype=torch.float64, device=self.device) if return_err else None
        prm = params if params is not None else self.ik_params
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().so101_ik_track(
                self.model._h, C.byref(prm), x_soa.data_ptr(),
                quat_soa.data_ptr() if quat_soa is not None else None,
                q0_soa.data_ptr() if q0_soa is not None else None, P, n, self.device_index,
                q_out.data_ptr(), status.data_ptr(), err.data_ptr() if err is not None else None,
                torch.cuda.current_stream(self.device).cuda_stream))
        out = (q_out.permute(2, 0, 1), status.t())
        return out + (err.t(),) if return_err else out

    def curves_on_device(self, traj_names: Sequence[str], idx: Optional[Sequence[int]] = None,
                         traj_scale: Optional[Sequence[float]] = None, centers=None, radius: float = 0.1) -> torch.Tensor:
        """traj_names: sequence of 'Fig8' / 'Circle', or a bool array (True = Fig8).
        `reference_curve` for n tracks at once as device tensor ops (no per-track host loop, nothing uploaded but
        the per-track parameters): -> xyz [n, P, 3] float64.  Same formulas; sin/cos come from the device math library,
        so way-points agree with the numpy version to ~1e-16, not bitwise."""
        n = len(traj_names)
        dev = self.device
        if isinstance(traj_names, (np.ndarray, torch.Tensor)) and traj_names.dtype in (np.bool_, torch.bool):
            fig8 = torch.as_tensor(traj_names, device=dev)[:, None]          # True = Fig8, False = Circle (large batches)
        else:
            for nm in traj_names:
                if nm not in ("Fig8", "Circle"):
                    raise ValueError(f"未知的轨迹名称: {nm}")   # reference text
            fig8 = torch.tensor([nm == "Fig8" for nm in traj_names], device=dev)[:, None]
        ix = torch.as_tensor(self.idx if idx is None else idx, device=dev).expand(n) if idx is None \
            else torch.as_tensor(idx, device=dev)
        yz = (ix == 1)[:, None]
        sc = torch.full((n,), float(self.traj_scale), dtype=torch.float64, device=dev) if traj_scale is None \
            else torch.as_tensor(traj_scale, dtype=torch.float64, device=dev)
        default_c = torch.where(yz, torch.tensor([[0.4, 0.0, 0.2]], dtype=torch.float64, device=dev),
                                torch.tensor([[0.3, 0.0, 0.2]], dtype=torch.float64, device=dev))
        c = default_c if centers is None else torch.as_tensor(centers, dtype=torch.float64, device=dev)
        P = int(self.time_steps_per_sec * self.time_horizon)
        t = 1.6 + 0.02 * torch.linspace(0, self.time_horizon * 5, P, dtype=torch.float64, device=dev)[None]   # [1, P]
        s_, c_ = torch.sin(t), torch.cos(t)
        den = 1 + s_ ** 2
        a = (0.2 * sc)[:, None]
        u, v = 2 * a * s_ * c_ / den, a * c_ / den          # Fig8: (long in-plane axis, y)
        # Circle: y-z plane -> y = r cos, z = r sin; x-y plane -> x = r cos, y = r sin
        cu = radius * c_.expand(n, P)
        su = radius * s_.expand(n, P)
        x = torch.where(yz, c[:, 0:1].expand(n, P), c[:, 0:1] + torch.where(fig8, u, cu))
        y = c[:, 1:2] + torch.where(fig8, v, torch.where(yz, cu, su))
        z = torch.where(yz, c[:, 2:3] + torch.where(fig8, u, su), c[:, 2:3].expand(n, P))
        return torch.stack([x, y, z], dim=2).contiguous()

    def generate_batch(self, traj_names: Sequence[str], idx: Optional[Sequence[int]] = None,
                       traj_scale: Optional[Sequence[float]] = None, centers=None, target_orientation=None):
        """One reference curve per entry (name, plane, scale, centre) -> (xyz [n,P,3], q [n,P,num_joints], status [n,P])
        as device tensors; way-points are built on the device and all tracks are solved in one launch."""
        xyz = self.curves_on_device(traj_names, idx, traj_scale, centers)
        q, status = self.solve_tracks(xyz, target_orientation)
        return xyz, q[:, :, :self.num_joints], status

    # ---- reference API ------------------------------------------------------------------------------------------
    def _solve_ik(self, target_pos: np.ndarray, target_quat: Optional[np.ndarray]) -> Optional[np.ndarray]:
        """[REF control/TrajectoryGenerator.py:81-116]: one way-point from the generator's current joint vector; the
        joint vector is updated in place (inplace=True) even when the solve fails."""
        q, status, = self.solve_tracks(np.asarray(target_pos, dtype=np.float64).reshape(1, 1, 3), target_quat,
                                       q0=self.qpos)[:2]
        # the kernel restores the start vector on failure (what `generate` does next anyway)
        ok = bool(int(status[0, 0].item()) & IK_SUCCESS)
        if ok:
            self.qpos = q[0, 0].cpu().numpy().copy()
            return self.qpos[:self.num_joints].copy()
        return None

    def generate(self, traj_name: str = "Fig8", target_orientation=np.array([1.0, 0.0, 0.0, 0.0])):
        """[REF control/TrajectoryGenerator.py:118-213]."""
        xyz_coords, _ = reference_curve(traj_name, self.idx, self.time_horizon, self.time_steps_per_sec,
                                        self.traj_scale)
        q, status = self.solve_tracks(xyz_coords[None], target_orientation, q0=np.zeros(T.NV))
        status = status[0].cpu().numpy()
        if status[0] & IK_ABORTED:
            raise RuntimeError("轨迹的第一个点IK求解失败,请检查目标位置是否在机器人工作空间内。")   # reference text
        for i in np.nonzero((status & IK_SUCCESS) == 0)[0]:
            print(f"警告: 逆运动学在时间步 {i} (目标位置: {np.round(xyz_coords[i], 3)}) 求解失败。")
        joint_angles = q[0, :, :self.num_joints].cpu().numpy().copy()
        self.qpos = q[0, -1].cpu().numpy().copy()
        return xyz_coords, joint_angles, self.time_vector
