"""The control loop of the reference's `Koopman_MPC.py`, for n reference curves at once, on the device
(rows N3 + N4 joined to the stepper).

Reference loop [REF Koopman_MPC.py:83-90 (runBefore), 109-126 and 197-222 (runFunc / runMPC)]:
  * the reference trajectory is `state_all_ref = hstack([cartesian_points, joint_angle_traj])` — the 8-dim observation
    `[ee(3) | q(5)]` the Koopman model was trained on — produced by `CartesianTrajectoryGenerator.generate`;
  * runBefore: `qpos[:5] = joint_angle_traj[0]`, `mj_forward`, and the controller's state starts as `state_all_ref[0]`;
  * every frame: `qfrc_applied = qfrc_bias` (gravity compensation), the window `state_all_ref[k+1 : k+H+1]` is lifted
    row by row into a zero-initialised `[H, nz]` array (so the rows past the end of the trajectory stay ZERO in lifted
    space), `z0 = Psi(state)`, the MPC returns `u`, `a = clip(u, -0.5, 0.5)`, `s_next = env.step(a)`, the state becomes
    `s_next` and is appended to `actual_traj`; one way-point per env step, `total_frames` frames.

Here: the curves come from `TrajectoryGenerator.solve_tracks` (one IK launch), the MPC is the closed-form minimiser of
the same problem (`KoopmanModel.mpc_gains`), the environment is `SOARM101VecEnv(gravity_compensation=True)`; the loop
body is a handful of device launches per frame for all n curves.  Viewer, ZMQ bridge, the return-to-home phase and the
50 Hz sleep of the reference loop are its control plane and are not mirrored.
"""
from __future__ import annotations

from typing import Optional

import torch

from .koopman import KoopmanModel
from .vec_env import SOARM101VecEnv


class BatchedKoopmanMPC:
    def __init__(self, env: SOARM101VecEnv, model: KoopmanModel, cartesian_points: torch.Tensor,
                 joint_angle_traj: torch.Tensor, H: int = 10, clip: float = 0.5):
        """env: n environments (construct it with gravity_compensation=True to mirror [REF Koopman_MPC.py:119]);
        cartesian_points [n, P, 3], joint_angle_traj [n, P, 5]: what `generate` / `generate_batch` return."""
        n = env.num_envs
        cp = torch.as_tensor(cartesian_points, dtype=torch.float64, device=env.device)
        ja = torch.as_tensor(joint_angle_traj, dtype=torch.float64, device=env.device)
        if cp.dim() == 2:
            cp, ja = cp.unsqueeze(0), ja.unsqueeze(0)
        if cp.shape[0] != n or ja.shape[:2] != cp.shape[:2] or cp.shape[2] != 3 or ja.shape[2] != 5:
            raise ValueError(f"need cartesian_points [{n}, P, 3] and joint_angle_traj [{n}, P, 5]")
        self.env, self.model, self.H, self.clip = env, model, int(H), float(clip)
        self.state_all_ref = torch.cat([cp, ja], dim=2)                      # [n, P, 8]   (:50)
        self.total_frames = int(cp.shape[1])
        self.traj_index = 0
        self.actual_traj = []
        self.state_tensor: Optional[torch.Tensor] = None
        self.Kz, self.Kr = model.mpc_gains(self.H)
        self._zwin: Optional[torch.Tensor] = None

    def runBefore(self) -> None:
        """[REF Koopman_MPC.py:83-90]"""
        n = self.env.num_envs
        init = torch.cat([self.state_all_ref[:, 0, 3:8], torch.zeros((n, 5), dtype=torch.float64,
                                                                    device=self.env.device)], dim=1)
        obs0 = self.env.reset(options={"initial_state": init})[0]            # qpos[:5] = joint_angle_traj[0]; mj_forward
        self.obs0 = obs0.to(torch.float64).clone()
        self.state_tensor = self.state_all_ref[:, 0].clone()
        self.traj_index = 0
        self.actual_traj = []
        self.applied = []
        self._zwin = None

    def _lift_ref_row(self, j: int) -> torch.Tensor:
        """Lifted reference row j of every curve, ZERO past the end of the trajectory (the reference fills a
        zero-initialised array, [REF Koopman_MPC.py:199-203])."""
        if j < self.total_frames:
            return self.model.lift(self.state_all_ref[:, j])
        return torch.zeros((self.env.num_envs, self.model.nz), dtype=torch.float64, device=self.env.device)

    def runMPC(self) -> torch.Tensor:
        """One frame [REF Koopman_MPC.py:197-222] -> the applied control a [n, 5].

        Consecutive windows [k+1, k+H] overlap in H-1 rows, so the lifted window lives in a circular buffer: each frame
        lifts ONE new reference row per curve (the reference re-lifts all H) and the matching column blocks of the
        reference gain are rotated instead of the data."""
        n, H, nz, nu = self.env.num_envs, self.H, self.model.nz, self.model.nu
        k = self.traj_index
        if self._zwin is None:
            self._zwin = torch.empty((n, H, nz), dtype=torch.float64, device=self.env.device)
            for j in range(H):
                self._zwin[:, (k + j) % H] = self._lift_ref_row(k + 1 + j)
        else:
            self._zwin[:, (k - 1) % H] = self._lift_ref_row(k + H)     # the slot of row k (just consumed) <- row k+H
        # slot s holds row k+1+((s - k) mod H): window position j = (s - k) mod H  <=>  s = (k + j) mod H
        order = [(k + j) % H for j in range(H)]
        Kr_rot = torch.empty((nu, H, nz), dtype=torch.float64, device=self.env.device)
        Kr_rot[:, order] = self.Kr[:nu].reshape(nu, H, nz)
        z0 = self.model.lift(self.state_tensor)
        u = z0 @ self.Kz[:nu].t() + self._zwin.reshape(n, H * nz) @ Kr_rot.reshape(nu, H * nz).t()
        a = torch.clamp(u, -self.clip, self.clip)                            # get_control [REF MPC_Controler.py:149]
        s_next = self.env.step(a)[0]
        self.state_tensor = s_next.to(torch.float64).clone()
        self.actual_traj.append(self.state_tensor)
        self.applied.append(a)
        self.traj_index += 1
        return a

    def run(self, frames: Optional[int] = None) -> torch.Tensor:
        """runBefore + `frames` (default: all) frames -> actual_traj [n, frames, 8]."""
        self.runBefore()
        for _ in range(self.total_frames if frames is None else int(frames)):
            self.runMPC()
        return torch.stack(self.actual_traj, dim=1)

    def dataset_rows(self) -> torch.Tensor:
        """The closed-loop run as Koopman training data in the layout of `generate_physics_based_data`
        [REF SOARM101/SOARM101_DataCollection.py:108-134]: rows [n, frames, 13] float64 = [u_i (5) | ee_i (3) | q_i (5)]
        with u_i the control applied FROM observation i (s_0 = the observation after runBefore's reset, s_{i+1} =
        step(u_i)) - on-policy data along the Cartesian curves, consumable by the reference's `Collater`."""
        if not self.applied:
            raise RuntimeError("run() first")
        states = torch.stack([self.obs0] + self.actual_traj[:-1], dim=1)     # s_0 .. s_{frames-1}
        return torch.cat([torch.stack(self.applied, dim=1), states], dim=2).contiguous()
