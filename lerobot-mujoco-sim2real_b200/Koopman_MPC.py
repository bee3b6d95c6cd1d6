"""The control loop of the reference's `Koopman_MPC.py`, for n reference curves at once, on the device
(rows N3 + N4 joined to the stepper).

Reference loop [REF Koopman_MPC.py:83-90 (runBefore), 109-126 and 197-222 (runFunc / runMPC)]:
  * the reference trajectory is `state_all_ref = hstack([cartesian_points, joint_angle_traj])` — the 8-dim observation
    `[ee(3) | q(5)]` the Koopman model was trained on — produced by `CartesianTrajectoryGenerator.generate`;
  * runBefore: `qpos[:5] = joint_angle_traj[0]`, `mj_forward`, and the controller's state starts as `state_all_ref[0]`;
  * every frame: `qfrc_applied = qfrc_bias` (gravity compensation), the window `state_all_ref[k+1 : k+H+1]` is lifted
    row by row into a zero-initialised `[H, nz]` array (so the rows past the end of the trajectory stay ZERO in lifted
    space), `z0 = Psi(state)`, the MPC (`args.MPC_type`: 'delta_mpc' by default [REF args.py:75], or 'mpc') returns
    `u_opt`, `u0 = u_opt[0] + u_prev`, `a = clip(u0, -0.5, 0.5)` [REF control/MPC_Controler.py:143-152], the loop then
    sets `u_prev = u0` (unclipped, in both modes [REF Koopman_MPC.py:217]), `s_next = env.step(a)`, the state becomes
    `s_next` and is appended to `actual_traj`; one way-point per env step, `total_frames` frames.
    In 'mpc' mode `u_prev` still enters `u0` (the controller adds it whatever the formulation), so that mode integrates
    its own output; the reference runs 'delta_mpc', where `u_prev` is also the problem's parameter.  Both are mirrored
    as written.

Here: the curves come from `TrajectoryGenerator.solve_tracks` (one IK launch); the MPC is the closed-form minimiser of
the same problem (`KoopmanModel.mpc_gains3`); the reference part of the control law is folded once for the whole
trajectory (`KoopmanModel.feedforward`: every reference row lifted once instead of H times); a frame is TWO launches
for all n curves: `so101_koopman_mpc_step` (encoder MLP + gain product + clip, fused) and the stepper's
`so101_batch_step_flags` with the gravity compensation as a launch flag.  No torch matmul on the path.
Viewer, ZMQ bridge, the return-to-home phase and the 50 Hz sleep of the reference loop are its control plane and are
not mirrored.
"""
from __future__ import annotations

from typing import Optional

import torch

from .koopman import KoopmanModel
from .vec_env import SOARM101VecEnv


class BatchedKoopmanMPC:
    def __init__(self, env: SOARM101VecEnv, model: KoopmanModel, cartesian_points: torch.Tensor,
                 joint_angle_traj: torch.Tensor, H: int = 10, clip: float = 0.5, MPC_type: str = "delta_mpc"):
        """env: n environments (construct it with gravity_compensation=True to mirror [REF Koopman_MPC.py:119]);
        cartesian_points [n, P, 3], joint_angle_traj [n, P, 5]: what `generate` / `generate_batch` return;
        MPC_type: 'delta_mpc' (the reference's default, [REF args.py:75]) or 'mpc'."""
        n = env.num_envs
        cp = torch.as_tensor(cartesian_points, dtype=torch.float64, device=env.device)
        ja = torch.as_tensor(joint_angle_traj, dtype=torch.float64, device=env.device)
        if cp.dim() == 2:
            cp, ja = cp.unsqueeze(0), ja.unsqueeze(0)
        if cp.shape[0] != n or ja.shape[:2] != cp.shape[:2] or cp.shape[2] != 3 or ja.shape[2] != 5:
            raise ValueError(f"need cartesian_points [{n}, P, 3] and joint_angle_traj [{n}, P, 5]")
        if MPC_type not in ("mpc", "delta_mpc"):
            raise ValueError(f"MPC_type must be 'mpc' or 'delta_mpc', got {MPC_type!r}")
        self.env, self.model, self.H, self.clip, self.MPC_type = env, model, int(H), float(clip), MPC_type
        self.state_all_ref = torch.cat([cp, ja], dim=2).contiguous()         # [n, P, 8]   (:50)
        self.total_frames = int(cp.shape[1])
        self.traj_index = 0
        self.actual_traj = []
        self.state_tensor: Optional[torch.Tensor] = None
        # reference part of the control law, uff [n, rows, nu]: folded in runBefore for the frames that will run (each
        # reference row is lifted once)
        self.uff: Optional[torch.Tensor] = None
        self._uff_frames = 0
        self.u_prev = torch.zeros((model.nu, n), dtype=torch.float64, device=env.device)
        self._ctrl = torch.zeros((model.nu, n), dtype=env.torch_dtype, device=env.device)

    def _fold_reference(self, frames: int) -> None:
        """uff for frames [0, frames): needs reference rows [1, frames + H]."""
        frames = min(int(frames), self.total_frames)
        if self.uff is not None and self._uff_frames >= frames:
            return
        rows = min(self.total_frames, frames + self.H)
        ref = self.state_all_ref if rows == self.total_frames else self.state_all_ref[:, :rows].contiguous()
        self.uff = self.model.feedforward(ref, self.H, self.MPC_type)        # [n, rows, nu]
        # a window that reaches past `rows` but not past the trajectory would miss rows: only frames < rows - H are
        # complete unless the fold went to the end of the trajectory
        self._uff_frames = self.total_frames if rows == self.total_frames else rows - self.H

    def runBefore(self, frames: Optional[int] = None) -> None:
        """[REF Koopman_MPC.py:83-90]; frames: how many frames will run (default: the whole trajectory)"""
        self._fold_reference(self.total_frames if frames is None else frames)
        n = self.env.num_envs
        init = torch.cat([self.state_all_ref[:, 0, 3:8], torch.zeros((n, 5), dtype=torch.float64,
                                                                    device=self.env.device)], dim=1)
        obs0 = self.env.reset(options={"initial_state": init})[0]            # qpos[:5] = joint_angle_traj[0]; mj_forward
        self.obs0 = obs0.to(torch.float64).clone()
        self.state_tensor = self.state_all_ref[:, 0].clone()
        self.traj_index = 0
        self.actual_traj = []
        self.applied = []
        self.u_prev.zero_()                                                  # MPCController.__init__: u_prev = zeros

    def runMPC(self) -> torch.Tensor:
        """One frame [REF Koopman_MPC.py:197-222] -> the applied control a [n, 5]."""
        n = self.env.num_envs
        k = self.traj_index
        if k >= self._uff_frames:
            self._fold_reference(self.total_frames)
        a = torch.empty((n, self.model.nu), dtype=torch.float64, device=self.env.device)
        if k == 0:      # the controller's first state is the reference's first row (float64), then the env's observations
            self.model.mpc_step(self.state_tensor.contiguous(), False, self.uff, k, self.u_prev, self._ctrl, a, self.H,
                                self.MPC_type, self.clip)
        else:           # the stepper's own observation buffer (float32, structure of arrays): no copy, no transpose
            self.model.mpc_step(self.env._obs, True, self.uff, k, self.u_prev, self._ctrl, a, self.H, self.MPC_type,
                                self.clip)
        self.env.step_soa(self._ctrl)                                        # gravity compensation: a flag of the launch
        self.state_tensor = self.env._obs.t().to(torch.float64)
        self.actual_traj.append(self.state_tensor)
        self.applied.append(a)
        self.traj_index += 1
        return a

    def run(self, frames: Optional[int] = None) -> torch.Tensor:
        """runBefore + `frames` (default: all) frames -> actual_traj [n, frames, 8]."""
        self.runBefore(frames)
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
