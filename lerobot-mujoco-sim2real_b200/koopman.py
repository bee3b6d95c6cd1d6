"""Koopman model of the reference on the device side of the batched simulator.

The reference trains `Koopmanlinear(x_dim=8, u_dim=5, encode_layers=[8,64,64,64,64,24])` on data from the stepping path
[REF models/KoopmanBase.py:12-60, args.py:103] and drives the arm with an MPC over that model
[REF control/MPC_Controler.py, Koopman_MPC.py:197-222].  This module provides what that loop needs next to the
batched stepper:

  * `KoopmanModel`      weights (from a state_dict / .npz), the lift  z = [x, encoder(x)]  on the current torch device;
  * `score`             n control sequences rolled through  z+ = A z + B u  and costed as the reference's MPC does, in
                        one CUDA launch through the C ABI (`so101_koopman_score`) - the model-side twin of `shoot`;
  * `mpc_gains/mpc_control`  the minimiser of the reference's MPC problem in closed form.  The NLP handed to IPOPT
                        [REF MPC_Controler.py:65-107] has no constraints and a cost that is quadratic in u (linear model,
                        `linearize_B` is the identity for the linear model), so its solution is one linear solve:
                        u* = -(G'QG + R)^-1 G'Q (F z0 - zref); the gain matrices depend on the model only.
Only torch plumbing and the C ABI are used; nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from . import tables as T


class KoopmanModel:
    def __init__(self, weights: Dict[str, np.ndarray], device: Optional[torch.device] = None):
        """weights: the reference checkpoint's state_dict as numpy arrays (keys `x_encode_net.linear_{i}.weight/bias`,
        `lA.weight`, `lB.weight`) - e.g. `torch.load(best_model.pt)` converted, or an .npz."""
        self.device = torch.device(device if device is not None else "cuda")
        w = {k: np.asarray(v, dtype=np.float64) for k, v in weights.items()}
        self.A = np.ascontiguousarray(w["lA.weight"])              # [nz, nz]
        self.B = np.ascontiguousarray(w["lB.weight"])              # [nz, nu]
        self.nz, self.nu = self.B.shape
        self.layers = []
        i = 0
        while f"x_encode_net.linear_{i}.weight" in w:
            self.layers.append((torch.as_tensor(w[f"x_encode_net.linear_{i}.weight"], device=self.device),
                                torch.as_tensor(w[f"x_encode_net.linear_{i}.bias"], device=self.device)))
            i += 1
        self.x_dim = self.layers[0][0].shape[1]
        assert self.nz == self.x_dim + self.layers[-1][0].shape[0] and self.A.shape == (self.nz, self.nz)
        self._gains = {}

    @classmethod
    def from_npz(cls, path: str, device: Optional[torch.device] = None) -> "KoopmanModel":
        return cls(dict(np.load(path)), device)

    # ---- lift [REF KoopmanBase.py:44-47]: z = concat(x, MLP(x)), ReLU between layers, none after the last ----
    def lift(self, x: torch.Tensor) -> torch.Tensor:
        x = x.to(self.device, torch.float64)
        h = x
        for k, (W, b) in enumerate(self.layers):
            h = h @ W.t() + b
            if k != len(self.layers) - 1:
                h = torch.relu(h)
        return torch.cat([x, h], dim=-1)

    # ---- n control sequences under the model [REF MPC_Controler.py:76-84] ------------------------------------------
    def score(self, z0, U: torch.Tensor, zref=None, q_weight: float = 50.0, r_weight: float = 0.5,
              nobs: Optional[int] = None, want_pred: bool = True) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        """z0 [nz] (host or device), U [H, nu, n] device tensor (float64/float32, as `SOARM101VecEnv.shoot` takes),
        zref [H, nz] or None -> (Xhat [n, H+1, nobs] float32 predicted observations, cost [n] float64)."""
        assert U.is_cuda and U.is_contiguous() and U.dim() == 3 and U.shape[1] == self.nu
        H, _, n = U.shape
        nobs = self.x_dim if nobs is None else nobs
        z0h = np.ascontiguousarray(torch.as_tensor(z0).detach().cpu().numpy().reshape(self.nz), dtype=np.float64)
        zr = None
        if zref is not None:
            zr = np.ascontiguousarray(torch.as_tensor(zref).detach().cpu().numpy().reshape(H, self.nz), dtype=np.float64)
        dt = {torch.float64: T.F64, torch.float32: T.F32}[U.dtype]
        Xhat = torch.empty((n, H + 1, nobs), dtype=torch.float32, device=U.device) if want_pred else None
        cost = torch.empty((n,), dtype=torch.float64, device=U.device)
        _lib.check(_lib.lib().so101_koopman_score(
            self.A.ctypes.data, self.B.ctypes.data, self.nz, self.nu, z0h.ctypes.data,
            None if zr is None else zr.ctypes.data, float(q_weight), float(r_weight), U.data_ptr(), H, n, dt,
            U.device.index or 0, nobs, None if Xhat is None else Xhat.data_ptr(), cost.data_ptr(),
            torch.cuda.current_stream(U.device).cuda_stream))
        return Xhat, cost

    # ---- the reference's MPC problem solved exactly ------------------------------------------------------------------
    def mpc_gains(self, H: int = 10, q_weight: float = 50.0, r_weight: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (Kz [H*nu, nz], Kr [H*nu, H*nz]) with u* = Kz z0 + Kr vec(zref): the minimiser of
        sum_t q |z_{t+1} - zref_t|^2 + r |u_t|^2, z_{t+1} = A z_t + B u_t [REF MPC_Controler.py:65-98]."""
        key = (H, q_weight, r_weight)
        if key not in self._gains:
            nz, nu = self.nz, self.nu
            F = np.zeros((H * nz, nz)); G = np.zeros((H * nz, H * nu))
            Ap = np.eye(nz)
            pows = [np.eye(nz)]
            for t in range(H):
                Ap = self.A @ Ap
                pows.append(Ap)
                F[t * nz:(t + 1) * nz] = Ap                                   # z_{t+1} = A^{t+1} z0 + sum_s A^{t-s} B u_s
                for s in range(t + 1):
                    G[t * nz:(t + 1) * nz, s * nu:(s + 1) * nu] = pows[t - s] @ self.B
            Hs = q_weight * G.T @ G + r_weight * np.eye(H * nu)
            Kr = np.linalg.solve(Hs, q_weight * G.T)
            Kz = -Kr @ F
            self._gains[key] = (torch.as_tensor(Kz, device=self.device), torch.as_tensor(Kr, device=self.device))
        return self._gains[key]

    def mpc_control(self, x: torch.Tensor, xref: torch.Tensor, H: int = 10, clip: float = 0.5) -> torch.Tensor:
        """One MPC step for a batch: x [N, 8] observations, xref [N, H, 8] reference observations (lifted here, as
        [REF Koopman_MPC.py:200-204] does) -> first optimal control [N, nu], clipped like `get_control`
        [REF MPC_Controler.py:160]."""
        Kz, Kr = self.mpc_gains(H)
        N = x.shape[0]
        z0 = self.lift(x)                                                       # [N, nz]
        zref = self.lift(xref.reshape(N * H, -1)).reshape(N, H * self.nz)
        u = z0 @ Kz[:self.nu].t() + zref @ Kr[:self.nu].t()
        return torch.clamp(u, -clip, clip)

    # ---- sampling MPC through the physics ("replaces IPOPT with sampling MPC") -----------------------
    def shooting_mpc(self, shooter, state0, x: torch.Tensor, xref: torch.Tensor, sigma: float = 0.05,
                     clip: float = 0.5, q_weight: float = 50.0, r_weight: float = 0.5, flags: int = 0,
                     seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """One sampling-MPC step: B = shooter.num_envs candidate control sequences [H, nu] are rolled through the
        PHYSICS from the shared physics state `state0` (18 doubles: qpos, qvel, qacc_warmstart) in one `k_shoot` launch
        (BASELINE config 5) and costed with the reference's MPC cost on the lifted outcome,
            sum_t q |psi(x_{t+1}) - psi(xref_t)|^2 + r |u_t|^2          [REF control/MPC_Controler.py:65-98],
        candidate 0 = the closed-form minimiser under the Koopman MODEL (what `mpc_control` applies), the others =
        that sequence plus N(0, sigma^2) noise, all clipped to +-clip.  x [8] = current observation, xref [H, 8].
        -> (U_best [H, nu], costs [B] (physics-evaluated), index of the best).  The best candidate can only improve on
        the model's answer as judged by the true dynamics."""
        H = int(xref.shape[0])
        B = shooter.num_envs
        dev, dt = shooter.device, shooter.torch_dtype
        Kz, Kr = self.mpc_gains(H, q_weight, r_weight)
        z0 = self.lift(x.reshape(1, -1))                                        # [1, nz]
        zref = self.lift(xref)                                                  # [H, nz]
        u_model = (z0 @ Kz.t() + zref.reshape(1, -1) @ Kr.t()).reshape(H, self.nu)
        g = torch.Generator(device=dev).manual_seed(int(seed))
        noise = torch.randn((H, self.nu, B), generator=g, dtype=torch.float64, device=dev) * sigma
        noise[:, :, 0] = 0.0
        U = torch.clamp(u_model[:, :, None] + noise, -clip, clip).to(dt).contiguous()          # [H, nu, B]
        X = shooter.shoot(state0, U, flags=flags)                                # [B, H+1, 8] float32, physics
        Z = self.lift(X[:, 1:].reshape(B * H, -1).double()).reshape(B, H, self.nz)
        costs = q_weight * ((Z - zref[None]) ** 2).sum(dim=(1, 2)) + r_weight * (U.double() ** 2).sum(dim=(0, 1))
        best = int(torch.argmin(costs).item())
        return U[:, :, best].double(), costs, best
