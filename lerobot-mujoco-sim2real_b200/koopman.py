"""Koopman model of the reference on the device side of the batched simulator.

The reference trains `Koopmanlinear(x_dim=8, u_dim=5, encode_layers=[8,64,64,64,64,24])` on data from the stepping path
[REF models/KoopmanBase.py:12-60, args.py:103] and drives the arm with an MPC over that model
[REF control/MPC_Controler.py, Koopman_MPC.py:197-222].  This module provides what that loop needs next to the
batched stepper:

  * `KoopmanModel`      weights (from a state_dict / .npz), the lift  z = [x, encoder(x)]  on the current torch device;
  * `score`             n control sequences rolled through  z+ = A z + B u  and costed as the reference's MPC does, in
                        one CUDA launch through the C ABI (`so101_koopman_score`) - the model-side twin of `shoot`;
  * `mpc_gains/mpc_control`  the minimiser of the reference's MPC problem in closed form, for both of its formulations:
                        'mpc' (decision variable u) and 'delta_mpc' (decision variable delta_u, u_t = u_prev + sum delta_u,
                        cost on delta_u; the reference's default [REF args.py:75]).  The NLP handed to IPOPT
                        [REF MPC_Controler.py:65-141] has no constraints and a cost that is quadratic in the decision
                        variable (linear model, `linearize_B` is the identity for it), so its solution is one linear solve;
                        the gain matrices depend on the model only.
  * `lift_device/feedforward/mpc_step`  the loop body on the device: fused encoder MLP + gain product + clip kernels
                        behind the C ABI (`so101_koopman_lift`, `_feedforward`, `_mpc_step`); `lift` (torch matmuls) stays
                        as the plain reference the kernels are tested against.
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
        self._host_layers = [(np.ascontiguousarray(w[f"x_encode_net.linear_{j}.weight"]),
                              np.ascontiguousarray(w[f"x_encode_net.linear_{j}.bias"])) for j in range(len(self.layers))]
        self._h = None            # So101Koopman* (created on first use of a device kernel)
        self._dev_gains = None    # key of the gains currently uploaded to it

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and _lib is not None and _lib._LIB is not None:
            _lib._LIB.so101_koopman_destroy(h)
            self._h = None

    # ---- the device-side encoder / MPC kernels -----------------------------------------------------------------------
    def _handle(self):
        if self._h is None:
            L = _lib.lib()
            n = len(self._host_layers)
            dims = (C.c_int32 * (n + 1))(self.x_dim, *[Wb[0].shape[0] for Wb in self._host_layers])
            Wp = (C.c_void_p * n)(*[Wb[0].ctypes.data for Wb in self._host_layers])
            bp = (C.c_void_p * n)(*[Wb[1].ctypes.data for Wb in self._host_layers])
            h = C.c_void_p()
            _lib.check(L.so101_koopman_create(n, dims, Wp, bp, self.device.index or 0, C.byref(h)))
            self._h = h
        return self._h

    @staticmethod
    def _x_layout(x: torch.Tensor, x_dim: int) -> int:
        """dtype code of an observation tensor (float64 or float32)"""
        return {torch.float64: T.F64, torch.float32: T.F32}[x.dtype]

    def lift_device(self, x: torch.Tensor, soa: bool = False) -> torch.Tensor:
        """z = [x, encoder(x)] by the fused MLP kernel: x [n, >= x_dim] rows, or (soa) [x_dim, n]; float64 or float32."""
        assert x.is_cuda and x.is_contiguous() and x.dim() == 2
        n = x.shape[1] if soa else x.shape[0]
        Z = torch.empty((n, self.nz), dtype=torch.float64, device=x.device)
        _lib.check(_lib.lib().so101_koopman_lift(self._handle(), x.data_ptr(), self._x_layout(x, self.x_dim), 1 if soa else 0,
                                                 0 if soa else x.shape[1], n, Z.data_ptr(),
                                                 torch.cuda.current_stream(x.device).cuda_stream))
        return Z

    def _upload_gains(self, H: int, mpc_type: str, q_weight: float = 50.0, r_weight: float = 0.5):
        key = (H, q_weight, r_weight, mpc_type)
        if self._dev_gains != key:
            Kz, Kr, Ku = self.mpc_gains3(H, q_weight, r_weight, mpc_type)
            nu = self.nu
            kz = np.ascontiguousarray(Kz[:nu].cpu().numpy()); kr = np.ascontiguousarray(Kr[:nu].cpu().numpy())
            ku = np.ascontiguousarray(Ku[:nu].cpu().numpy())
            _lib.check(_lib.lib().so101_koopman_set_gains(self._handle(), H, nu, kz.ctypes.data, kr.ctypes.data, ku.ctypes.data))
            self._dev_gains = key

    def feedforward(self, xref: torch.Tensor, H: int = 10, mpc_type: str = "delta_mpc") -> torch.Tensor:
        """Reference part of the MPC law for whole trajectories: xref [n, P, x_dim] float64 ->
        uff [n, P, nu], uff[e, k] = sum_t Kr_t Psi(xref[e, k+1+t]) over the window rows that exist (rows past the end of
        the trajectory are zero in lifted space, as the reference leaves them [REF Koopman_MPC.py:199-203])."""
        assert xref.is_cuda and xref.is_contiguous() and xref.dtype == torch.float64 and xref.shape[2] == self.x_dim
        self._upload_gains(H, mpc_type)
        n, P = xref.shape[:2]
        uff = torch.empty((n, P, self.nu), dtype=torch.float64, device=xref.device)
        _lib.check(_lib.lib().so101_koopman_feedforward(self._handle(), xref.data_ptr(), n, P, uff.data_ptr(),
                                                        torch.cuda.current_stream(xref.device).cuda_stream))
        return uff

    def mpc_step(self, obs: torch.Tensor, soa: bool, uff: Optional[torch.Tensor], frame: int, u_prev: torch.Tensor,
                 ctrl: torch.Tensor, a_out: Optional[torch.Tensor] = None, H: int = 10, mpc_type: str = "delta_mpc",
                 clip: float = 0.5) -> None:
        """One frame of the reference's loop body for all envs, one launch: lift(obs), u0 = Kz z0 + uff[:, frame] +
        Ku u_prev + u_prev, a = clip(u0), u_prev <- u0 [REF MPC_Controler.py:143-152, Koopman_MPC.py:212-217].
        obs: rows [n, >= x_dim] or (soa) [x_dim, n]; u_prev [nu, n] float64 (in/out); ctrl [nu, n] (out, the stepper's
        control rows, its dtype); a_out [n, nu] float64 or None."""
        self._upload_gains(H, mpc_type)
        n = obs.shape[1] if soa else obs.shape[0]
        assert u_prev.shape == (self.nu, n) and u_prev.dtype == torch.float64 and u_prev.is_contiguous()
        assert ctrl.shape[0] >= self.nu and ctrl.shape[1] == n and ctrl.is_contiguous()
        up, us = (None, 0)
        if uff is not None:
            assert uff.is_contiguous() and uff.shape[0] == n and uff.shape[2] == self.nu
            up, us = uff.data_ptr() + frame * self.nu * 8, uff.shape[1] * self.nu
        _lib.check(_lib.lib().so101_koopman_mpc_step(
            self._handle(), obs.data_ptr(), self._x_layout(obs, self.x_dim), 1 if soa else 0, 0 if soa else obs.shape[1],
            up, us, u_prev.data_ptr(), ctrl.data_ptr(), {torch.float64: T.F64, torch.float32: T.F32}[ctrl.dtype],
            None if a_out is None else a_out.data_ptr(), float(clip), n, torch.cuda.current_stream(obs.device).cuda_stream))

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
    def mpc_gains(self, H: int = 10, q_weight: float = 50.0, r_weight: float = 0.5, mpc_type: str = "mpc"):
        """'mpc'  -> (Kz [H*nu, nz], Kr [H*nu, H*nz]) with u* = Kz z0 + Kr vec(zref): the minimiser of
                     sum_t q |z_{t+1} - zref_t|^2 + r |u_t|^2, z_{t+1} = A z_t + B u_t [REF MPC_Controler.py:65-98];
        'delta_mpc' -> (Kz, Kr, Ku [H*nu, nu]) with delta_u* = Kz z0 + Kr vec(zref) + Ku u_prev: the minimiser of
                     sum_t q |z_{t+1} - zref_t|^2 + r |delta_u_t|^2 with u_t = u_prev + sum_{s<=t} delta_u_s
                     [REF MPC_Controler.py:100-141].  `mpc_gains3` returns three matrices for either formulation."""
        if mpc_type == "mpc":
            return self.mpc_gains3(H, q_weight, r_weight, "mpc")[:2]
        return self.mpc_gains3(H, q_weight, r_weight, mpc_type)

    def mpc_gains3(self, H: int = 10, q_weight: float = 50.0, r_weight: float = 0.5, mpc_type: str = "delta_mpc"):
        """-> (Kz, Kr, Ku) for either formulation (Ku = 0 for 'mpc')."""
        if mpc_type not in ("mpc", "delta_mpc"):
            raise ValueError(f"MPC_type must be 'mpc' or 'delta_mpc', got {mpc_type!r}")
        key = (H, q_weight, r_weight, mpc_type)
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
            if mpc_type == "mpc":
                Hs = q_weight * G.T @ G + r_weight * np.eye(H * nu)
                Kr = np.linalg.solve(Hs, q_weight * G.T)
                Kz = -Kr @ F
                Ku = np.zeros((H * nu, nu))
            else:
                S = np.kron(np.tril(np.ones((H, H))), np.eye(nu))             # u = S delta_u + E u_prev
                E = np.kron(np.ones((H, 1)), np.eye(nu))
                Gd = G @ S
                Hs = q_weight * Gd.T @ Gd + r_weight * np.eye(H * nu)
                Kr = np.linalg.solve(Hs, q_weight * Gd.T)
                Kz = -Kr @ F
                Ku = -Kr @ (G @ E)
            self._gains[key] = tuple(torch.as_tensor(K, device=self.device) for K in (Kz, Kr, Ku))
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

    # ---- bilinear model variants (DBKN / IBKN) --------------------------------------------------------------------------
    def set_bilinear(self, H_weight: np.ndarray, u_z: bool = False) -> None:
        """Attach the bilinear term of `KoopmanBlinear` [REF models/KoopmanBase.py:62-110]: z+ = A z + B u + H (z (x) u),
        `H_weight` = net.H.weight [nz, nz*nu] (u_z: the Kronecker product is taken as u (x) z and permuted, as
        `get_Hi_numpy` does).  The reference's MPC then freezes B over the horizon at B_total(z0) = B + sum_j z0_j H_hat_j
        (`linearize_B`, [REF control/MPC_Controler.py:46-63]), so the problem stays an unconstrained quadratic - with gains
        that depend on z0: one 50 x 50 solve per env (`mpc_control_bilinear`, batched tensor algebra: this variant is not
        on the measured hot path; the reference ships no bilinear checkpoint)."""
        Hd = np.asarray(H_weight, dtype=np.float64)
        nz, nu = self.nz, self.nu
        assert Hd.shape == (nz, nz * nu)
        if u_z:   # columns ordered (u_i, z_j) -> (z_j, u_i)
            Hd = Hd.reshape(nz, nu, nz).transpose(0, 2, 1).reshape(nz, nz * nu)
        self.H_hat = torch.as_tensor(Hd.reshape(nz, nz, nu).transpose(1, 0, 2).copy(), device=self.device)   # [j][nz][nu]

    def mpc_control_bilinear(self, x: torch.Tensor, xref: torch.Tensor, u_prev: Optional[torch.Tensor] = None, H: int = 10,
                             mpc_type: str = "delta_mpc", q_weight: float = 50.0, r_weight: float = 0.5,
                             clip: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
        """One controller step of the bilinear variants for a batch: x [N, x_dim], xref [N, H, x_dim] (rows past the end of a
        trajectory: pass zeros and `valid` rows only are lifted - here all rows are lifted), u_prev [N, nu] (zeros if None)
        -> (u0 [N, nu] unclipped = u_opt[0] + u_prev, a = clip(u0)) [REF control/MPC_Controler.py:143-152]."""
        N, nz, nu = x.shape[0], self.nz, self.nu
        dev = self.device
        u_prev = torch.zeros((N, nu), dtype=torch.float64, device=dev) if u_prev is None else u_prev.to(dev, torch.float64)
        z0 = self.lift(x)                                                        # [N, nz]
        zref = self.lift(xref.reshape(N * H, -1)).reshape(N, H * nz)
        A = torch.as_tensor(self.A, device=dev)
        Bt = torch.as_tensor(self.B, device=dev)[None] + torch.einsum("nj,jab->nab", z0, self.H_hat)   # linearize_B
        pows = [torch.eye(nz, dtype=torch.float64, device=dev)]
        for _ in range(H):
            pows.append(A @ pows[-1])
        P = torch.stack(pows)                                                    # [H+1, nz, nz]
        AB = torch.einsum("kab,nbc->nkac", P[:H], Bt)                            # A^k B_total, k = 0..H-1
        G = torch.zeros((N, H, nz, H, nu), dtype=torch.float64, device=dev)
        for t in range(H):
            for s_ in range(t + 1):
                G[:, t, :, s_, :] = AB[:, t - s_]
        G = G.reshape(N, H * nz, H * nu)
        Fz = torch.einsum("kab,nb->nka", P[1:], z0).reshape(N, H * nz)           # F z0
        if mpc_type == "delta_mpc":
            S = torch.kron(torch.tril(torch.ones((H, H), dtype=torch.float64, device=dev)), torch.eye(nu, dtype=torch.float64, device=dev))
            Gd = G @ S
            off = Fz + (G.reshape(N, H * nz, H, nu).sum(dim=2) @ u_prev[:, :, None])[:, :, 0] - zref
        elif mpc_type == "mpc":
            Gd, off = G, Fz - zref
        else:
            raise ValueError(f"MPC_type must be 'mpc' or 'delta_mpc', got {mpc_type!r}")
        Hs = q_weight * Gd.transpose(1, 2) @ Gd + r_weight * torch.eye(H * nu, dtype=torch.float64, device=dev)
        rhs = -q_weight * (Gd.transpose(1, 2) @ off[:, :, None])
        sol = torch.linalg.solve(Hs, rhs)[:, :nu, 0]
        u0 = sol + u_prev
        return u0, torch.clamp(u0, -clip, clip)

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
