"""TEST INFRASTRUCTURE ONLY - numpy restatement of the reference's Koopman model and MPC cost.

Follows [REF models/KoopmanBase.py:44-57] (lift z = [x, MLP(x)], z+ = lA z + lB u, x = lC z = first 8 coordinates)
and [REF control/MPC_Controler.py:65-98] (horizon roll-out and cost with Q = 50 I over the lifted state, R = 0.5 I;
`state_full` is always true there).  Pinned by the reference's own checkpoint (tests/golden/koopman_dkuc.npz, exported
from results/SOARM101/11_27/DKUC/best_model.pt) through the fingerprints in tests/test_oracle.py.
"""
import numpy as np


def lift(W, x):
    h = x
    n = sum(1 for k in W if k.startswith("x_encode_net.linear_") and k.endswith(".weight"))
    for i in range(n):
        h = h @ W[f"x_encode_net.linear_{i}.weight"].T + W[f"x_encode_net.linear_{i}.bias"]
        if i != n - 1:
            h = np.maximum(h, 0)
    return np.concatenate([x, h], -1)


def score(W, z0, U, zref=None, q=50.0, r=0.5, nobs=8):
    """z0 [nz], U [H, nu, n], zref [H, nz] -> (Xhat [n, H+1, nobs], cost [n]) by plain loops over time."""
    A, B = W["lA.weight"], W["lB.weight"]
    H, nu, n = U.shape
    z = np.tile(z0, (n, 1))
    X = [z[:, :nobs].copy()]
    cost = np.zeros(n)
    for t in range(H):
        u = U[t].T                                  # [n, nu]
        z = z @ A.T + u @ B.T
        if zref is not None:
            cost += q * ((z - zref[t]) ** 2).sum(-1)
        cost += r * (u ** 2).sum(-1)
        X.append(z[:, :nobs].copy())
    return np.stack(X, 1), cost


def linearize_B(W, z0, u_z=False):
    """B_total = Bd + sum_j z0[j] H_hat_j with H_hat_j as `KoopmanBlinear.get_Hi_numpy` builds them
    [REF control/MPC_Controler.py:46-63, models/KoopmanBase.py:85-110]; the model's own B when it has no H layer."""
    B = W["lB.weight"]
    if "H.weight" not in W:
        return B
    nz, nu = B.shape
    Hd = W["H.weight"]
    if u_z:       # P of build_permutation_matrix: vec(u (x) z) -> vec(z (x) u)
        P = np.zeros((nu * nz, nu * nz))
        for i in range(nu):
            for j in range(nz):
                P[j * nu + i, i * nz + j] = 1
        Hd = Hd @ P.T
    Bt = B.copy()
    for j in range(nz):
        Bt = Bt + z0[j] * Hd[:, j * nu:(j + 1) * nu]
    return Bt


def mpc_solve(W, z0, zref, H, q=50.0, r=0.5, mpc_type="mpc", u_prev=None, u_z=False):
    """The minimiser IPOPT converges to for the reference's unconstrained, quadratic problem.

    'mpc'       [REF control/MPC_Controler.py:65-98]:  decision u,  cost sum_t q |z_{t+1} - zref_t|^2 + r |u_t|^2
    'delta_mpc' [REF control/MPC_Controler.py:100-141]: decision delta_u, u_t = u_{t-1} + delta_u_t (u_{-1} = u_prev),
                cost sum_t q |z_{t+1} - zref_t|^2 + r |delta_u_t|^2
    Restated as the reference writes it - the model is rolled forward step by step for a given decision vector - and
    solved as a linear least-squares problem: the residual vector is affine in the decision, its Jacobian is formed by
    rolling unit perturbations.  Shares no structure with the closed-form gains of the product (koopman.py)."""
    A, B = W["lA.weight"], linearize_B(W, z0, u_z)      # the bilinear variants freeze B at B_total(z0) over the horizon
    nz, nu = B.shape
    u_prev = np.zeros(nu) if u_prev is None else np.asarray(u_prev, dtype=np.float64)

    def residual(x):
        x = x.reshape(H, nu)
        z, u, res = z0.copy(), u_prev.copy(), []
        for t in range(H):
            u = u + x[t] if mpc_type == "delta_mpc" else x[t]
            z = A @ z + B @ u
            res.append(np.sqrt(q) * (z - zref[t]))
            res.append(np.sqrt(r) * x[t])
        return np.concatenate(res)

    r0 = residual(np.zeros(H * nu))
    J = np.stack([residual(e) - r0 for e in np.eye(H * nu)], axis=1)
    return np.linalg.lstsq(J, -r0, rcond=None)[0].reshape(H, nu)


class Controller:
    """`MPCController.get_control` + the `u_prev` hand-off of `Koopman_MPC.runMPC`
    [REF control/MPC_Controler.py:143-152, Koopman_MPC.py:206-217]: u0 = u_opt[0] + u_eso (= 0) + u_prev,
    a = clip(u0, -0.5, 0.5); get_control stores a as u_prev in 'delta_mpc' mode, and runMPC then overwrites it with the
    unclipped u0 in BOTH modes; in 'delta_mpc' mode u_prev is also a parameter of the problem."""

    def __init__(self, W, H=10, mpc_type="delta_mpc", clip=0.5):
        self.W, self.H, self.mpc_type, self.clip = W, H, mpc_type, clip
        self.u_prev = np.zeros(W["lB.weight"].shape[1])

    def step(self, state, zref):
        u_opt = mpc_solve(self.W, lift(self.W, state), zref, self.H, mpc_type=self.mpc_type,
                          u_prev=self.u_prev if self.mpc_type == "delta_mpc" else None)
        u0 = u_opt[0] + self.u_prev
        a = np.clip(u0, -self.clip, self.clip)
        if self.mpc_type == "delta_mpc":
            self.u_prev = a.copy()
        self.u_prev = u0          # runMPC: `self.mpc_controller.u_prev = u`
        return u0, a


def k_linear_loss(W, x, u, start_idx, gamma=0.99, pre_length=5):
    """The reference's training loss [REF models/losses.py:54-129] for the linear model (loss_name "mse"): from
    x0 = x[:, start_idx] the lifted state is rolled `pre_length` steps under z+ = lA z + lB u (u_encoder is the
    identity) and compared, gamma-weighted, (i) in lifted space with the lift of the true next state (koopman_loss),
    (ii) after decoding with lC (pred_loss), (iii) recon_loss = mse(lC lift(x1), x1); total = their sum; all divided
    by the sum of the weights.  stable_Loss = 1e-3 * sum(max(|eig(lA)| - 1, 0)) [REF :18-21, :117].
    x [n, steps, 8], u [n, steps, 5] -> dict of floats."""
    A, B, Cm = W["lA.weight"], W["lB.weight"], W["lC.weight"]
    mse = lambda a, b: float(((a - b) ** 2).mean())
    z = lift(W, x[:, start_idx])
    out = dict(koopman_loss=0.0, pred_loss=0.0, recon_loss=0.0, total_loss=0.0)
    beta, cont = 1.0, 0.0
    for i in range(start_idx, start_idx + pre_length):
        z = z @ A.T + u[:, i] @ B.T
        x1 = x[:, i + 1]
        z1 = lift(W, x1)
        k, p, r = mse(z, z1), mse(z @ Cm.T, x1), mse(z1 @ Cm.T, x1)
        out["koopman_loss"] += beta * k; out["pred_loss"] += beta * p; out["recon_loss"] += beta * r
        out["total_loss"] += beta * (k + p + r)
        cont += beta
        beta *= gamma
    out = {k: v / cont for k, v in out.items()}
    out["stable_Loss"] = 1e-3 * float(np.clip(np.abs(np.linalg.eigvals(A)) - 1.0, 0.0, None).sum())
    return out


PUBLISHED_TRAIN = dict(koopman_loss=4.1822929293272153e-07, pred_loss=9.377657459082502e-07,
                       recon_loss=4.393824788019431e-09, total_loss=1.360388863313245e-06,
                       stable_Loss=1.9319259122283284e-06)   # [REF results/SOARM101/11_27/DKUC/best_scores.json:3-8]


def check_training_loss_fingerprints(W, x, u, label=""):
    """Four more statistics the reference holds from REAL MuJoCo data of this path: the epoch-429 means of its training
    loss terms [REF results/.../best_scores.json:3-8], recomputed with `k_linear_loss` [REF models/losses.py:54-129]
    restated in oracle/koopman_oracle.py on train-shaped data (21-row 'random' trajectories, float32 like the
    reference's TensorDataset) averaged over all 15 window starts (the reference draws one per batch).  A model that
    fitted its own training set scores a little better there than on fresh data, hence the one-sided bands."""
    x = x.astype(np.float32).astype(np.float64); u = u.astype(np.float32).astype(np.float64)
    acc = {}
    for s in range(x.shape[1] - 5 - 1 + 1):                   # random.randint(0, steps - pre_length - 1) is inclusive
        for k, v in k_linear_loss(W, x, u, s).items():
            acc.setdefault(k, []).append(v)
    got = {k: float(np.mean(v)) for k, v in acc.items()}
    print(label, {k: f"{got[k]:.3e} (published {PUBLISHED_TRAIN[k]:.3e})" for k in got})
    for k in ("koopman_loss", "pred_loss", "total_loss"):
        assert 0.9 * PUBLISHED_TRAIN[k] < got[k] < 1.3 * PUBLISHED_TRAIN[k], (k, got[k])
    assert 0.8 * PUBLISHED_TRAIN["recon_loss"] < got["recon_loss"] < 1.2 * PUBLISHED_TRAIN["recon_loss"]
    assert abs(got["stable_Loss"] / PUBLISHED_TRAIN["stable_Loss"] - 1) < 0.02     # model-only statistic
    return got
