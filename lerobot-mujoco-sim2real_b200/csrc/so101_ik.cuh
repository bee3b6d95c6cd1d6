// so101_ik.cuh - batched site-pose inverse kinematics along Cartesian way-point tracks.
//
// The reference turns a Cartesian reference curve (Fig8 / Circle) into joint angles one way-point at a time with
// dm_control.utils.inverse_kinematics.qpos_from_site_pose, warm-started from the previous way-point
// [REF control/TrajectoryGenerator.py:81-116 (_solve_ik), :180-210 (the way-point loop of generate)].  dm_control is a
// third-party dependency that is not vendored; its published algorithm is restated in the test oracle (ik_oracle).  Here the
// same iteration runs for n independent tracks at once: one thread per track, way-points visited in order, joint
// vector, site frame, Jacobian and the small Gram matrix in registers.
//
// Not a translation of the numpy code:
//   * forward kinematics and the site Jacobian come from the link-local chain tables of DevModel (E_k, r_k: one 3x3
//     product per link; hinge axis = third column of the link's world rotation, anchor = its origin), not from
//     mj_kinematics + mj_comPos + mj_jacSite's cdof/subtree_com detour;
//   * the update is computed on the SMALLER Gram matrix: position-only targets (3 x nd Jacobian) use
//     J'(JJ' + reg I)^-1 err resp. J'(JJ')^+ err with a 3x3 factorisation, pose targets (6 x nd) use the nd x nd normal
//     equations.  (J'J + reg I)^-1 J' = J'(JJ' + reg I)^-1 and pinv(J'J) J' = pinv(J) = J' pinv(JJ') hold exactly and
//     J'J, JJ' share their non-zero spectrum, so numpy's `lstsq(J'J, J'err, rcond=-1)` cut (singular values
//     <= DBL_EPSILON * largest are dropped) is the same cut on either side;
//   * the pseudo-inverse is the plain inverse (LDL') whenever every pivot is far above the cut, and a cyclic Jacobi
//     eigen-decomposition of that symmetric matrix (3x3 or 6x6, unrolled) otherwise.
// Included at the end of so101_capi.cu (same translation unit: shares fail()/CUDA_TRY/DeviceGuard/So101Model).
#pragma once

namespace so101 {

struct IkModel {
  double E[NV][9], r[NV][3], qpos0[NV];
  double site[3];       // site position in the frame of link site_link
  double site_rot[9];   // site axes in that frame (row-major)
  int32_t site_link, pad_;
};

namespace ik {

constexpr double MINVAL = 1e-15;  // mjMINVAL
constexpr double EPS = 2.220446049250313e-16;
constexpr double PI = 3.14159265358979323846;

// mju_mat2Quat + mju_normalize4
__device__ __forceinline__ void mat2quat(const double* m, double* q) {
  if (m[0] + m[4] + m[8] > 0) {
    q[0] = 0.5 * sqrt(1 + m[0] + m[4] + m[8]);
    const double s = 0.25 / q[0];
    q[1] = s * (m[7] - m[5]); q[2] = s * (m[2] - m[6]); q[3] = s * (m[3] - m[1]);
  } else if (m[0] > m[4] && m[0] > m[8]) {
    q[1] = 0.5 * sqrt(1 + m[0] - m[4] - m[8]);
    const double s = 0.25 / q[1];
    q[0] = s * (m[7] - m[5]); q[2] = s * (m[1] + m[3]); q[3] = s * (m[2] + m[6]);
  } else if (m[4] > m[8]) {
    q[2] = 0.5 * sqrt(1 - m[0] + m[4] - m[8]);
    const double s = 0.25 / q[2];
    q[0] = s * (m[2] - m[6]); q[1] = s * (m[1] + m[3]); q[3] = s * (m[5] + m[7]);
  } else {
    q[3] = 0.5 * sqrt(1 - m[0] - m[4] + m[8]);
    const double s = 0.25 / q[3];
    q[0] = s * (m[3] - m[1]); q[1] = s * (m[2] + m[6]); q[2] = s * (m[5] + m[7]);
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else { const double i = 1 / n; q[0] *= i; q[1] *= i; q[2] *= i; q[3] *= i; }
}

// err_rot = mju_quat2Vel(target * conj(site), 1)
__device__ __forceinline__ void rot_error(const double* tq, const double* sq, double* e) {
  const double n0 = sq[0], n1 = -sq[1], n2 = -sq[2], n3 = -sq[3];
  const double w = tq[0] * n0 - tq[1] * n1 - tq[2] * n2 - tq[3] * n3;
  double x = tq[0] * n1 + tq[1] * n0 + tq[2] * n3 - tq[3] * n2;
  double y = tq[0] * n2 - tq[1] * n3 + tq[2] * n0 + tq[3] * n1;
  double z = tq[0] * n3 + tq[1] * n2 - tq[2] * n1 + tq[3] * n0;
  const double sn = sqrt(x * x + y * y + z * z);
  if (sn < MINVAL) { x = 1; y = 0; z = 0; }
  else { const double i = 1 / sn; x *= i; y *= i; z *= i; }
  double speed = 2 * atan2(sn, w);
  if (speed > PI) speed -= 2 * PI;
  e[0] = x * speed; e[1] = y * speed; e[2] = z * speed;
}

// x = (G + diag(add))^-1 b for a symmetric positive definite N x N system (LDL', unrolled).  Returns the smallest pivot.
template <int N>
__device__ __forceinline__ double spd_solve(const double (&G)[N][N], const double (&add)[N], const double (&b)[N],
                                            double (&x)[N]) {
  double L[N][N], d[N];
  double dmin = 1e300;
#pragma unroll
  for (int j = 0; j < N; j++) {
    double s = G[j][j] + add[j];
#pragma unroll
    for (int k = 0; k < j; k++) s -= L[j][k] * L[j][k] * d[k];
    d[j] = s;
    dmin = fmin(dmin, s);
    const double inv = 1 / s;
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      double t = G[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) t -= L[i][k] * L[j][k] * d[k];
      L[i][j] = t * inv;
    }
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
    double t = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) t -= L[i][k] * x[k];
    x[i] = t;
  }
#pragma unroll
  for (int i = 0; i < N; i++) x[i] /= d[i];
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    double t = x[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) t -= L[k][i] * x[k];
    x[i] = t;
  }
  return dmin;
}

// x = G^+ b with numpy's rcond=-1 cut.  When every LDL' pivot is far above the cut (WELL * largest diagonal entry; the
// cut itself is 2.2e-16 * lambda_max) no eigenvalue can be dropped, the pseudo-inverse IS the inverse and the
// factorisation's answer is returned (the usual case: ~100 instructions instead of ~1500 for the eigen-decomposition);
// otherwise - near a kinematic singularity, or masked-out dofs in the 6 x 6 case - the Jacobi path decides which
// directions exist.
template <int N>
__device__ __forceinline__ void sym_pinv_apply(double (&A)[N][N], const double (&b)[N], double (&x)[N]);
// `pad[i]` = 1 for rows/columns that are identically zero by construction (dofs that may not move): they are given a
// unit diagonal for the factorisation, which leaves x_i = 0 and the other unknowns untouched.
template <int N>
__device__ __forceinline__ void sym_solve_or_pinv(double (&G)[N][N], const double (&pad)[N], const double (&b)[N],
                                                  double (&x)[N]) {
  constexpr double WELL = 1e-6;
  double gmax = 0;
#pragma unroll
  for (int i = 0; i < N; i++) gmax = fmax(gmax, G[i][i]);
  const double dmin = spd_solve<N>(G, pad, b, x);
  if (dmin > WELL * gmax) return;       // false for NaN pivots too
  sym_pinv_apply<N>(G, b, x);
}

// x = G^+ b for symmetric G, eigenvalues with |lambda| <= DBL_EPSILON * max|lambda| dropped (numpy lstsq, rcond=-1).
// Cyclic Jacobi: G -> V' G V diagonal; b is rotated along (c = V' b), x = V diag^+ c.
template <int N>
__device__ __forceinline__ void sym_pinv_apply(double (&A)[N][N], const double (&b)[N], double (&x)[N]) {
  double V[N][N];
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 16; sweep++) {
    double off = 0, dg = 0;
#pragma unroll
    for (int p = 0; p < N; p++) {
      dg += A[p][p] * A[p][p];
#pragma unroll
      for (int q = p + 1; q < N; q++) off += A[p][q] * A[p][q];
    }
    if (off <= 1e-40 * dg) break;
#pragma unroll
    for (int p = 0; p < N - 1; p++) {
#pragma unroll
      for (int q = p + 1; q < N; q++) {
        const double apq = A[p][q];
        if (apq != 0.0) {
          const double theta = (A[q][q] - A[p][p]) / (2 * apq);
          const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1));
          const double c = rsqrt(t * t + 1), s = t * c;
          A[p][p] -= t * apq;
          A[q][q] += t * apq;
          A[p][q] = 0; A[q][p] = 0;
#pragma unroll
          for (int r = 0; r < N; r++) {
            if (r != p && r != q) {
              const double arp = A[r][p], arq = A[r][q];
              A[r][p] = c * arp - s * arq; A[p][r] = A[r][p];
              A[r][q] = s * arp + c * arq; A[q][r] = A[r][q];
            }
            const double vrp = V[r][p], vrq = V[r][q];
            V[r][p] = c * vrp - s * vrq;
            V[r][q] = s * vrp + c * vrq;
          }
        }
      }
    }
  }
  double lmax = 0;
#pragma unroll
  for (int i = 0; i < N; i++) lmax = fmax(lmax, fabs(A[i][i]));
  const double cut = EPS * lmax;
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = 0;
#pragma unroll
  for (int k = 0; k < N; k++) {
    double ck = 0;
#pragma unroll
    for (int i = 0; i < N; i++) ck += V[i][k] * b[i];
    const double w = fabs(A[k][k]) > cut ? ck / A[k][k] : 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) x[i] += V[i][k] * w;
  }
}

}  // namespace ik

// One thread per track.  POSE = the target has an orientation (6-row error) or not (3 rows).
template <bool POSE>
__global__ void __launch_bounds__(128)
k_ik_track(const __grid_constant__ IkModel m, const So101IkParams prm, const double* __restrict__ xyz,
           const double* __restrict__ quat, const double* __restrict__ q0, int P, int64_t n,
           double* __restrict__ q_out, int32_t* __restrict__ status, double* __restrict__ err_out) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  double q[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) q[k] = q0 ? q0[(int64_t)k * n + b] : m.qpos0[k];
  double tq[4] = {1, 0, 0, 0};
  if (POSE) {
#pragma unroll
    for (int c = 0; c < 4; c++) tq[c] = quat[(int64_t)c * n + b];
  }
  bool aborted = false;
  for (int p = 0; p < P; p++) {
    const int64_t o3 = ((int64_t)p * 3) * n + b;
    const double tx = xyz[o3], ty = xyz[o3 + n], tz = xyz[o3 + 2 * n];
    double qs[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) qs[k] = q[k];
    bool success = false;
    int steps = 0;
    double err_norm = 0;
    if (!aborted) {
      for (steps = 0; steps < prm.max_steps; steps++) {
        // ---- forward kinematics of the chain up to the site's link: axes, anchors, site frame
        double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pos[3] = {0, 0, 0};
        double ax[NV][3], an[NV][3];
#pragma unroll
        for (int k = 0; k < NV; k++) {
          if (k <= m.site_link) {
            double sn, cs;
            sincos(q[k] - m.qpos0[k], &sn, &cs);
            const double* E = m.E[k];
            const double* rk = m.r[k];
#pragma unroll
            for (int i = 0; i < 3; i++) pos[i] += R[3 * i] * rk[0] + R[3 * i + 1] * rk[1] + R[3 * i + 2] * rk[2];
            double RE[9];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
              for (int j = 0; j < 3; j++)
                RE[3 * i + j] = R[3 * i] * E[j] + R[3 * i + 1] * E[3 + j] + R[3 * i + 2] * E[6 + j];
#pragma unroll
            for (int i = 0; i < 3; i++) {
              R[3 * i] = cs * RE[3 * i] + sn * RE[3 * i + 1];
              R[3 * i + 1] = cs * RE[3 * i + 1] - sn * RE[3 * i];
              R[3 * i + 2] = RE[3 * i + 2];
              ax[k][i] = R[3 * i + 2];
              an[k][i] = pos[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 3; i++) { ax[k][i] = 0; an[k][i] = 0; }
          }
        }
        double sp[3];
#pragma unroll
        for (int i = 0; i < 3; i++)
          sp[i] = pos[i] + R[3 * i] * m.site[0] + R[3 * i + 1] * m.site[1] + R[3 * i + 2] * m.site[2];
        // ---- error
        constexpr int M = POSE ? 6 : 3;
        double e[M];
        e[0] = tx - sp[0]; e[1] = ty - sp[1]; e[2] = tz - sp[2];
        err_norm = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
        if (POSE) {
          double sm[9], sq[4], er[3];
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++)
              sm[3 * i + j] = R[3 * i] * m.site_rot[j] + R[3 * i + 1] * m.site_rot[3 + j] + R[3 * i + 2] * m.site_rot[6 + j];
          ik::mat2quat(sm, sq);
          ik::rot_error(tq, sq, er);
          e[M - 3] = er[0]; e[M - 2] = er[1]; e[M - 1] = er[2];
          err_norm += sqrt(er[0] * er[0] + er[1] * er[1] + er[2] * er[2]) * prm.rot_weight;
        }
        if (err_norm < prm.tol) { success = true; break; }
        // ---- site Jacobian: column k = [axis_k x (site - anchor_k) ; axis_k] for the moving dofs below the site
        double J[M][NV];
#pragma unroll
        for (int k = 0; k < NV; k++) {
          const bool on = ((prm.dof_mask >> k) & 1) && k <= m.site_link;
          const double dx = sp[0] - an[k][0], dy = sp[1] - an[k][1], dz = sp[2] - an[k][2];
          J[0][k] = on ? ax[k][1] * dz - ax[k][2] * dy : 0.0;
          J[1][k] = on ? ax[k][2] * dx - ax[k][0] * dz : 0.0;
          J[2][k] = on ? ax[k][0] * dy - ax[k][1] * dx : 0.0;
          if (POSE) {
            J[M - 3][k] = on ? ax[k][0] : 0.0; J[M - 2][k] = on ? ax[k][1] : 0.0; J[M - 1][k] = on ? ax[k][2] : 0.0;
          }
        }
        const bool reg = err_norm > prm.reg_threshold && prm.reg_strength > 0;
        double upd[NV];
        if (!POSE) {
          double G[3][3], y[3], e3[3] = {e[0], e[1], e[2]};
#pragma unroll
          for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) {
              double s = 0;
#pragma unroll
              for (int k = 0; k < NV; k++) s += J[i][k] * J[j][k];
              G[i][j] = s;
            }
          if (reg) {
            const double add[3] = {prm.reg_strength, prm.reg_strength, prm.reg_strength};
            ik::spd_solve<3>(G, add, e3, y);
          } else {
            const double pad[3] = {0, 0, 0};
            ik::sym_solve_or_pinv<3>(G, pad, e3, y);
          }
#pragma unroll
          for (int k = 0; k < NV; k++) upd[k] = J[0][k] * y[0] + J[1][k] * y[1] + J[2][k] * y[2];
        } else {
          double G[NV][NV], g[NV];
#pragma unroll
          for (int i = 0; i < NV; i++) {
            double s = 0;
#pragma unroll
            for (int r = 0; r < M; r++) s += J[r][i] * e[r];
            g[i] = s;
#pragma unroll
            for (int j = 0; j < NV; j++) {
              double t = 0;
#pragma unroll
              for (int r = 0; r < M; r++) t += J[r][i] * J[r][j];
              G[i][j] = t;
            }
          }
          double add[NV];
#pragma unroll
          for (int k = 0; k < NV; k++)
            add[k] = (((prm.dof_mask >> k) & 1) && k <= m.site_link) ? (reg ? prm.reg_strength : 0.0) : 1.0;
          if (reg) ik::spd_solve<NV>(G, add, g, upd);
          else ik::sym_solve_or_pinv<NV>(G, add, g, upd);
        }
        double un = 0;
#pragma unroll
        for (int k = 0; k < NV; k++) un += upd[k] * upd[k];
        un = sqrt(un);
        if (err_norm / un > prm.progress_thresh) break;
        const double sc = un > prm.max_update_norm ? prm.max_update_norm / un : 1.0;
#pragma unroll
        for (int k = 0; k < NV; k++) q[k] += upd[k] * sc;
      }
      if (steps == prm.max_steps) steps = prm.max_steps - 1;  // Python's loop variable after an exhausted range
      if (!success) {
        if (p == 0) aborted = true;  // the reference raises here [REF control/TrajectoryGenerator.py:207-208]
#pragma unroll
        for (int k = 0; k < NV; k++) q[k] = qs[k];
      }
    }
    // a failed solve repeats the previous answer (which q was restored to) [REF :198-204]
#pragma unroll
    for (int k = 0; k < NV; k++) q_out[((int64_t)p * NV + k) * n + b] = q[k];
    status[(int64_t)p * n + b] = (success ? 1 : 0) | (aborted ? 2 : 0) | (steps << 8);
    if (err_out) err_out[(int64_t)p * n + b] = err_norm;
  }
}

}  // namespace so101

extern "C" int so101_ik_track(const So101Model* model, const So101IkParams* params, const double* xyz,
                              const double* quat, const double* q0, int P, int64_t n, int device, double* q_out,
                              int32_t* status, double* err, void* stream) {
  using namespace so101;
  if (!model || !params) return fail(SO101_EINVAL, "null argument");
  if (P < 0 || n < 0) return fail(SO101_EINVAL, "ik_track: negative size");
  if (P > 0 && n > 0 && (!q_out || !status || !xyz)) return fail(SO101_EINVAL, "null argument");
  if (params->max_steps < 1 || !(params->tol >= 0)) return fail(SO101_EINVAL, "ik_track: need max_steps >= 1, tol >= 0");
  if ((params->dof_mask & ~((1 << NV) - 1)) != 0) return fail(SO101_EINVAL, "ik_track: dof_mask has bits beyond dof 5");
  const int ndev = so101_device_count();
  if (ndev <= 0) return fail(SO101_ENODEVICE, "no CUDA device visible: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(SO101_EINVAL, "device index out of range");
  if (P == 0 || n == 0) return SO101_OK;
  DeviceGuard g(device);
  if (!g.ok) return fail(SO101_ECUDA, "cudaSetDevice failed");
  IkModel im;
  std::memset(&im, 0, sizeof im);
  const DevModel<double>& d = model->d;
  const So101Tables& t = model->tables;
  std::memcpy(im.E, d.E, sizeof im.E);
  std::memcpy(im.r, d.r, sizeof im.r);
  std::memcpy(im.qpos0, d.qpos0, sizeof im.qpos0);
  std::memcpy(im.site, d.site, sizeof im.site);
  im.site_link = d.site_link;
  {  // site axes in the link frame F_k (z = hinge axis): A_k' * R(site_quat)
    double A[9], At[9], Rs[9];
    const double qn = t.site_quat[0] * t.site_quat[0] + t.site_quat[1] * t.site_quat[1] +
                      t.site_quat[2] * t.site_quat[2] + t.site_quat[3] * t.site_quat[3];
    if (!(qn > 0)) return fail(SO101_EMODEL, "ik_track: tables carry no site orientation (site_quat is zero)");
    hostbuild::z2vec(t.jnt_axis[d.site_link], A);
    hostbuild::mt(A, At);
    hostbuild::q2m(t.site_quat, Rs);
    hostbuild::mm(At, Rs, im.site_rot);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int blk = 128;
  const unsigned grid = (unsigned)((n + blk - 1) / blk);
  if (quat) k_ik_track<true><<<grid, blk, 0, st>>>(im, *params, xyz, quat, q0, P, n, q_out, status, err);
  else k_ik_track<false><<<grid, blk, 0, st>>>(im, *params, xyz, nullptr, q0, P, n, q_out, status, err);
  CUDA_TRY(cudaGetLastError());
  return SO101_OK;
}
