// so101_model.h — device-side constant tables of one compiled scene, and their host builder.
//
// The kernels do NOT follow MuJoCo's world-frame, COM-centred formulation (that is what
// oracle/so101_oracle.c restates).  They use link-local spatial algebra: every link k carries
// a frame F_k whose origin is the joint anchor and whose z axis is the hinge axis, so that
//   * the joint motion subspace is the constant S = [0 0 1 | 0 0 0],
//   * link inertias are constants (no per-step R*I*R' in world axes),
//   * forward kinematics is needed only for the observation site and the contact tripwire.
// The builder folds body_pos/body_quat/jnt_pos/jnt_axis/body_ipos/body_iquat/body_inertia of
// So101Tables into E_k (rotation F_k(pre-joint) -> F_{k-1}), r_k (origin of F_k in F_{k-1})
// and the 10-parameter spatial inertia of link k about the origin of F_k.
//
// DevModel<T> is passed to every kernel BY VALUE as a __grid_constant__ parameter: all table
// reads become constant-bank operands of the FMA instructions (no registers, no loads), and
// several models (scene A / scene B) can coexist in one process.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/so101_b200.h"

namespace so101 {

constexpr int NV = SO101_NV;
constexpr int TRIP_PER_LINK = 3;
// words of the per-env hull cache that belong to the self-collision test: budget of the critical pair (float bits), budget of
// all other pairs (float bits), the critical pair (16 a + b, -1: none yet)
constexpr int SELF_BUDGET_SLOT = SO101_MAXTRIP - 1, SELF_REST_SLOT = SO101_MAXTRIP - 2, SELF_PAIR_SLOT = SO101_MAXTRIP - 3;
constexpr int SELF_EXTRA = 4;      // colliding geoms without a tripwire box (self-collision test only)

template <typename T>
struct DevModel {
  // kinematic chain
  T E[NV][9];        // row-major; v_{k-1} = E_k * Rz(theta_k) * v_k
  T r[NV][3];        // origin of F_k expressed in F_{k-1} (k = 0: in the world frame)
  T I[NV][10];       // [Ixx Iyy Izz Ixy Ixz Iyz hx hy hz m] about the origin of F_k, in F_k axes
  T qpos0[NV];
  T armature[NV], damping[NV], stiffness[NV], qspring[NV];
  T accg[3];         // -gravity (spatial acceleration of the world frame)
  // friction-loss rows (constant per model)
  T fr_f[NV];        // frictionloss (0 => row absent)
  T fr_R[NV], fr_D[NV], fr_Rf[NV], fr_hRff[NV], fr_B[NV];
  // joint-limit rows
  T lim_lo[NV], lim_hi[NV], lim_margin[NV], lim_invw[NV], lim_K[NV], lim_B[NV];
  T lim_imp[NV][5];
  // actuators, re-indexed by dof
  T act_gain[NV], act_b0[NV], act_b1[NV], act_b2[NV], act_gear[NV];
  T ctrl_lo[NV], ctrl_hi[NV], frc_lo[NV], frc_hi[NV];
  // observation site in the frame of link site_link
  T site[3];
  // option / solver
  T h, tolerance, gtol_fac, scale;
  // contact tripwire: up to TRIP_PER_LINK oriented boxes per link, in the link frame
  T trip_c[NV][TRIP_PER_LINK][3], trip_ax[NV][TRIP_PER_LINK][9], trip_half[NV][TRIP_PER_LINK][3];
  T trip_rad[NV][TRIP_PER_LINK];   // |half|_2: bounding-sphere pre-check
  T trip_z, trip_qlo[NV], trip_qhi[NV];
  // table-plane contact rows (so101_contact.cuh): sliding friction, K / B of the contact solref, impedance
  // parameters, includemargin, translational body_invweight0 of every link, footprint of the table top
  T con_mu, con_K, con_B, con_margin, con_imp[5], con_tran[NV], con_box[4];
  T con_tilt[NV][3];   // tie-breaking tilt of the support direction (1e-7 * (2,3,4)/sqrt(29) in the BODY frame) in link axes
  // self-collision test (self_boxes_overlap): the boxes of the colliding geoms that have no tripwire box (link frame); the
  // others are the tripwire boxes above.  Kept small: with these tables the float32 kernels' parameters stay below 4 KB
  // (beyond it the float32 one-warp kernel lost 35 %).
  T sx_c[SELF_EXTRA][3], sx_ax[SELF_EXTRA][9], sx_half[SELF_EXTRA][3], sx_rad[SELF_EXTRA];
  T self_rlen[NV];   // |r_k|: with the distance of a box's far corner from its link origin, the lever arm of the joints between two boxes
  int32_t trip_n[NV];   // <- first non-T field (see hostbuild::convert)
  int32_t trip_geom[NV][TRIP_PER_LINK];   // hull (= tripwire box) index of slot b of link k
  // device pointers of the hull data as (lo, hi) words - int32 so that DevModel<double> and DevModel<float> share the
  // layout of this tail; patched per batch (so101_batch_create), 0 = no hulls: vert, adj_start, adj, cube
  int32_t hull_ptr[8];
  int32_t hull_res, con_enabled;
  int32_t ntrip;
  int32_t site_link;
  int32_t iterations, ls_iterations;
  int32_t limited_mask, ctrllim_mask, frclim_mask, nfriction, ctrl_of_dof[NV];
  // flat list of all boxes for the self-collision test: link, and slot b of trip_*[link][b] (>= 0) or -(1 + index into sx_*)
  int32_t sb_n, sb_link[SO101_MAXTRIP], sb_slot[SO101_MAXTRIP];
  int32_t self_boxes;   // 1: poses outside the joint box get the box-box test; 0: they are flagged (SO101_OPT_SELF_TEST)
  int32_t any_damping, any_stiffness, pad_;
};

namespace hostbuild {

inline void q2m(const double q[4], double m[9]) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double w = q[0] / n, x = q[1] / n, y = q[2] / n, z = q[3] / n;
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
inline void mm(const double a[9], const double b[9], double c[9]) {
  double t[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  std::memcpy(c, t, sizeof t);
}
inline void mt(const double a[9], double c[9]) {
  double t[9] = {a[0], a[3], a[6], a[1], a[4], a[7], a[2], a[5], a[8]};
  std::memcpy(c, t, sizeof t);
}
inline void mv(const double a[9], const double v[3], double o[3]) {
  double t[3] = {a[0] * v[0] + a[1] * v[1] + a[2] * v[2], a[3] * v[0] + a[4] * v[1] + a[5] * v[2],
                 a[6] * v[0] + a[7] * v[1] + a[8] * v[2]};
  std::memcpy(o, t, sizeof t);
}
// rotation taking z to the unit vector a (identity when a == z)
inline void z2vec(const double a[3], double A[9]) {
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double s = std::sqrt(a[0] * a[0] + a[1] * a[1]);
  if (s < 1e-14) {
    if (a[2] > 0) { std::memcpy(A, I, sizeof I); return; }
    const double F[9] = {1, 0, 0, 0, -1, 0, 0, 0, -1};  // half turn about x
    std::memcpy(A, F, sizeof F);
    return;
  }
  // Rodrigues about k = z x a / |z x a| by angle acos(a_z)
  double kx = -a[1] / s, ky = a[0] / s, c = a[2], sn = s, v = 1 - c;
  A[0] = c + kx * kx * v; A[1] = kx * ky * v;     A[2] = ky * sn;
  A[3] = kx * ky * v;     A[4] = c + ky * ky * v; A[5] = -kx * sn;
  A[6] = -ky * sn;        A[7] = kx * sn;         A[8] = c;
}

// getimpedance of engine_core_constraint.c (host side, for the constant friction rows)
inline double impedance(const double* solimp, double pos, double margin) {
  if (solimp[0] == solimp[1] || solimp[2] <= 1e-15) return 0.5 * (solimp[0] + solimp[1]);
  double x = std::fabs((pos - margin) / solimp[2]);
  if (x >= 1 || x <= 0) return x >= 1 ? solimp[1] : solimp[0];
  double y;
  if (solimp[4] == 1) y = x;
  else if (x <= solimp[3]) y = std::pow(x, solimp[4]) / std::pow(solimp[3], solimp[4] - 1);
  else y = 1 - std::pow(1 - x, solimp[4]) / std::pow(1 - solimp[3], solimp[4] - 1);
  return solimp[0] + y * (solimp[1] - solimp[0]);
}

// Returns "" on success, else the reason the model is outside the supported subset.
inline std::string build(const So101Tables& t, DevModel<double>& m) {
  std::memset(&m, 0, sizeof m);
  if (t.abi_version != SO101_ABI_VERSION) return "tables ABI version mismatch";
  if (t.nv != NV || t.nu != NV) return "model must have exactly 6 hinge dofs and 6 actuators";
  if (t.nbody < NV + 1 || t.nbody > SO101_MAXBODY) return "unsupported body count";
  if (t.ntrip < 0 || t.ntrip > SO101_MAXTRIP) return "ntrip exceeds SO101_MAXTRIP: a truncated tripwire could miss a contact";
  if (t.nself < 0 || t.ntrip + t.nself > SO101_MAXTRIP) return "ntrip + nself exceeds SO101_MAXTRIP";
  if (t.ntrip > SELF_PAIR_SLOT) return "ntrip leaves no room for the self-collision budget words";
  // ---- chain check: link k = jnt_body[k]; parent(link k) == link k-1; link 0 hangs off fixed bodies
  int link[NV];
  for (int k = 0; k < NV; k++) {
    link[k] = t.jnt_body[k];
    if (link[k] <= 0 || link[k] >= t.nbody || t.body_jnt[link[k]] != k) return "joint/body tables inconsistent";
    if (k > 0 && t.body_parent[link[k]] != link[k - 1]) return "kinematic tree is not a serial chain";
  }
  for (int b = 1; b < t.nbody; b++) {
    bool is_link = false;
    for (int k = 0; k < NV; k++) is_link |= (link[k] == b);
    if (!is_link && t.body_jnt[b] >= 0) return "joint on a non-chain body";
    if (!is_link) {  // fixed bodies are only allowed between the world and link 0
      bool on_root_path = false;
      for (int a = t.body_parent[link[0]]; a > 0; a = t.body_parent[a]) on_root_path |= (a == b);
      if (!on_root_path) return "fixed body off the root path (would need inertia merging)";
    }
  }
  // ---- world pose of the fixed parent of link 0
  double Rb[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pb[3] = {0, 0, 0};
  {
    int path[SO101_MAXBODY], n = 0;
    for (int a = t.body_parent[link[0]]; a > 0; a = t.body_parent[a]) path[n++] = a;
    for (int i = n - 1; i >= 0; i--) {
      double Rc[9], o[3];
      q2m(t.body_quat[path[i]], Rc);
      mv(Rb, t.body_pos[path[i]], o);
      for (int c = 0; c < 3; c++) pb[c] += o[c];
      mm(Rb, Rc, Rb);
    }
  }
  double Aprev[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, jprev[3] = {0, 0, 0};
  for (int k = 0; k < NV; k++) {
    int b = link[k];
    double Rc[9], A[9], At[9], Apt[9], tmp[9], E[9];
    q2m(t.body_quat[b], Rc);
    z2vec(t.jnt_axis[k], A);
    mt(A, At);
    mt(Aprev, Apt);
    mm(Rc, A, tmp);
    mm(Apt, tmp, E);  // E_k = A_{k-1}^T Rc_k A_k
    double off[3], rj[3], r[3];
    mv(Rc, t.jnt_pos[k], rj);
    for (int c = 0; c < 3; c++) off[c] = t.body_pos[b][c] + rj[c] - jprev[c];
    mv(Apt, off, r);
    if (k == 0) {  // fold the fixed base pose: frame "-1" is the world
      double E0[9], r0[3];
      mm(Rb, E, E0);
      mv(Rb, r, r0);
      for (int c = 0; c < 3; c++) r0[c] += pb[c];
      std::memcpy(E, E0, sizeof E0);
      std::memcpy(r, r0, sizeof r0);
    }
    for (int c = 0; c < 9; c++) m.E[k][c] = E[c];
    for (int c = 0; c < 3; c++) m.r[k][c] = r[c];
    // spatial inertia of link k about the origin of F_k, in F_k axes
    double Ri[9], AtRi[9], d3[3], com[3];
    q2m(t.body_iquat[b], Ri);
    mm(At, Ri, AtRi);
    for (int c = 0; c < 3; c++) d3[c] = t.body_ipos[b][c] - t.jnt_pos[k][c];
    mv(At, d3, com);
    double Ic[9] = {0};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        for (int p = 0; p < 3; p++) Ic[3 * i + j] += AtRi[3 * i + p] * t.body_inertia[b][p] * AtRi[3 * j + p];
    double mass = t.body_mass[b], cc = com[0] * com[0] + com[1] * com[1] + com[2] * com[2];
    m.I[k][0] = Ic[0] + mass * (cc - com[0] * com[0]);
    m.I[k][1] = Ic[4] + mass * (cc - com[1] * com[1]);
    m.I[k][2] = Ic[8] + mass * (cc - com[2] * com[2]);
    m.I[k][3] = Ic[1] - mass * com[0] * com[1];
    m.I[k][4] = Ic[2] - mass * com[0] * com[2];
    m.I[k][5] = Ic[5] - mass * com[1] * com[2];
    m.I[k][6] = mass * com[0]; m.I[k][7] = mass * com[1]; m.I[k][8] = mass * com[2];
    m.I[k][9] = mass;
    if (b == t.site_body) {
      double s3[3];
      for (int c = 0; c < 3; c++) s3[c] = t.site_pos[c] - t.jnt_pos[k][c];
      mv(At, s3, m.site);
      m.site_link = k;
    }
    for (int i = 0; i < t.ntrip && i < SO101_MAXTRIP; i++) {
      if (t.trip_body[i] != b) continue;
      int slot = m.trip_n[k];
      if (slot >= TRIP_PER_LINK) return "more than 3 tripwire boxes on one link";
      double c3[3];
      for (int c = 0; c < 3; c++) c3[c] = t.trip_center[i][c] - t.jnt_pos[k][c];
      mv(At, c3, m.trip_c[k][slot]);
      for (int a = 0; a < 3; a++) mv(At, &t.trip_axes[i][3 * a], &m.trip_ax[k][slot][3 * a]);
      for (int c = 0; c < 3; c++) m.trip_half[k][slot][c] = t.trip_half[i][c];
      m.trip_geom[k][slot] = i;
      m.trip_rad[k][slot] = std::sqrt(t.trip_half[i][0] * t.trip_half[i][0] + t.trip_half[i][1] * t.trip_half[i][1] +
                                      t.trip_half[i][2] * t.trip_half[i][2]) * (1 + 1e-5);  // margin covers f32 rounding of |zw|
      m.trip_n[k] = slot + 1;
      m.sb_link[m.sb_n] = k; m.sb_slot[m.sb_n] = slot; m.sb_n++;
    }
    for (int i = t.ntrip; i < t.ntrip + t.nself; i++) {   // geoms without a tripwire box: self-collision test only
      if (t.trip_body[i] != b) continue;
      int g = 0;
      while (g < SELF_EXTRA && m.sx_rad[g] != 0) g++;
      if (g == SELF_EXTRA) return "more than 4 colliding geoms without a tripwire box";
      double c3[3];
      for (int c = 0; c < 3; c++) c3[c] = t.trip_center[i][c] - t.jnt_pos[k][c];
      mv(At, c3, m.sx_c[g]);
      for (int a = 0; a < 3; a++) mv(At, &t.trip_axes[i][3 * a], &m.sx_ax[g][3 * a]);
      for (int c = 0; c < 3; c++) m.sx_half[g][c] = t.trip_half[i][c];
      m.sx_rad[g] = std::sqrt(t.trip_half[i][0] * t.trip_half[i][0] + t.trip_half[i][1] * t.trip_half[i][1] +
                              t.trip_half[i][2] * t.trip_half[i][2]) * (1 + 1e-5);
      m.sb_link[m.sb_n] = k; m.sb_slot[m.sb_n] = -(1 + g); m.sb_n++;
    }
    std::memcpy(Aprev, A, sizeof A);
    std::memcpy(jprev, t.jnt_pos[k], sizeof jprev);
  }
  {
    bool site_ok = false;
    for (int k = 0; k < NV; k++) site_ok |= (link[k] == t.site_body);
    if (!site_ok) return "observation site must sit on a chain link";
  }
  m.ntrip = t.ntrip < SO101_MAXTRIP ? t.ntrip : SO101_MAXTRIP;
  for (int i = 0; i < t.ntrip + t.nself; i++) {
    bool ok = false;
    for (int k = 0; k < NV; k++) ok |= (link[k] == t.trip_body[i]);
    if (!ok) return "tripwire box on a non-chain body";
  }
  for (int k = 0; k < NV; k++) m.self_rlen[k] = std::sqrt(m.r[k][0] * m.r[k][0] + m.r[k][1] * m.r[k][1] + m.r[k][2] * m.r[k][2]);
  m.self_boxes = m.sb_n > 0;
  m.trip_z = t.trip_plane_z;
  for (int k = 0; k < NV; k++) { m.trip_qlo[k] = t.trip_qbox[k][0]; m.trip_qhi[k] = t.trip_qbox[k][1]; }

  // contact rows: K, B as mj_makeImpedance derives them from the contact's solref (refsafe), see the limit rows below
  if (t.con_enabled) {
    if (t.con_condim != 3) return "contact: only condim 3";
    if (t.con_solimp[4] != 1.0 && t.con_solimp[4] != 2.0) return "contact solimp power must be 1 or 2 on the CUDA path";
    double c0 = t.con_solref[0], c1 = t.con_solref[1], cd = t.con_solimp[1];
    if (c0 > 0 && c0 < 2 * t.timestep) c0 = 2 * t.timestep;
    if (c0 > 0) { m.con_K = 1 / std::fmax(1e-15, cd * cd * c0 * c0 * c1 * c1); m.con_B = 2 / std::fmax(1e-15, cd * c0); }
    else { m.con_K = -c0 / std::fmax(1e-15, cd * cd); m.con_B = -c1 / std::fmax(1e-15, cd); }
    m.con_mu = t.con_friction[0];
    m.con_margin = t.con_margin;
    for (int c = 0; c < 5; c++) m.con_imp[c] = t.con_solimp[c];
    for (int c = 0; c < 4; c++) m.con_box[c] = t.con_box[c];
    for (int k = 0; k < NV; k++) {
      m.con_tran[k] = t.body_invweight0[0][0] + t.body_invweight0[link[k]][0];
      const double tilt[3] = {2e-7 / 5.385164807134504, 3e-7 / 5.385164807134504, 4e-7 / 5.385164807134504};
      double A[9], At[9];
      z2vec(t.jnt_axis[k], A);
      mt(A, At);
      mv(At, tilt, m.con_tilt[k]);
    }
    m.con_enabled = 1;
  }
  for (int c = 0; c < 3; c++) m.accg[c] = -t.gravity[c];
  m.h = t.timestep;
  m.tolerance = t.tolerance;
  m.iterations = t.iterations;
  m.ls_iterations = t.ls_iterations;
  m.scale = 1.0 / (t.meaninertia * NV);
  // gtol = tolerance * ls_tolerance * |search| / scale
  m.gtol_fac = t.tolerance * t.ls_tolerance / m.scale;
  for (int k = 0; k < NV; k++) {
    m.qpos0[k] = t.qpos0[k];
    m.armature[k] = t.dof_armature[k];
    m.damping[k] = t.dof_damping[k];
    m.stiffness[k] = t.jnt_stiffness[k];
    m.qspring[k] = t.qpos_spring[k];
    if (t.dof_damping[k] > 0) m.any_damping = 1;
    if (t.jnt_stiffness[k] != 0) m.any_stiffness = 1;
    // friction row k
    double f = t.dof_frictionloss[k] > 0 ? t.dof_frictionloss[k] : 0.0;
    double sr0 = t.dof_solref[k][0], sr1 = t.dof_solref[k][1];
    if (sr0 > 0 && sr0 < 2 * t.timestep) sr0 = 2 * t.timestep;
    double imp = impedance(t.dof_solimp[k], 0.0, 0.0);
    double R = std::fmax(1e-15, (1 - imp) * t.dof_invweight0[k] / imp);
    double dmax = t.dof_solimp[k][1];
    double B = sr0 > 0 ? 2 / std::fmax(1e-15, dmax * sr0) : -sr1 / std::fmax(1e-15, dmax);
    m.fr_f[k] = f; m.fr_R[k] = R; m.fr_D[k] = f > 0 ? 1 / R : 0.0; m.fr_Rf[k] = R * f;
    m.fr_hRff[k] = 0.5 * R * f * f; m.fr_B[k] = B;
    if (f > 0) m.nfriction++;
    // limit row k
    if (t.jnt_limited[k]) {
      m.limited_mask |= 1 << k;
      if (t.jnt_solimp[k][4] != 1.0 && t.jnt_solimp[k][4] != 2.0)
        return "joint-limit solimp power must be 1 or 2 on the CUDA path";
      if (!(t.jnt_range[k][1] - t.jnt_range[k][0] > 2 * t.jnt_margin[k]))
        return "joint range narrower than 2*margin (both limit rows could be active)";
    }
    m.lim_lo[k] = t.jnt_range[k][0]; m.lim_hi[k] = t.jnt_range[k][1];
    m.lim_margin[k] = t.jnt_margin[k]; m.lim_invw[k] = t.dof_invweight0[k];
    double l0 = t.jnt_solref[k][0], l1 = t.jnt_solref[k][1], ldmax = t.jnt_solimp[k][1];
    if (l0 > 0 && l0 < 2 * t.timestep) l0 = 2 * t.timestep;
    if (l0 > 0) {
      m.lim_K[k] = 1 / std::fmax(1e-15, ldmax * ldmax * l0 * l0 * l1 * l1);
      m.lim_B[k] = 2 / std::fmax(1e-15, ldmax * l0);
    } else {
      m.lim_K[k] = -l0 / std::fmax(1e-15, ldmax * ldmax);
      m.lim_B[k] = -l1 / std::fmax(1e-15, ldmax);
    }
    for (int c = 0; c < 5; c++) m.lim_imp[k][c] = t.jnt_solimp[k][c];
    m.ctrl_of_dof[k] = -1;
  }
  for (int i = 0; i < NV; i++) {  // actuators re-indexed by dof (exactly one per dof)
    int k = t.act_dof[i];
    if (k < 0 || k >= NV || m.ctrl_of_dof[k] != -1) return "need exactly one actuator per joint";
    m.ctrl_of_dof[k] = i;
    m.act_gain[k] = t.act_gain[i]; m.act_gear[k] = t.act_gear[i];
    m.act_b0[k] = t.act_bias[i][0]; m.act_b1[k] = t.act_bias[i][1]; m.act_b2[k] = t.act_bias[i][2];
    m.ctrl_lo[k] = t.act_ctrlrange[i][0]; m.ctrl_hi[k] = t.act_ctrlrange[i][1];
    m.frc_lo[k] = t.act_forcerange[i][0]; m.frc_hi[k] = t.act_forcerange[i][1];
    if (t.act_ctrllimited[i]) m.ctrllim_mask |= 1 << k;
    if (t.act_forcelimited[i]) m.frclim_mask |= 1 << k;
    if (m.ctrl_of_dof[k] != k) return "actuator i must drive joint i (ctrl rows are indexed by joint)";
  }
  return "";
}

// hull vertex of a geom on body `body` (body frame) -> the frame of its link (origin = joint anchor, z = hinge axis)
inline bool hull_to_link(const So101Tables& t, int body, const double v[3], double out[3]) {
  for (int k = 0; k < NV; k++) {
    if (t.jnt_body[k] != body) continue;
    double A[9], At[9], d3[3];
    z2vec(t.jnt_axis[k], A);
    mt(A, At);
    for (int c = 0; c < 3; c++) d3[c] = v[c] - t.jnt_pos[k][c];
    mv(At, d3, out);
    return true;
  }
  return false;
}

template <typename T>
inline void convert(const DevModel<double>& a, DevModel<T>& b) {
  // field-wise cast; the two instantiations share the layout up to the element type
  const int nT = offsetof(DevModel<double>, trip_n) / sizeof(double);
  const double* src = reinterpret_cast<const double*>(&a);
  T* dst = reinterpret_cast<T*>(&b);
  for (int i = 0; i < nT; i++) dst[i] = static_cast<T>(src[i]);
  // integer tail: from trip_n up to and including pad_ (identical layout in both instantiations)
  const size_t tail = offsetof(DevModel<double>, pad_) + sizeof(int32_t) - offsetof(DevModel<double>, trip_n);
  std::memcpy(reinterpret_cast<char*>(&b) + offsetof(DevModel<T>, trip_n),
              reinterpret_cast<const char*>(&a) + offsetof(DevModel<double>, trip_n), tail);
}

}  // namespace hostbuild
}  // namespace so101
