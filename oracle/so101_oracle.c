/*
 * so101_oracle.c — CPU restatement (fp64, one env at a time) of mujoco.mj_step / mj_forward
 * for hinge trees, as called by the reference at SOARM101/SOARM101_Env.py:87,102,132.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product path (the CUDA
 * library behind include/so101_b200.h) never calls it and has no CPU fallback.
 *
 * PARITY UNPINNED.  The arithmetic of this path lives in the third-party `mujoco` PyPI wheel
 * (not vendored in the reference, no version pin; usage implies >= 2.3.3, artifacts dated
 * 2025-11 => most likely 3.3.x).  MuJoCo cannot be installed in the build container (no
 * wheel, no network) and the reference ships no golden vectors, tests or datasets for this
 * path.  What follows restates the published algorithm of MuJoCo 3.x function by function:
 *
 *   mj_kinematics, mj_comPos, mj_crb, mj_factorM, mj_solveM      engine_core_smooth.c
 *   mj_comVel, mj_rne, mj_passive                                 engine_core_smooth.c / engine_passive.c
 *   mj_makeConstraint (friction-loss + joint-limit rows),
 *   mj_makeImpedance, mj_referenceConstraint, mj_constraintUpdate engine_core_constraint.c
 *   mj_fwdActuation, mj_fwdAcceleration, mj_fwdConstraint,
 *   mj_Euler (implicit joint damping), mj_checkPos/Vel/Acc        engine_forward.c
 *   mj_solNewton -> mj_solPrimal, PrimalSearch (exact line search) engine_solver.c
 *
 * Each function below names the MuJoCo routine it follows.  What is pinned in-container:
 * the SURVEY.md Appendix-B anchors (mass, site positions, M(0), gravity bias), physics
 * invariants, and statistical fingerprints of the reference's trained Koopman model, the only
 * artefact in the tree that has seen real MuJoCo output: one-step MSE 7.5e-7 (training loss
 * 9.4e-7) and the published 200-step open-loop MAE (6.84e-3) reproduced at 6.8e-3
 * (tests/test_oracle.py).  No bit-level golden vector of mj_step exists, hence "unpinned".
 * Collision detection is restated for ONE case, the one the hot-path scene reaches: the convex
 * hull of a mesh geom against the top face of the static table box (mjc_Convex returns one contact per geom pair:
 * normal = the face normal, dist = lowest hull point - face height, pos = midway between the two witness points),
 * with mj_instantiateContact (condim 3, pyramidal cone: rows J_n +- mu J_t), mj_diagApprox and the pyramidal
 * R = 2 mu^2 R[first] adjustment of mj_makeImpedance.  Mesh-mesh self collision is not restated (tripwire flag).
 *
 * Deliberately generic (body_parent tree, sparse qM in MuJoCo's dof_Madr layout, dense efc_J)
 * so that it shares no structure with the chain-specialised CUDA kernels it checks.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/so101_b200.h"

#define NV SO101_NV
#define NB SO101_MAXBODY
#define NM 21           /* max entries of sparse qM (serial chain of 6) */
#define MAXCON 10       /* one contact per colliding geom (SO101_MAXTRIP boxes at most; the reference scene has 10) */
#define MAXEFC 64       /* 6 friction + 12 limit rows + 4 rows per contact, padded */
#define mjMINVAL 1e-15
#define mjMAXVAL 1e10

enum { ST_SATISFIED = 0, ST_QUADRATIC = 1, ST_LINEARNEG = 2, ST_LINEARPOS = 3 };
enum { ROW_FRICTION = 0, ROW_LIMIT = 1, ROW_CONTACT = 2 };

typedef struct OracleData {
  /* state */
  double qpos[NV], qvel[NV], qacc_warmstart[NV], ctrl[NV], qfrc_applied[NV], time;
  /* position-dependent */
  double xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], ximat[NB][9];
  double xanchor[NV][3], xaxis[NV][3], site_xpos[3];
  double subtree_com[NB][3], cinert[NB][10], crb[NB][10], cdof[NV][6];
  double qM[NM], qLD[NM], qLDiagInv[NV];
  /* velocity-dependent */
  double cvel[NB][6], cdof_dot[NV][6];
  double qfrc_bias[NV], qfrc_passive[NV], qfrc_actuator[NV], actuator_force[NV];
  double qfrc_smooth[NV], qacc_smooth[NV], qfrc_constraint[NV], qacc[NV];
  /* constraints */
  int32_t nefc, nf;
  int32_t efc_type[MAXEFC], efc_id[MAXEFC], efc_state[MAXEFC];
  double efc_J[MAXEFC][NV], efc_pos[MAXEFC], efc_margin[MAXEFC], efc_frictionloss[MAXEFC];
  double efc_diagApprox[MAXEFC], efc_R[MAXEFC], efc_D[MAXEFC], efc_KBIP[MAXEFC][4];
  double efc_vel[MAXEFC], efc_aref[MAXEFC], efc_b[MAXEFC], efc_force[MAXEFC];
  /* diagnostics */
  int32_t solver_niter, solver_nls, warning_bad, used_warmstart;
  double solver_cost;
  /* tree bookkeeping (filled by so101o_init) */
  int32_t dof_parent[NV], dof_Madr[NV], dof_body[NV], body_root[NB], nM;
  double body_subtreemass[NB];
  int32_t ls_evals_iter[8];   /* diagnostics: PrimalEval count of the first 8 line searches of the last solve */
  /* contacts (mj_collision for hull-vs-table-top; filled only when hulls are loaded, so101o_set_hulls) */
  int32_t ncon, con_unsupported;            /* unsupported: a contact beyond the table's footprint */
  int32_t con_geom[MAXCON], con_vert[MAXCON];
  double con_dist[MAXCON], con_pos[MAXCON][3], con_frame[MAXCON][9], con_gap[MAXCON];
} OracleData;

/* ---------------------------------------------------------------------------------------- */
/* small vector helpers (engine_util_blas.c / engine_util_spatial.c)                          */
/* ---------------------------------------------------------------------------------------- */
static double dotn(const double* a, const double* b, int n) {
  double s = 0;
  for (int i = 0; i < n; i++) s += a[i] * b[i];
  return s;
}
static void cross3(double r[3], const double a[3], const double b[3]) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}
static void mulMatVec3(double r[3], const double m[9], const double v[3]) {
  r[0] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  r[1] = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  r[2] = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
}
/* mju_mulQuat */
static void mulQuat(double r[4], const double a[4], const double b[4]) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                 a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                 a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  memcpy(r, t, sizeof t);
}
/* mju_rotVecQuat */
static void rotVecQuat(double r[3], const double v[3], const double q[4]) {
  if (v[0] == 0 && v[1] == 0 && v[2] == 0) { r[0] = r[1] = r[2] = 0; return; }
  if (q[0] == 1 && q[1] == 0 && q[2] == 0 && q[3] == 0) { r[0] = v[0]; r[1] = v[1]; r[2] = v[2]; return; }
  double t[3] = {q[0] * v[0] + q[2] * v[2] - q[3] * v[1],
                 q[0] * v[1] + q[3] * v[0] - q[1] * v[2],
                 q[0] * v[2] + q[1] * v[1] - q[2] * v[0]};
  double o[3] = {v[0] + 2 * (q[2] * t[2] - q[3] * t[1]),
                 v[1] + 2 * (q[3] * t[0] - q[1] * t[2]),
                 v[2] + 2 * (q[1] * t[1] - q[2] * t[0])};
  r[0] = o[0]; r[1] = o[1]; r[2] = o[2];
}
/* mju_axisAngle2Quat */
static void axisAngle2Quat(double q[4], const double axis[3], double angle) {
  if (angle == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double s = sin(angle * 0.5);
  q[0] = cos(angle * 0.5);
  q[1] = axis[0] * s; q[2] = axis[1] * s; q[3] = axis[2] * s;
}
/* mju_normalize4 */
static void normalize4(double q[4]) {
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n < mjMINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
  else if (fabs(n - 1) > mjMINVAL) { double s = 1 / n; q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s; }
}
/* mju_quat2Mat */
static void quat2Mat(double m[9], const double q[4]) {
  if (q[0] == 1 && q[1] == 0 && q[2] == 0 && q[3] == 0) {
    m[0] = 1; m[1] = 0; m[2] = 0; m[3] = 0; m[4] = 1; m[5] = 0; m[6] = 0; m[7] = 0; m[8] = 1;
    return;
  }
  double q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  double q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  double q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[4] = q00 - q11 + q22 - q33; m[8] = q00 - q11 - q22 + q33;
  m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02);
  m[3] = 2 * (q12 + q03); m[5] = 2 * (q23 - q01);
  m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01);
}
/* mju_inertCom */
static void inertCom(double res[10], const double inert[3], const double mat[9], const double dif[3],
                     double mass) {
  double tmp[9] = {mat[0] * inert[0], mat[3] * inert[0], mat[6] * inert[0],
                   mat[1] * inert[1], mat[4] * inert[1], mat[7] * inert[1],
                   mat[2] * inert[2], mat[5] * inert[2], mat[8] * inert[2]};
  res[0] = mat[0] * tmp[0] + mat[1] * tmp[3] + mat[2] * tmp[6];
  res[1] = mat[3] * tmp[1] + mat[4] * tmp[4] + mat[5] * tmp[7];
  res[2] = mat[6] * tmp[2] + mat[7] * tmp[5] + mat[8] * tmp[8];
  res[3] = mat[0] * tmp[1] + mat[1] * tmp[4] + mat[2] * tmp[7];
  res[4] = mat[0] * tmp[2] + mat[1] * tmp[5] + mat[2] * tmp[8];
  res[5] = mat[3] * tmp[2] + mat[4] * tmp[5] + mat[5] * tmp[8];
  res[0] += mass * (dif[1] * dif[1] + dif[2] * dif[2]);
  res[1] += mass * (dif[0] * dif[0] + dif[2] * dif[2]);
  res[2] += mass * (dif[0] * dif[0] + dif[1] * dif[1]);
  res[3] -= mass * dif[0] * dif[1];
  res[4] -= mass * dif[0] * dif[2];
  res[5] -= mass * dif[1] * dif[2];
  res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2];
  res[9] = mass;
}
/* mju_mulInertVec */
static void mulInertVec(double r[6], const double i[10], const double v[6]) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
/* mju_crossMotion / mju_crossForce */
static void crossMotion(double r[6], const double vel[6], const double v[6]) {
  double a[3], b[3];
  cross3(r, vel, v);
  cross3(a, vel, v + 3);
  cross3(b, vel + 3, v);
  r[3] = a[0] + b[0]; r[4] = a[1] + b[1]; r[5] = a[2] + b[2];
}
static void crossForce(double r[6], const double vel[6], const double f[6]) {
  double a[3], b[3];
  cross3(a, vel, f);
  cross3(b, vel + 3, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  cross3(r + 3, vel, f + 3);
}

/* ---------------------------------------------------------------------------------------- */
/* tree bookkeeping                                                                           */
/* ---------------------------------------------------------------------------------------- */
size_t so101o_sizeof_data(void) { return sizeof(OracleData); }

int so101o_init(const So101Tables* m, OracleData* d) {
  memset(d, 0, sizeof *d);
  if (m->nv != NV || m->nbody > NB || m->nbody < 2) return -1;
  for (int j = 0; j < NV; j++) d->dof_body[j] = m->jnt_body[j];
  /* dof_parentid: dof of the nearest ancestor body that has a joint */
  for (int j = 0; j < NV; j++) {
    int b = m->body_parent[m->jnt_body[j]], p = -1;
    while (b > 0) { if (m->body_jnt[b] >= 0) { p = m->body_jnt[b]; break; } b = m->body_parent[b]; }
    if (p >= j) return -2; /* dofs must be numbered parents first */
    d->dof_parent[j] = p;
  }
  int adr = 0;
  for (int j = 0; j < NV; j++) {
    d->dof_Madr[j] = adr;
    for (int k = j; k >= 0; k = d->dof_parent[k]) adr++;
  }
  if (adr > NM) return -3;
  d->nM = adr;
  for (int b = 0; b < m->nbody; b++) {
    int r = b;
    while (r > 0 && m->body_parent[r] > 0) r = m->body_parent[r];
    d->body_root[b] = r;
  }
  for (int b = 0; b < m->nbody; b++) d->body_subtreemass[b] = m->body_mass[b];
  for (int b = m->nbody - 1; b > 0; b--) d->body_subtreemass[m->body_parent[b]] += d->body_subtreemass[b];
  for (int j = 0; j < NV; j++) d->qpos[j] = m->qpos0[j];
  return 0;
}

/* mj_resetData */
void so101o_reset(const So101Tables* m, OracleData* d) {
  for (int j = 0; j < NV; j++) {
    d->qpos[j] = m->qpos0[j];
    d->qvel[j] = d->qacc_warmstart[j] = d->ctrl[j] = d->qfrc_applied[j] = d->qacc[j] = 0;
  }
  d->time = 0;
  d->warning_bad = 0;
}

/* ---------------------------------------------------------------------------------------- */
/* position stage                                                                             */
/* ---------------------------------------------------------------------------------------- */
/* mj_kinematics */
static void kinematics(const So101Tables* m, OracleData* d) {
  d->xquat[0][0] = 1; d->xquat[0][1] = d->xquat[0][2] = d->xquat[0][3] = 0;
  memset(d->xpos[0], 0, sizeof d->xpos[0]);
  quat2Mat(d->xmat[0], d->xquat[0]);
  for (int i = 1; i < m->nbody; i++) {
    double xpos[3], xquat[4];
    int pid = m->body_parent[i];
    if (pid) {
      mulMatVec3(xpos, d->xmat[pid], m->body_pos[i]);
      for (int k = 0; k < 3; k++) xpos[k] += d->xpos[pid][k];
      mulQuat(xquat, d->xquat[pid], m->body_quat[i]);
    } else {
      memcpy(xpos, m->body_pos[i], sizeof xpos);
      memcpy(xquat, m->body_quat[i], sizeof xquat);
    }
    int j = m->body_jnt[i];
    if (j >= 0) {
      double qloc[4], vec[3];
      rotVecQuat(d->xaxis[j], m->jnt_axis[j], xquat);
      rotVecQuat(d->xanchor[j], m->jnt_pos[j], xquat);
      for (int k = 0; k < 3; k++) d->xanchor[j][k] += xpos[k];
      axisAngle2Quat(qloc, m->jnt_axis[j], d->qpos[j] - m->qpos0[j]);
      mulQuat(xquat, xquat, qloc);
      rotVecQuat(vec, m->jnt_pos[j], xquat);
      for (int k = 0; k < 3; k++) xpos[k] = d->xanchor[j][k] - vec[k];
    }
    normalize4(xquat);
    memcpy(d->xquat[i], xquat, sizeof xquat);
    memcpy(d->xpos[i], xpos, sizeof xpos);
    quat2Mat(d->xmat[i], xquat);
  }
  /* inertial frames (mj_local2Global) */
  for (int i = 1; i < m->nbody; i++) {
    double q[4];
    mulMatVec3(d->xipos[i], d->xmat[i], m->body_ipos[i]);
    for (int k = 0; k < 3; k++) d->xipos[i][k] += d->xpos[i][k];
    mulQuat(q, d->xquat[i], m->body_iquat[i]);
    quat2Mat(d->ximat[i], q);
  }
  /* the observation site */
  mulMatVec3(d->site_xpos, d->xmat[m->site_body], m->site_pos);
  for (int k = 0; k < 3; k++) d->site_xpos[k] += d->xpos[m->site_body][k];
}

/* mj_comPos */
static void comPos(const So101Tables* m, OracleData* d) {
  memset(d->subtree_com, 0, sizeof d->subtree_com);
  for (int i = m->nbody - 1; i >= 0; i--) {
    for (int k = 0; k < 3; k++) d->subtree_com[i][k] += d->xipos[i][k] * m->body_mass[i];
    if (i) {
      int p = m->body_parent[i];
      for (int k = 0; k < 3; k++) d->subtree_com[p][k] += d->subtree_com[i][k];
    }
    if (d->body_subtreemass[i] < mjMINVAL) {
      memcpy(d->subtree_com[i], d->xipos[i], sizeof d->xipos[i]);
    } else {
      double s = 1.0 / fmax(mjMINVAL, d->body_subtreemass[i]);
      for (int k = 0; k < 3; k++) d->subtree_com[i][k] *= s;
    }
  }
  for (int i = 1; i < m->nbody; i++) {
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = d->xipos[i][k] - d->subtree_com[d->body_root[i]][k];
    inertCom(d->cinert[i], m->body_inertia[i], d->ximat[i], off, m->body_mass[i]);
  }
  for (int j = 0; j < NV; j++) {
    double off[3];
    int bi = m->jnt_body[j];
    for (int k = 0; k < 3; k++) off[k] = d->subtree_com[d->body_root[bi]][k] - d->xanchor[j][k];
    /* mju_dofCom */
    memcpy(d->cdof[j], d->xaxis[j], 3 * sizeof(double));
    cross3(d->cdof[j] + 3, d->xaxis[j], off);
  }
}

/* mj_crb */
static void crb(const So101Tables* m, OracleData* d) {
  memcpy(d->crb, d->cinert, sizeof d->crb);
  for (int i = m->nbody - 1; i > 0; i--) {
    int p = m->body_parent[i];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[p][k] += d->crb[i][k];
  }
  memset(d->qM, 0, sizeof d->qM);
  for (int i = 0; i < NV; i++) {
    double buf[6];
    int adr = d->dof_Madr[i];
    d->qM[adr] = m->dof_armature[i];
    mulInertVec(buf, d->crb[d->dof_body[i]], d->cdof[i]);
    for (int j = i; j >= 0; j = d->dof_parent[j]) d->qM[adr++] += dotn(d->cdof[j], buf, 6);
  }
}

/* mj_factorI: in-place L'*D*L of a matrix in qM layout */
static void factorI(const OracleData* d, const double* M, double* qLD, double* qLDiagInv) {
  memcpy(qLD, M, NM * sizeof(double));
  for (int k = NV - 1; k >= 0; k--) {
    int Madr_kk = d->dof_Madr[k];
    if (qLD[Madr_kk] < mjMINVAL) qLD[Madr_kk] = mjMINVAL;
    int Madr_ki = Madr_kk + 1;
    int i = d->dof_parent[k];
    while (i >= 0) {
      double tmp = qLD[Madr_ki] / qLD[Madr_kk];
      int cnt = 0;
      for (int a = i; a >= 0; a = d->dof_parent[a]) cnt++;
      for (int c = 0; c < cnt; c++) qLD[d->dof_Madr[i] + c] -= qLD[Madr_ki + c] * tmp;
      qLD[Madr_ki] = tmp;
      i = d->dof_parent[i];
      Madr_ki++;
    }
  }
  for (int i = 0; i < NV; i++) qLDiagInv[i] = 1.0 / qLD[d->dof_Madr[i]];
}

/* mj_solveLD */
static void solveLD(const OracleData* d, double* x, const double* qLD, const double* qLDiagInv) {
  for (int i = NV - 1; i >= 0; i--) {
    double tmp = x[i];
    if (tmp != 0) {
      int adr = d->dof_Madr[i] + 1;
      for (int j = d->dof_parent[i]; j >= 0; j = d->dof_parent[j]) x[j] -= qLD[adr++] * tmp;
    }
  }
  for (int i = 0; i < NV; i++) x[i] *= qLDiagInv[i];
  for (int i = 0; i < NV; i++) {
    int adr = d->dof_Madr[i] + 1;
    double tmp = x[i];
    for (int j = d->dof_parent[i]; j >= 0; j = d->dof_parent[j]) tmp -= qLD[adr++] * x[j];
    x[i] = tmp;
  }
}

/* mj_mulM */
static void mulM(const OracleData* d, double* res, const double* vec) {
  for (int i = 0; i < NV; i++) res[i] = 0;
  for (int i = 0; i < NV; i++) {
    int adr = d->dof_Madr[i];
    res[i] += d->qM[adr] * vec[i];
    adr++;
    for (int j = d->dof_parent[i]; j >= 0; j = d->dof_parent[j]) {
      res[i] += d->qM[adr] * vec[j];
      res[j] += d->qM[adr] * vec[i];
      adr++;
    }
  }
}

/* mj_fullM */
static void fullM(const OracleData* d, double* dst /*NV*NV*/, const double* M) {
  memset(dst, 0, NV * NV * sizeof(double));
  for (int i = 0; i < NV; i++) {
    int adr = d->dof_Madr[i];
    for (int j = i; j >= 0; j = d->dof_parent[j]) {
      dst[i * NV + j] = M[adr];
      dst[j * NV + i] = M[adr];
      adr++;
    }
  }
}

/* ---------------------------------------------------------------------------------------- */
/* constraints                                                                                */
/* ---------------------------------------------------------------------------------------- */
/* getimpedance (engine_core_constraint.c) */
static double impedance(const double* solimp, double pos, double margin) {
  if (solimp[0] == solimp[1] || solimp[2] <= mjMINVAL) return 0.5 * (solimp[0] + solimp[1]);
  double x = (pos - margin) / solimp[2];
  if (x < 0) x = -x;
  if (x >= 1 || x <= 0) return x >= 1 ? solimp[1] : solimp[0];
  double y;
  if (solimp[4] == 1) y = x;
  else if (x <= solimp[3]) {
    double a = 1 / pow(solimp[3], solimp[4] - 1);
    y = a * pow(x, solimp[4]);
  } else {
    double b = 1 / pow(1 - solimp[3], solimp[4] - 1);
    y = 1 - b * pow(1 - x, solimp[4]);
  }
  return solimp[0] + y * (solimp[1] - solimp[0]);
}

/* ---- convex hulls of the colliding geoms (test infrastructure: set once, read by every thread) ---- */
static struct {
  int ngeom, nvert;
  int* vert_start;
  double* vert;
} g_hulls = {0, 0, NULL, NULL};
void so101o_set_hulls(const So101Hulls* h) {
  free(g_hulls.vert_start); free(g_hulls.vert);
  g_hulls.ngeom = g_hulls.nvert = 0; g_hulls.vert_start = NULL; g_hulls.vert = NULL;
  if (!h || h->ngeom <= 0) return;
  g_hulls.ngeom = h->ngeom; g_hulls.nvert = h->nvert;
  g_hulls.vert_start = (int*)malloc(sizeof(int) * (h->ngeom + 1));
  g_hulls.vert = (double*)malloc(sizeof(double) * 3 * h->nvert);
  memcpy(g_hulls.vert_start, h->vert_start, sizeof(int) * (h->ngeom + 1));
  memcpy(g_hulls.vert, h->vert, sizeof(double) * 3 * h->nvert);
}

/* mj_collision, restricted to (table box top face, mesh hull) pairs: mjc_Convex gives ONE contact per pair, at the
   deepest hull vertex (brute force over the hull's vertices; the first lowest one wins), normal +z (from the box,
   geom1, into the mesh, geom2), dist = z_min - z_top, pos midway between the witness points.  con_gap = how much
   higher the runner-up vertex is (ties make the witness point ambiguous; tests skip them). */
static void collision(const So101Tables* m, OracleData* d) {
  d->ncon = 0; d->con_unsupported = 0;
  if (!m->con_enabled || g_hulls.ngeom != m->ntrip) return;
  for (int g = 0; g < m->ntrip && d->ncon < MAXCON; g++) {
    int b = m->trip_body[g];
    const double* R = d->xmat[b];
    {   /* broadphase: the geom's bounding box (box contains hull contains mesh) clears the plane -> no contact */
      double c[3], ext = 0;
      mulMatVec3(c, R, m->trip_center[g]);
      for (int ax = 0; ax < 3; ax++) {
        const double* u = m->trip_axes[g] + 3 * ax;
        ext += fabs(R[6] * u[0] + R[7] * u[1] + R[8] * u[2]) * m->trip_half[g][ax];
      }
      if (d->xpos[b][2] + c[2] - ext >= m->trip_plane_z + m->con_margin) continue;
    }
    /* witness vertex = support vertex of the direction "down" (world -z in the body frame).  Extruded shapes have
       edges that stay parallel to the table whatever the joint angles (hinge axes 1-3 are horizontal): a whole edge
       is lowest and the witness point is ambiguous (MuJoCo's answer there depends on its GJK/EPA iterates).  Rule
       here: the direction is tilted by 1e-7 towards e = (2,3,4)/sqrt(29) in the body frame, which orders tied
       vertices deterministically and moves the selected point by < 1e-9 m in height. */
    static const double tilt[3] = {2e-7 / 5.385164807134504, 3e-7 / 5.385164807134504, 4e-7 / 5.385164807134504};
    const double dsel[3] = {-R[6] + tilt[0], -R[7] + tilt[1], -R[8] + tilt[2]};
    double sbest = -1e300, s2 = -1e300;
    int arg = -1;
    for (int i = g_hulls.vert_start[g]; i < g_hulls.vert_start[g + 1]; i++) {
      const double* v = g_hulls.vert + 3 * i;
      double sv = dsel[0] * v[0] + dsel[1] * v[1] + dsel[2] * v[2];
      if (sv > sbest) { s2 = sbest; sbest = sv; arg = i; }
      else if (sv > s2) s2 = sv;
    }
    double zmin, z2 = 0;
    {
      const double* v = g_hulls.vert + 3 * arg;
      zmin = d->xpos[b][2] + (R[6] * v[0] + R[7] * v[1] + R[8] * v[2]);
      z2 = zmin + (sbest - s2);
    }
    double dist = zmin - m->trip_plane_z;
    if (!(dist < m->con_margin)) continue;
    int c = d->ncon++;
    const double* v = g_hulls.vert + 3 * arg;
    double p[3];
    mulMatVec3(p, R, v);
    for (int k = 0; k < 3; k++) p[k] += d->xpos[b][k];
    d->con_geom[c] = g; d->con_vert[c] = arg; d->con_dist[c] = dist; d->con_gap[c] = z2 - zmin;
    d->con_pos[c][0] = p[0]; d->con_pos[c][1] = p[1]; d->con_pos[c][2] = p[2] - 0.5 * dist;
    /* mju_makeFrame for the normal (0,0,1): y = (0,1,0), z = x cross y = (-1,0,0) */
    const double frame[9] = {0, 0, 1, 0, 1, 0, -1, 0, 0};
    memcpy(d->con_frame[c], frame, sizeof frame);
    if (p[0] < m->con_box[0] || p[0] > m->con_box[1] || p[1] < m->con_box[2] || p[1] > m->con_box[3])
      d->con_unsupported = 1;   /* beyond the top face: an edge contact of the box, not restated */
  }
}

static void addRow(OracleData* d, int type, int id, int dof, double jac, double pos, double margin,
                   double floss, double diagApprox) {
  int r = d->nefc++;
  memset(d->efc_J[r], 0, sizeof d->efc_J[r]);
  d->efc_J[r][dof] = jac;
  d->efc_type[r] = type; d->efc_id[r] = id;
  d->efc_pos[r] = pos; d->efc_margin[r] = margin; d->efc_frictionloss[r] = floss;
  d->efc_diagApprox[r] = diagApprox;
}

/* mj_makeConstraint (mj_instantiateFriction, mj_instantiateLimit) + mj_makeImpedance */
static void makeConstraint(const So101Tables* m, OracleData* d) {
  d->nefc = 0;
  for (int i = 0; i < NV; i++)
    if (m->dof_frictionloss[i] > 0)
      addRow(d, ROW_FRICTION, i, i, 1.0, 0.0, 0.0, m->dof_frictionloss[i], m->dof_invweight0[i]);
  d->nf = d->nefc;
  for (int i = 0; i < NV; i++) {
    if (!m->jnt_limited[i]) continue;
    double margin = m->jnt_margin[i], value = d->qpos[i];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[i][(side + 1) / 2] - value);
      if (dist < margin) addRow(d, ROW_LIMIT, i, i, -side, dist, margin, 0.0, m->dof_invweight0[i]);
    }
  }
  /* mj_instantiateContact: condim 3, pyramidal cone.  Jacobian of the contact point on the mesh's body (the table is
     the world: zero Jacobian) by mj_jac, rotated into the contact frame; rows J_n + mu J_tk and J_n - mu J_tk. */
  for (int c = 0; c < d->ncon && d->nefc + 4 <= MAXEFC; c++) {
    int b = m->trip_body[d->con_geom[c]];
    double jacp[3][NV], off[3];
    memset(jacp, 0, sizeof jacp);
    for (int k = 0; k < 3; k++) off[k] = d->con_pos[c][k] - d->subtree_com[d->body_root[b]][k];
    for (int a = b; a > 0; a = m->body_parent[a]) {
      int j = m->body_jnt[a];
      if (j < 0) continue;
      double tmp[3];
      cross3(tmp, d->cdof[j], off);                        /* cdof_ang x (point - com) + cdof_lin */
      for (int k = 0; k < 3; k++) jacp[k][j] = tmp[k] + d->cdof[j][3 + k];
    }
    double Jc[3][NV];
    for (int r3 = 0; r3 < 3; r3++)
      for (int j = 0; j < NV; j++)
        Jc[r3][j] = d->con_frame[c][3 * r3] * jacp[0][j] + d->con_frame[c][3 * r3 + 1] * jacp[1][j] +
                    d->con_frame[c][3 * r3 + 2] * jacp[2][j];
    double tran = m->body_invweight0[0][0] + m->body_invweight0[b][0];
    for (int k = 1; k < 3; k++) {
      double fri = m->con_friction[0];   /* mjContact.friction = (sliding, sliding, torsional, rolling, rolling) */
      for (int sgn = 1; sgn >= -1; sgn -= 2) {
        int r = d->nefc++;
        for (int j = 0; j < NV; j++) d->efc_J[r][j] = Jc[0][j] + sgn * fri * Jc[k][j];
        d->efc_type[r] = ROW_CONTACT; d->efc_id[r] = c;
        d->efc_pos[r] = d->con_dist[c]; d->efc_margin[r] = m->con_margin; d->efc_frictionloss[r] = 0;
        d->efc_diagApprox[r] = tran + fri * fri * tran;    /* mj_diagApprox, pyramidal, translational dims */
      }
    }
  }
  for (int r = 0; r < d->nefc; r++) {
    int id = d->efc_id[r];
    const double* solref = d->efc_type[r] == ROW_FRICTION ? m->dof_solref[id] : d->efc_type[r] == ROW_LIMIT ? m->jnt_solref[id] : m->con_solref;
    const double* solimp = d->efc_type[r] == ROW_FRICTION ? m->dof_solimp[id] : d->efc_type[r] == ROW_LIMIT ? m->jnt_solimp[id] : m->con_solimp;
    double sr0 = solref[0], sr1 = solref[1];
    if (sr0 > 0 && sr0 < 2 * m->timestep) sr0 = 2 * m->timestep; /* refsafe */
    double imp = impedance(solimp, d->efc_pos[r], d->efc_margin[r]);
    d->efc_R[r] = fmax(mjMINVAL, (1 - imp) * d->efc_diagApprox[r] / imp);
    d->efc_D[r] = 1 / d->efc_R[r];
    double K, B, dmax = solimp[1];
    if (sr0 > 0) {
      K = 1 / fmax(mjMINVAL, dmax * dmax * sr0 * sr0 * sr1 * sr1);
      B = 2 / fmax(mjMINVAL, dmax * sr0);
    } else {
      K = -sr0 / fmax(mjMINVAL, dmax * dmax);
      B = -sr1 / fmax(mjMINVAL, dmax);
    }
    if (d->efc_type[r] == ROW_FRICTION) K = 0;
    d->efc_KBIP[r][0] = K; d->efc_KBIP[r][1] = B; d->efc_KBIP[r][2] = imp; d->efc_KBIP[r][3] = 0;
  }
  /* mj_makeImpedance, frictional contacts: every row of a pyramid gets R = 2 mu^2 R[first row] (mu = friction[0] /
     sqrt(impratio), impratio 1) */
  for (int r = 0; r < d->nefc; r++) {
    if (d->efc_type[r] != ROW_CONTACT) continue;
    double mu = m->con_friction[0];
    double Rpy = 2 * mu * mu * d->efc_R[r];
    for (int j = 0; j < 4; j++) { d->efc_R[r + j] = Rpy; d->efc_D[r + j] = 1 / Rpy; }
    r += 3;
  }
}

/* mj_referenceConstraint */
static void referenceConstraint(OracleData* d) {
  for (int r = 0; r < d->nefc; r++) {
    d->efc_vel[r] = dotn(d->efc_J[r], d->qvel, NV);
    d->efc_aref[r] = -d->efc_KBIP[r][1] * d->efc_vel[r] -
                     d->efc_KBIP[r][0] * d->efc_KBIP[r][2] * (d->efc_pos[r] - d->efc_margin[r]);
  }
}

/* mj_constraintUpdate: forces, states, cost from jar = J*qacc - aref */
static void constraintUpdate(OracleData* d, const double* jar, double* cost, int flg_force) {
  double s = 0;
  for (int i = 0; i < d->nefc; i++) {
    d->efc_force[i] = -d->efc_D[i] * jar[i];
    if (i < d->nf) {
      double f = d->efc_frictionloss[i], R = d->efc_R[i];
      if (jar[i] <= -R * f) {
        s += -0.5 * R * f * f - f * jar[i];
        d->efc_force[i] = f;
        d->efc_state[i] = ST_LINEARNEG;
      } else if (jar[i] >= R * f) {
        s += -0.5 * R * f * f + f * jar[i];
        d->efc_force[i] = -f;
        d->efc_state[i] = ST_LINEARPOS;
      } else {
        s += 0.5 * d->efc_D[i] * jar[i] * jar[i];
        d->efc_state[i] = ST_QUADRATIC;
      }
      continue;
    }
    if (jar[i] >= 0) {
      d->efc_force[i] = 0;
      d->efc_state[i] = ST_SATISFIED;
    } else {
      s += 0.5 * d->efc_D[i] * jar[i] * jar[i];
      d->efc_state[i] = ST_QUADRATIC;
    }
  }
  if (cost) *cost = s;
  if (flg_force) {
    for (int j = 0; j < NV; j++) d->qfrc_constraint[j] = 0;
    for (int i = 0; i < d->nefc; i++)
      for (int j = 0; j < NV; j++) d->qfrc_constraint[j] += d->efc_J[i][j] * d->efc_force[i];
  }
}

/* ---------------------------------------------------------------------------------------- */
/* Newton solver (engine_solver.c: mj_solPrimal with flg_Newton, pyramidal => no cone Hessian)*/
/* ---------------------------------------------------------------------------------------- */
typedef struct {
  double Jaref[MAXEFC], Jv[MAXEFC], Ma[NV], Mv[NV], grad[NV], Mgrad[NV], search[NV];
  double quad[MAXEFC][3], quadGauss[3], H[NV * NV], cost;
  int LSiter;
} Primal;
typedef struct { double alpha, cost, deriv[2]; } Pnt;

/* mju_cholFactor / mju_cholSolve (dense, lower triangle) */
static void cholFactor(double* mat, int n) {
  for (int j = 0; j < n; j++) {
    double tmp = mat[j * (n + 1)];
    if (j) tmp -= dotn(mat + j * n, mat + j * n, j);
    if (tmp < mjMINVAL) tmp = mjMINVAL;
    mat[j * (n + 1)] = sqrt(tmp);
    tmp = 1 / mat[j * (n + 1)];
    for (int i = j + 1; i < n; i++) mat[i * n + j] = (mat[i * n + j] - dotn(mat + i * n, mat + j * n, j)) * tmp;
  }
}
static void cholSolve(double* res, const double* mat, const double* vec, int n) {
  for (int i = 0; i < n; i++) res[i] = vec[i];
  for (int i = 0; i < n; i++) {
    if (i) res[i] -= dotn(mat + i * n, res, i);
    res[i] /= mat[i * (n + 1)];
  }
  for (int i = n - 1; i >= 0; i--) {
    if (i < n - 1)
      for (int j = i + 1; j < n; j++) res[i] -= mat[j * n + i] * res[j];
    res[i] /= mat[i * (n + 1)];
  }
}

/* MakeHessian: H = M + J' * diag(D[quadratic]) * J, factorised.  MuJoCo updates the factor
   incrementally when rows change state (HessianIncremental); a full refactorisation gives the
   same matrix up to rounding. */
static void makeHessian(const OracleData* d, Primal* c) {
  fullM(d, c->H, d->qM);
  for (int r = 0; r < d->nefc; r++) {
    if (d->efc_state[r] != ST_QUADRATIC) continue;
    for (int i = 0; i < NV; i++) {
      if (d->efc_J[r][i] == 0) continue;
      for (int j = 0; j < NV; j++) c->H[i * NV + j] += d->efc_D[r] * d->efc_J[r][i] * d->efc_J[r][j];
    }
  }
  cholFactor(c->H, NV);
}

/* PrimalUpdateConstraint */
static void primalUpdateConstraint(OracleData* d, Primal* c) {
  constraintUpdate(d, c->Jaref, &c->cost, 1);
  double Gauss = 0;
  for (int i = 0; i < NV; i++) Gauss += 0.5 * (c->Ma[i] - d->qfrc_smooth[i]) * (d->qacc[i] - d->qacc_smooth[i]);
  c->quadGauss[0] = Gauss;
  c->cost += Gauss;
}
/* PrimalUpdateGradient (Newton) */
static void primalUpdateGradient(OracleData* d, Primal* c) {
  for (int i = 0; i < NV; i++) c->grad[i] = c->Ma[i] - d->qfrc_smooth[i] - d->qfrc_constraint[i];
  cholSolve(c->Mgrad, c->H, c->grad, NV);
}
/* PrimalPrepare */
static void primalPrepare(const OracleData* d, Primal* c) {
  c->quadGauss[1] = dotn(c->search, c->Ma, NV) - dotn(d->qfrc_smooth, c->search, NV);
  c->quadGauss[2] = 0.5 * dotn(c->search, c->Mv, NV);
  for (int i = 0; i < d->nefc; i++) {
    double DJ0 = d->efc_D[i] * c->Jaref[i];
    c->quad[i][0] = 0.5 * c->Jaref[i] * DJ0;
    c->quad[i][1] = c->Jv[i] * DJ0;
    c->quad[i][2] = 0.5 * c->Jv[i] * d->efc_D[i] * c->Jv[i];
  }
}
/* PrimalEval */
static void primalEval(const OracleData* d, Primal* c, Pnt* p, double alpha) {
  double qt[3] = {c->quadGauss[0], c->quadGauss[1], c->quadGauss[2]};
  for (int i = 0; i < d->nefc; i++) {
    double x = c->Jaref[i] + alpha * c->Jv[i];
    if (i < d->nf) {
      double f = d->efc_frictionloss[i], Rf = d->efc_R[i] * f;
      if (-Rf < x && x < Rf) {
        qt[0] += c->quad[i][0]; qt[1] += c->quad[i][1]; qt[2] += c->quad[i][2];
      } else if (x <= -Rf) {
        qt[0] += f * (-0.5 * Rf - c->Jaref[i]); qt[1] += -f * c->Jv[i];
      } else {
        qt[0] += f * (-0.5 * Rf + c->Jaref[i]); qt[1] += f * c->Jv[i];
      }
      continue;
    }
    if (x < 0) { qt[0] += c->quad[i][0]; qt[1] += c->quad[i][1]; qt[2] += c->quad[i][2]; }
  }
  p->alpha = alpha;
  p->cost = alpha * alpha * qt[2] + alpha * qt[1] + qt[0];
  p->deriv[0] = 2 * alpha * qt[2] + qt[1];
  p->deriv[1] = 2 * qt[2];
  if (p->deriv[1] <= 0) p->deriv[1] = mjMINVAL;
  c->LSiter++;
}
/* updateBracket */
static int updateBracket(const OracleData* d, Primal* c, Pnt* p, const Pnt cand[3], Pnt* pnext) {
  int flag = 0;
  for (int i = 0; i < 3; i++) {
    if (p->deriv[0] < 0 && cand[i].deriv[0] < 0 && p->deriv[0] < cand[i].deriv[0]) { *p = cand[i]; flag = 1; }
    else if (p->deriv[0] > 0 && cand[i].deriv[0] > 0 && p->deriv[0] > cand[i].deriv[0]) { *p = cand[i]; flag = 2; }
  }
  if (flag) primalEval(d, c, pnext, p->alpha - p->deriv[0] / p->deriv[1]);
  return flag;
}
/* PrimalSearch */
static double primalSearch(const So101Tables* m, OracleData* d, Primal* c) {
  Pnt p0, p1, p2, pmid, p1next, p2next;
  c->LSiter = 0;
  double snorm = sqrt(dotn(c->search, c->search, NV));
  if (snorm < mjMINVAL) return 0;
  double scale = 1 / (m->meaninertia * (NV > 1 ? NV : 1));
  double gtol = m->tolerance * m->ls_tolerance * snorm / scale;
  mulM(d, c->Mv, c->search);
  for (int i = 0; i < d->nefc; i++) c->Jv[i] = dotn(d->efc_J[i], c->search, NV);
  primalPrepare(d, c);
  primalEval(d, c, &p0, 0);
  primalEval(d, c, &p1, p0.alpha - p0.deriv[0] / p0.deriv[1]);
  if (p0.cost < p1.cost) p1 = p0;
  if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
  int dir = p1.deriv[0] < 0 ? +1 : -1;
  p2 = p1;
  while (p1.deriv[0] * dir <= -gtol && c->LSiter < m->ls_iterations) {
    p2 = p1;
    primalEval(d, c, &p1, p1.alpha - p1.deriv[0] / p1.deriv[1]);
    if (fabs(p1.deriv[0]) < gtol) return p1.alpha;
  }
  if (c->LSiter >= m->ls_iterations) return p1.alpha;
  p2next = p1;
  primalEval(d, c, &p1next, p1.alpha - p1.deriv[0] / p1.deriv[1]);
  while (c->LSiter < m->ls_iterations) {
    primalEval(d, c, &pmid, 0.5 * (p1.alpha + p2.alpha));
    Pnt cand[3] = {p1next, p2next, pmid};
    double bestcost = 0;
    int bestind = -1;
    for (int i = 0; i < 3; i++)
      if (fabs(cand[i].deriv[0]) < gtol && (bestind == -1 || cand[i].cost < bestcost)) {
        bestcost = cand[i].cost;
        bestind = i;
      }
    if (bestind >= 0) return cand[bestind].alpha;
    int b1 = updateBracket(d, c, &p1, cand, &p1next);
    int b2 = updateBracket(d, c, &p2, cand, &p2next);
    if (!b1 && !b2) return pmid.alpha;
  }
  if (p1.cost <= p2.cost && p1.cost < p0.cost) return p1.alpha;
  if (p2.cost <= p1.cost && p2.cost < p0.cost) return p2.alpha;
  return 0;
}

/*
 * ALTERNATIVE UNDER TEST, not MuJoCo: exact minimiser of the line-search objective.  Along the search
 * direction f'(alpha) is continuous, piecewise linear and non-decreasing (Huber friction rows, one-sided
 * limit rows); PrimalSearch above converges to its root within gtol.  walkSearch finds the root directly,
 * one linear piece at a time (next zone breakpoint -> slope/curvature at the piece's midpoint -> stop when
 * the piece's own root lies inside it).  The CUDA kernels use this form; so101o_set_line_search(1) switches
 * the oracle to it so that the equivalence can be measured on the CPU (tests/test_oracle.py).
 */
static int g_line_search = 0;   /* 0: PrimalSearch (MuJoCo, default)   1: walkSearch */
void so101o_set_line_search(int mode) { g_line_search = mode; }

static void slopeAt(const OracleData* d, const Primal* c, double alpha, double* d0, double* d1) {
  double s0 = c->quadGauss[1] + 2 * c->quadGauss[2] * alpha, s1 = 2 * c->quadGauss[2];
  for (int i = 0; i < d->nefc; i++) {
    double x = c->Jaref[i] + alpha * c->Jv[i];
    if (i < d->nf) {
      double f = d->efc_frictionloss[i], Rf = d->efc_R[i] * f;
      if (-Rf < x && x < Rf) { s0 += d->efc_D[i] * x * c->Jv[i]; s1 += d->efc_D[i] * c->Jv[i] * c->Jv[i]; }
      else s0 += (x <= -Rf ? -f : f) * c->Jv[i];
    } else if (x < 0) {
      s0 += d->efc_D[i] * x * c->Jv[i];
      s1 += d->efc_D[i] * c->Jv[i] * c->Jv[i];
    }
  }
  *d0 = s0;
  *d1 = s1 > 0 ? s1 : mjMINVAL;
}

static double walkSearch(OracleData* d, Primal* c) {
  c->LSiter = 0;
  if (sqrt(dotn(c->search, c->search, NV)) < mjMINVAL) return 0;
  mulM(d, c->Mv, c->search);
  for (int i = 0; i < d->nefc; i++) c->Jv[i] = dotn(d->efc_J[i], c->search, NV);
  primalPrepare(d, c);
  double alpha = 0;
  for (int it = 0; it < 2 * MAXEFC + 2; it++) {
    double nb = INFINITY;   /* next zone breakpoint beyond alpha */
    for (int i = 0; i < d->nefc; i++) {
      double jv = c->Jv[i];
      if (jv == 0) continue;
      if (i < d->nf) {
        double Rf = d->efc_R[i] * d->efc_frictionloss[i];
        double b1 = (Rf - c->Jaref[i]) / jv, b2 = (-Rf - c->Jaref[i]) / jv;
        if (b1 > alpha && b1 < nb) nb = b1;
        if (b2 > alpha && b2 < nb) nb = b2;
      } else {
        double b1 = -c->Jaref[i] / jv;
        if (b1 > alpha && b1 < nb) nb = b1;
      }
    }
    int last = !(nb < INFINITY);
    double mid = last ? alpha + 1 : 0.5 * (alpha + nb), d0, d1;
    slopeAt(d, c, mid, &d0, &d1);
    c->LSiter++;
    double root = mid - d0 / d1;
    if (root <= nb || last) return root > alpha ? root : alpha;
    alpha = nb;
  }
  return alpha;
}

/* mj_solPrimal(flg_Newton=1) */
static void solNewton(const So101Tables* m, OracleData* d) {
  Primal c;
  memset(&c, 0, sizeof c);
  mulM(d, c.Ma, d->qacc);
  for (int i = 0; i < d->nefc; i++) c.Jaref[i] = dotn(d->efc_J[i], d->qacc, NV) - d->efc_aref[i];
  primalUpdateConstraint(d, &c);
  makeHessian(d, &c);
  primalUpdateGradient(d, &c);
  for (int i = 0; i < NV; i++) c.search[i] = -c.Mgrad[i];
  double scale = 1 / (m->meaninertia * (NV > 1 ? NV : 1));
  int iter = 0, nls = 0;
  memset(d->ls_evals_iter, 0, sizeof d->ls_evals_iter);
  while (iter < m->iterations) {
    double alpha = g_line_search ? walkSearch(d, &c) : primalSearch(m, d, &c);
    nls += c.LSiter;
    if (iter < 8) d->ls_evals_iter[iter] = c.LSiter;
    if (alpha == 0) break;
    for (int i = 0; i < NV; i++) { d->qacc[i] += alpha * c.search[i]; c.Ma[i] += alpha * c.Mv[i]; }
    for (int i = 0; i < d->nefc; i++) c.Jaref[i] += alpha * c.Jv[i];
    double oldcost = c.cost;
    int oldstate[MAXEFC], changed = 0;
    memcpy(oldstate, d->efc_state, sizeof oldstate);
    primalUpdateConstraint(d, &c);
    for (int i = 0; i < d->nefc; i++)
      changed |= (oldstate[i] == ST_QUADRATIC) != (d->efc_state[i] == ST_QUADRATIC);
    if (changed) makeHessian(d, &c);
    primalUpdateGradient(d, &c);
    double improvement = scale * (oldcost - c.cost);
    double gradient = scale * sqrt(dotn(c.grad, c.grad, NV));
    iter++;
    if (improvement < m->tolerance || gradient < m->tolerance) break;
    for (int i = 0; i < NV; i++) c.search[i] = -c.Mgrad[i];
  }
  d->solver_niter = iter;
  d->solver_nls = nls;
  d->solver_cost = c.cost;
}

/* ---------------------------------------------------------------------------------------- */
/* pipeline                                                                                   */
/* ---------------------------------------------------------------------------------------- */
static void fwdPosition(const So101Tables* m, OracleData* d) {
  kinematics(m, d);
  comPos(m, d);
  crb(m, d);
  factorI(d, d->qM, d->qLD, d->qLDiagInv);
  collision(m, d);
  makeConstraint(m, d);
}

/* mj_comVel */
static void comVel(const So101Tables* m, OracleData* d) {
  memset(d->cvel[0], 0, sizeof d->cvel[0]);
  for (int i = 1; i < m->nbody; i++) {
    double cvel[6];
    memcpy(cvel, d->cvel[m->body_parent[i]], sizeof cvel);
    int j = m->body_jnt[i];
    if (j >= 0) {
      crossMotion(d->cdof_dot[j], cvel, d->cdof[j]);
      for (int k = 0; k < 6; k++) cvel[k] += d->cdof[j][k] * d->qvel[j];
    }
    memcpy(d->cvel[i], cvel, sizeof cvel);
  }
}

/* mj_rne with flg_acc = 0 */
static void rne(const So101Tables* m, OracleData* d) {
  double cacc[NB][6], cfrc[NB][6];
  memset(cacc, 0, sizeof cacc);
  memset(cfrc, 0, sizeof cfrc);
  for (int k = 0; k < 3; k++) cacc[0][3 + k] = -m->gravity[k];
  for (int i = 1; i < m->nbody; i++) {
    double tmp[6], tmp1[6];
    int j = m->body_jnt[i], p = m->body_parent[i];
    for (int k = 0; k < 6; k++) cacc[i][k] = cacc[p][k] + (j >= 0 ? d->cdof_dot[j][k] * d->qvel[j] : 0.0);
    mulInertVec(cfrc[i], d->cinert[i], cacc[i]);
    mulInertVec(tmp, d->cinert[i], d->cvel[i]);
    crossForce(tmp1, d->cvel[i], tmp);
    for (int k = 0; k < 6; k++) cfrc[i][k] += tmp1[k];
  }
  for (int i = m->nbody - 1; i > 0; i--) {
    int p = m->body_parent[i];
    if (p) for (int k = 0; k < 6; k++) cfrc[p][k] += cfrc[i][k];
  }
  for (int j = 0; j < NV; j++) d->qfrc_bias[j] = dotn(d->cdof[j], cfrc[d->dof_body[j]], 6);
}

static void fwdVelocity(const So101Tables* m, OracleData* d) {
  comVel(m, d);
  /* mj_passive: joint springs and dampers */
  for (int j = 0; j < NV; j++)
    d->qfrc_passive[j] = -m->jnt_stiffness[j] * (d->qpos[j] - m->qpos_spring[j]) - m->dof_damping[j] * d->qvel[j];
  referenceConstraint(d);
  rne(m, d);
}

/* mj_fwdActuation (stateless actuators, joint transmission) */
static void fwdActuation(const So101Tables* m, OracleData* d) {
  for (int j = 0; j < NV; j++) d->qfrc_actuator[j] = 0;
  for (int i = 0; i < m->nu; i++) {
    int j = m->act_dof[i];
    double gear = m->act_gear[i];
    double ctrl = d->ctrl[i];
    if (m->act_ctrllimited[i]) ctrl = fmax(m->act_ctrlrange[i][0], fmin(m->act_ctrlrange[i][1], ctrl));
    double length = gear * d->qpos[j], velocity = gear * d->qvel[j];
    double force = m->act_gain[i] * ctrl + m->act_bias[i][0] + m->act_bias[i][1] * length + m->act_bias[i][2] * velocity;
    if (m->act_forcelimited[i]) force = fmax(m->act_forcerange[i][0], fmin(m->act_forcerange[i][1], force));
    d->actuator_force[i] = force;
    d->qfrc_actuator[j] += gear * force;
  }
}

/* mj_fwdAcceleration */
static void fwdAcceleration(OracleData* d) {
  for (int j = 0; j < NV; j++) {
    d->qfrc_smooth[j] = d->qfrc_passive[j] - d->qfrc_bias[j] + d->qfrc_applied[j] + d->qfrc_actuator[j];
    d->qacc_smooth[j] = d->qfrc_smooth[j];
  }
  solveLD(d, d->qacc_smooth, d->qLD, d->qLDiagInv);
}

/*
 * ALTERNATIVE UNDER TEST, not MuJoCo: starting point of the Newton solver.  MuJoCo starts from the cheaper of
 * qacc_warmstart and qacc_smooth and needs 3 iterations in ~43 % of the steps of this scene.  The CUDA kernels
 * start from the per-dof closed-form minimiser of the diagonal problem
 *     0.5 * M_ii (a - qacc_smooth_i)^2 + huber_i(a - aref_i)          (friction rows only),
 * which identifies the active set of the coupled problem in > 99.5 % of the steps (M is armature dominated), so
 * that the first Newton step is already exact.  The cost is strictly convex, the minimiser unique: same result up
 * to rounding.  so101o_set_solver_start(1) switches the oracle to it (tests/test_oracle.py).
 */
static int g_solver_start = 0;   /* 0: MuJoCo's warm start (default)   1: prox start   2: direct active set, prox fallback */
void so101o_set_solver_start(int mode) { g_solver_start = mode; }

static void proxStart(OracleData* d) {
  double Mf[NV * NV];
  fullM(d, Mf, d->qM);
  memcpy(d->qacc, d->qacc_smooth, sizeof d->qacc);
  for (int i = 0; i < d->nf; i++) {
    int dof = d->efc_id[i];
    double Mii = Mf[dof * NV + dof], as = d->qacc_smooth[dof], ar = d->efc_aref[i];
    double f = d->efc_frictionloss[i], Rf = d->efc_R[i] * f, D = d->efc_D[i];
    double aq = (Mii * as + D * ar) / (Mii + D);          /* stationary point of the quadratic zone */
    double ap = as - f / Mii, an = as + f / Mii;          /* ... of the two linear zones */
    double a;
    if (fabs(aq - ar) < Rf) a = aq;
    else if (ap - ar >= Rf) a = ap;
    else if (an - ar <= -Rf) a = an;
    else a = ar + (as > ar ? Rf : -Rf);
    d->qacc[dof] = a;
  }
}

/* Alternative under test: the direct active-set solve of the CUDA kernels (csrc/so101_physics.cuh active_set_guess /
 * active_set_accept).  Zones of the friction rows guessed from the per-dof problem, ONE solve of
 * (M + J'DJ restricted to the quadratic rows) a = qfrc_smooth + ..., accepted iff every row lies in the zone it was
 * assumed in (KKT on a strictly convex piecewise-quadratic cost => global minimiser).  Only without limit rows;
 * returns 0 if not applicable / rejected.  so101o_set_solver_start(2) enables it (tests/test_oracle.py);
 * so101o_set_direct_retries(1) adds the kernels' second attempt (zones of the rejected candidate). */
static int g_direct_retries = 0;
static long g_direct_retry_count = 0;
void so101o_set_direct_retries(int n) { g_direct_retries = n; }
static int directActiveSet(OracleData* d) {
  if (d->nefc != d->nf) return 0;                        /* limit rows present */
  double Mf[NV * NV], H[NV * NV], rhs[NV], zone[MAXEFC];
  fullM(d, Mf, d->qM);
  memcpy(H, Mf, sizeof H);
  memcpy(rhs, d->qfrc_smooth, sizeof rhs);
  for (int i = 0; i < d->nf; i++) {
    int dof = d->efc_id[i];
    double Mii = Mf[dof * NV + dof], ar = d->efc_aref[i], f = d->efc_frictionloss[i], D = d->efc_D[i];
    double t = Mii * (d->qacc_smooth[dof] - ar);
    int quad = fabs(t) < d->efc_R[i] * f * Mii + f;
    zone[i] = quad ? 0.0 : (t >= 0 ? 1.0 : -1.0);
    if (quad) { H[dof * NV + dof] += D; rhs[dof] += D * ar; }
    else rhs[dof] -= zone[i] * f;
  }
  double a[NV];
  cholFactor(H, NV);
  cholSolve(a, H, rhs, NV);
  int ok = 1;
  for (int i = 0; i < d->nf; i++) {
    int dof = d->efc_id[i];
    double jar = a[dof] - d->efc_aref[i], Rf = d->efc_R[i] * d->efc_frictionloss[i];
    if (zone[i] == 0.0 ? !(fabs(jar) < Rf) : !(zone[i] * jar > Rf)) ok = 0;
  }
  for (int retry = 0; !ok && retry < g_direct_retries; retry++) {
    /* active-set iteration: zones of the rejected candidate, solve again */
    memcpy(H, Mf, sizeof H);
    memcpy(rhs, d->qfrc_smooth, sizeof rhs);
    for (int i = 0; i < d->nf; i++) {
      int dof = d->efc_id[i];
      double ar = d->efc_aref[i], f = d->efc_frictionloss[i], D = d->efc_D[i], jar = a[dof] - ar, Rf = d->efc_R[i] * f;
      zone[i] = fabs(jar) < Rf ? 0.0 : (jar > 0 ? 1.0 : -1.0);
      if (zone[i] == 0.0) { H[dof * NV + dof] += D; rhs[dof] += D * ar; }
      else rhs[dof] -= zone[i] * f;
    }
    cholFactor(H, NV);
    cholSolve(a, H, rhs, NV);
    ok = 1;
    for (int i = 0; i < d->nf; i++) {
      int dof = d->efc_id[i];
      double jar = a[dof] - d->efc_aref[i], Rf = d->efc_R[i] * d->efc_frictionloss[i];
      if (zone[i] == 0.0 ? !(fabs(jar) < Rf) : !(zone[i] * jar > Rf)) ok = 0;
    }
    g_direct_retry_count++;
  }
  if (!ok) return 0;
  memcpy(d->qacc, a, sizeof a);
  double jarv[MAXEFC], cost;
  for (int i = 0; i < d->nefc; i++) jarv[i] = dotn(d->efc_J[i], d->qacc, NV) - d->efc_aref[i];
  constraintUpdate(d, jarv, &cost, 1);                   /* efc_force, qfrc_constraint at the solution */
  d->solver_niter = 1;
  d->solver_nls = 0;
  return 1;
}

/* mj_fwdConstraint (warmstart + Newton) */
static void fwdConstraint(const So101Tables* m, OracleData* d) {
  if (!d->nefc) {
    memcpy(d->qacc, d->qacc_smooth, sizeof d->qacc);
    memset(d->qfrc_constraint, 0, sizeof d->qfrc_constraint);
    d->solver_niter = d->solver_nls = 0;
    return;
  }
  for (int i = 0; i < d->nefc; i++) d->efc_b[i] = dotn(d->efc_J[i], d->qacc_smooth, NV) - d->efc_aref[i];
  /* warmstart(): the better of qacc_warmstart and qacc_smooth */
  double jar[MAXEFC], Ma[NV], cost_warm, cost_smooth;
  memcpy(d->qacc, d->qacc_warmstart, sizeof d->qacc);
  for (int i = 0; i < d->nefc; i++) jar[i] = dotn(d->efc_J[i], d->qacc_warmstart, NV) - d->efc_aref[i];
  constraintUpdate(d, jar, &cost_warm, 0);
  mulM(d, Ma, d->qacc_warmstart);
  for (int i = 0; i < NV; i++) cost_warm += 0.5 * (Ma[i] - d->qfrc_smooth[i]) * (d->qacc_warmstart[i] - d->qacc_smooth[i]);
  constraintUpdate(d, d->efc_b, &cost_smooth, 0);
  d->used_warmstart = 1;
  if (cost_warm > cost_smooth) { memcpy(d->qacc, d->qacc_smooth, sizeof d->qacc); d->used_warmstart = 0; }
  if (g_solver_start == 2 && !d->ncon) {
    if (directActiveSet(d)) return;
    proxStart(d);
  }
  if (g_solver_start == 1 && !d->ncon) proxStart(d);
  solNewton(m, d);
}

static int isBad(double x) { return isnan(x) || x > mjMAXVAL || x < -mjMAXVAL; }

/* mj_forward */
void so101o_forward(const So101Tables* m, OracleData* d) {
  fwdPosition(m, d);
  fwdVelocity(m, d);
  fwdActuation(m, d);
  fwdAcceleration(d);
  fwdConstraint(m, d);
}

/* mj_Euler (implicit in joint damping) + mj_advance */
static void euler(const So101Tables* m, OracleData* d) {
  double qacc[NV];
  int dof_damping = 0;
  for (int i = 0; i < NV; i++) if (m->dof_damping[i] > 0) { dof_damping = 1; break; }
  if (!dof_damping) {
    memcpy(qacc, d->qacc, sizeof qacc);
  } else {
    double MhB[NM], qH[NM], qHDiagInv[NV];
    memcpy(MhB, d->qM, sizeof MhB);
    for (int i = 0; i < NV; i++) MhB[d->dof_Madr[i]] += m->timestep * m->dof_damping[i];
    factorI(d, MhB, qH, qHDiagInv);
    for (int i = 0; i < NV; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    solveLD(d, qacc, qH, qHDiagInv);
  }
  for (int i = 0; i < NV; i++) d->qvel[i] += qacc[i] * m->timestep;
  for (int i = 0; i < NV; i++) d->qpos[i] += d->qvel[i] * m->timestep;
  d->time += m->timestep;
  memcpy(d->qacc_warmstart, d->qacc, sizeof d->qacc);
}

/* mj_step.  mj_checkPos/Vel/Acc reset the env in MuJoCo; here the env is flagged (and reset,
   as MuJoCo does) so batch tests can see it. */
void so101o_step(const So101Tables* m, OracleData* d) {
  int bad = 0;
  for (int i = 0; i < NV; i++) bad |= isBad(d->qpos[i]) | isBad(d->qvel[i]);
  if (bad) { so101o_reset(m, d); d->warning_bad = 1; }
  so101o_forward(m, d);
  bad = 0;
  for (int i = 0; i < NV; i++) bad |= isBad(d->qacc[i]);
  if (bad) { so101o_reset(m, d); d->warning_bad = 1; so101o_forward(m, d); }
  euler(m, d);
}

/* dense M for tests */
void so101o_fullM(const OracleData* d, double* dst) { fullM(d, dst, d->qM); }

/* ---------------------------------------------------------------------------------------- */
/* Philox4x32-10 counter RNG — the control/reset stream shared (by specification, not by code) */
/* with the CUDA rollout kernel: key = (seed_lo, seed_hi), counter = (env_lo, env_hi, step, tag) */
/* ---------------------------------------------------------------------------------------- */
static void philox4x32_10(uint32_t ctr[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * ctr[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * ctr[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ ctr[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ ctr[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
void so101o_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key[2], uint32_t out[4]) {
  memcpy(out, ctr_in, 4 * sizeof(uint32_t));
  philox4x32_10(out, key[0], key[1]);
}
/* 8 uniforms in [0,1) with 32 bits each for (seed, env, step, stream): dims 0..3 from block 0,
   4..7 from block 1 */
void so101o_uniform8(uint64_t seed, int64_t env, uint32_t step, uint32_t stream, double out[8]) {
  for (uint32_t blk = 0; blk < 2; blk++) {
    uint32_t c[4] = {(uint32_t)env, (uint32_t)((uint64_t)env >> 32), step, stream * 2u + blk};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    for (int k = 0; k < 4; k++) out[4 * blk + k] = (double)c[k] * (1.0 / 4294967296.0);
  }
}

enum { STREAM_RESET = 0, STREAM_CTRL = 1, STREAM_FREQ = 2, STREAM_AMP = 3, STREAM_PHASE = 4 };

/* control u_t of one env [REF SOARM101_DataCollection.py:57-74 (sin/chirp), :115,132 (random)] */
static void gen_ctrl(const So101CtrlSpec* s, int64_t env, int t, const double* freq, const double* amp,
                     const double* phase, double u[SO101_NU_ENV]) {
  if (s->kind == SO101_CTRL_RANDOM) {
    double r[8];
    so101o_uniform8(s->seed, env, (uint32_t)t, STREAM_CTRL, r);
    for (int k = 0; k < SO101_NU_ENV; k++) u[k] = (r[k] - 0.5) * 2 * s->amp;
  } else {
    for (int k = 0; k < SO101_NU_ENV; k++) {
      double f = freq[k];
      if (s->kind == SO101_CTRL_CHIRP) f = freq[k] + (s->freq_hi - s->freq_lo) * ((double)t / s->t_total);
      u[k] = amp[k] * sin(2 * M_PI * f * t + phase[k]);
    }
  }
}

/*
 * CPU restatement of SOARM101DataGenerator.generate_physics_based_data
 * [REF SOARM101_DataCollection.py:90-136] on top of SOARM101Env.reset/step
 * [REF SOARM101_Env.py:77-142]: rows[n][T+1][13] = [u_i | float32(ee_pos) | float32(qpos[0:5])].
 * u_tensor (nullable) = [T+1][5][n] for SO101_CTRL_TENSOR.  qpos0/qvel0 (nullable) [n][6]
 * override the random reset.  final_state (nullable) [n][18].  flags: SO101_ROLL_*.
 * Returns the number of physics steps executed.
 */
int64_t so101o_rollout(const So101Tables* m, const So101CtrlSpec* s, int64_t n, int T, int frame_skip,
                       const double* qpos0, const double* qvel0, double* rows, double* final_state,
                       uint32_t flags, int nthreads, int64_t* newton_iters) {
  int64_t iters_total = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static) reduction(+ : iters_total)
#endif
  for (int64_t e = 0; e < n; e++) {
    OracleData d;
    so101o_init(m, &d);
    so101o_reset(m, &d);
    int64_t env = s->env_offset + e;
    double freq[5] = {0}, amp[5] = {0}, phase[5] = {0}, r[8];
    if (s->kind == SO101_CTRL_SIN || s->kind == SO101_CTRL_CHIRP) {
      so101o_uniform8(s->seed, env, 0, STREAM_FREQ, r);
      for (int k = 0; k < 5; k++) freq[k] = s->freq_lo + (s->freq_hi - s->freq_lo) * r[k];
      so101o_uniform8(s->seed, env, 0, STREAM_AMP, r);
      for (int k = 0; k < 5; k++) amp[k] = -s->amp + 2 * s->amp * r[k];
      so101o_uniform8(s->seed, env, 0, STREAM_PHASE, r);
      for (int k = 0; k < 5; k++) phase[k] = 2 * M_PI * r[k];
    }
    if (qpos0) {
      for (int k = 0; k < NV; k++) { d.qpos[k] = qpos0[e * NV + k]; d.qvel[k] = qvel0 ? qvel0[e * NV + k] : 0.0; }
    } else {
      so101o_uniform8(s->seed, env, 0, STREAM_RESET, r);
      for (int k = 0; k < SO101_NU_ENV; k++) d.qpos[k] = s->reset_lo + (s->reset_hi - s->reset_lo) * r[k];
    }
    so101o_forward(m, &d);
    double u[SO101_NU_ENV];
    for (int t = 0; t <= T; t++) {
      if (t > 0) {
        for (int k = 0; k < SO101_NU_ENV; k++) d.ctrl[k] = u[k];
        if (flags & SO101_ROLL_GRAVCOMP_HOLD) { /* [REF Koopman_MPC.py:119] */
          so101o_forward(m, &d);
          memcpy(d.qfrc_applied, d.qfrc_bias, sizeof d.qfrc_applied);
        }
        for (int ss = 0; ss < frame_skip; ss++) {
          so101o_step(m, &d);
          iters_total += d.solver_niter;
        }
      }
      if (s->kind == SO101_CTRL_TENSOR) {
        const double* ut = (const double*)s->u;
        for (int k = 0; k < SO101_NU_ENV; k++) u[k] = ut[((int64_t)t * SO101_NU_ENV + k) * n + e];
      } else {
        gen_ctrl(s, env, t, freq, amp, phase, u);
      }
      if (rows) {
        double* row = rows + ((int64_t)e * (T + 1) + t) * SO101_ROW;
        for (int k = 0; k < SO101_NU_ENV; k++) row[k] = u[k];
        for (int k = 0; k < 3; k++) row[5 + k] = (double)(float)d.site_xpos[k];
        for (int k = 0; k < 5; k++) row[8 + k] = (double)(float)d.qpos[k];
      }
    }
    if (final_state) {
      for (int k = 0; k < NV; k++) {
        final_state[e * 18 + k] = d.qpos[k];
        final_state[e * 18 + 6 + k] = d.qvel[k];
        final_state[e * 18 + 12 + k] = d.qacc_warmstart[k];
      }
    }
  }
  if (newton_iters) *newton_iters = iters_total;
  return n * (int64_t)T * frame_skip;
}

/*
 * n independent envs advanced nsub physics steps from given states (teacher forcing):
 * state_in/out [n][18] = qpos, qvel, qacc_warmstart; ctrl [n][6]; qfrc_applied nullable [n][6];
 * obs (nullable) [n][8] unrounded doubles = site_xpos of the last forward, then qpos[0:5];
 * aux (nullable) [n][4] = Newton iterations, line-search evaluations, nefc, bad flag of the
 * last sub-step.
 */
void so101o_step_batch(const So101Tables* m, int64_t n, const double* state_in, const double* ctrl,
                       const double* qfrc_applied, int nsub, double* state_out, double* obs, double* aux,
                       int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
  for (int64_t e = 0; e < n; e++) {
    OracleData d;
    so101o_init(m, &d);
    so101o_reset(m, &d);
    for (int k = 0; k < NV; k++) {
      d.qpos[k] = state_in[e * 18 + k];
      d.qvel[k] = state_in[e * 18 + 6 + k];
      d.qacc_warmstart[k] = state_in[e * 18 + 12 + k];
      d.ctrl[k] = ctrl[e * NV + k];
      if (qfrc_applied) d.qfrc_applied[k] = qfrc_applied[e * NV + k];
    }
    if (nsub == 0) so101o_forward(m, &d);
    for (int ss = 0; ss < nsub; ss++) so101o_step(m, &d);
    for (int k = 0; k < NV; k++) {
      state_out[e * 18 + k] = d.qpos[k];
      state_out[e * 18 + 6 + k] = d.qvel[k];
      state_out[e * 18 + 12 + k] = nsub ? d.qacc_warmstart[k] : d.qacc[k];
    }
    if (obs) {
      for (int k = 0; k < 3; k++) obs[e * 8 + k] = d.site_xpos[k];
      for (int k = 0; k < 5; k++) obs[e * 8 + 3 + k] = d.qpos[k];
    }
    if (aux) {
      aux[e * 4 + 0] = d.solver_niter; aux[e * 4 + 1] = d.solver_nls;
      aux[e * 4 + 2] = d.nefc; aux[e * 4 + 3] = d.warning_bad;
    }
  }
}

/* contact situation of n states (qpos = state_in[e][0:6]): out[e] = {ncon, unsupported, smallest runner-up gap of
   the witness vertices (ties make MuJoCo's witness point ambiguous), deepest dist} */
void so101o_contact_probe(const So101Tables* m, int64_t n, const double* state_in, double* out, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
  for (int64_t e = 0; e < n; e++) {
    OracleData d;
    so101o_init(m, &d);
    so101o_reset(m, &d);
    for (int k = 0; k < NV; k++) d.qpos[k] = state_in[e * 18 + k];
    kinematics(m, &d);
    collision(m, &d);
    double gap = 1e300, dist = 0;
    for (int c = 0; c < d.ncon; c++) { gap = fmin(gap, d.con_gap[c]); dist = fmin(dist, d.con_dist[c]); }
    out[e * 4 + 0] = d.ncon; out[e * 4 + 1] = d.con_unsupported; out[e * 4 + 2] = gap; out[e * 4 + 3] = dist;
  }
}

/*
 * Shooting: B control sequences U[H][5][B] from one shared state0[18]; X[B][H+1][8] float32
 * observations (row 0 = observation of state0 after mj_forward).
 */
void so101o_shoot(const So101Tables* m, const double* state0, const double* U, int64_t B, int H,
                  int frame_skip, float* X, uint32_t flags, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(static)
#endif
  for (int64_t e = 0; e < B; e++) {
    OracleData d;
    so101o_init(m, &d);
    so101o_reset(m, &d);
    for (int k = 0; k < NV; k++) {
      d.qpos[k] = state0[k]; d.qvel[k] = state0[6 + k]; d.qacc_warmstart[k] = state0[12 + k];
    }
    so101o_forward(m, &d);
    for (int t = 0; t <= H; t++) {
      if (t > 0) {
        for (int k = 0; k < SO101_NU_ENV; k++) d.ctrl[k] = U[((int64_t)(t - 1) * SO101_NU_ENV + k) * B + e];
        if (flags & SO101_ROLL_GRAVCOMP_HOLD) {
          so101o_forward(m, &d);
          memcpy(d.qfrc_applied, d.qfrc_bias, sizeof d.qfrc_applied);
        }
        for (int ss = 0; ss < frame_skip; ss++) so101o_step(m, &d);
      }
      float* x = X + ((int64_t)e * (H + 1) + t) * 8;
      for (int k = 0; k < 3; k++) x[k] = (float)d.site_xpos[k];
      for (int k = 0; k < 5; k++) x[3 + k] = (float)d.qpos[k];
    }
  }
}

int so101o_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
