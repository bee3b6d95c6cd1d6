// so101_contact.cuh — table-plane contact: the path of an env whose collision bounding box dipped
// below the table top.  Reference: the scene enables contacts [REF SOARM101/SO101/scene_with_table_v.xml:28,31;
// so101_new_calib_v.xml:53-117]; mj_step then runs mj_collision (mesh hull vs box: mjc_Convex, one contact per geom
// pair), mj_instantiateContact (condim 3, pyramidal: rows J_n +- mu J_t), mj_makeImpedance and solves all rows with
// the Newton solver.  Restated for CPU in oracle/so101_oracle.c (collision, makeConstraint).
//
// This file is the RARE path (a few % of the envs of a long chirp rollout, none of a short random one):
//   1. contact_geometry: exact hull-vs-plane test of every hull whose bounding box tripped - support vertex in direction
//      -z by steepest ascent on the hull's edge graph, started from the previous step's vertex or a direction cube map -
//      and, only for hulls that do touch, world frames, contact points and the three Jacobian rows of each contact;
//   2. contact_active_set: the direct active-set iteration over friction (Huber), joint-limit and pyramid rows (inline);
//   3. contact_solve: the general dense Newton solve over the same rows with MuJoCo's warm-start pick, termination tests
//      and the exact piecewise-quadratic line search (out of line; only when 2 does not settle).
#pragma once

namespace so101 {

constexpr int MAXCON = SO101_MAXCON;
constexpr int MAXONE = NV + 4 * MAXCON;     // one-sided rows: joint limits + pyramid rows

// The active-set iteration (see contact_solve) is Newton's method with full steps on a piecewise-quadratic objective.  Where
// it does not settle it cycles between two pieces in which one friction row is saturated with opposite signs: each
// piece's minimiser lies beyond the row's (narrow) quadratic zone, the true minimiser inside it.  From the fourth solve on,
// a friction row whose candidate jumped from one saturated side to the other is therefore put into its quadratic zone.
// Measured on the oracle's contact steps of the benchmark workload (583 steps): plain iteration 2.6 % unsettled after
// 12 solves, with this rule none after 6 (mean 2.6 solves either way).
constexpr int ACTIVE_SET_ATTEMPTS = 10;
SO101_DEV void active_set_unflip(uint32_t zq, uint32_t zp, uint32_t& nzq, uint32_t& nzp) {
  const uint32_t flipped = ~zq & ~nzq & (zp ^ nzp) & ((1u << NV) - 1u);
  nzq |= flipped;
  nzp &= ~flipped;
}

template <typename T> SO101_DEV uint64_t hull_word_(const DevModel<T>& m, int which) {
  return ((uint64_t)(uint32_t)m.hull_ptr[2 * which + 1] << 32) | (uint64_t)(uint32_t)m.hull_ptr[2 * which];
}

// first vertex of a hull walk without a cached one: direction cube map (cold: once per hull and launch)
template <typename T>
__device__ __noinline__ int hull_cube_start(const DevModel<T>& m, int g, double d0, double d1, double d2) {
  const int32_t* cube = reinterpret_cast<const int32_t*>(hull_word_(m, 3));
  const int res = m.hull_res;
  const double d[3] = {d0, d1, d2};
  int ax = 0;
  if (fabs(d[1]) > fabs(d[ax])) ax = 1;
  if (fabs(d[2]) > fabs(d[ax])) ax = 2;
  const double dm = fabs(d[ax]);
  const double inv = dm > 0.0 ? 1.0 / dm : 0.0;
  const double u = d[(ax + 1) % 3] * inv, w = d[(ax + 2) % 3] * inv;
  int iu = (int)((u + 1.0) * 0.5 * res), iw = (int)((w + 1.0) * 0.5 * res);
  iu = iu < 0 ? 0 : (iu >= res ? res - 1 : iu);
  iw = iw < 0 ? 0 : (iw >= res ? res - 1 : iw);
  const int face = 2 * ax + (d[ax] < 0.0 ? 1 : 0);
  return cube[((g * 6 + face) * res + iu) * res + iw];
}

// support vertex of hull `g` for direction d (link frame): steepest ascent on the hull's edge graph from `start` (the
// support vertex of the previous step, if the caller kept it) or from a direction cube map.  The hulls are convex and the
// direction carries a generic tilt, so the ascent ends at the one global maximiser whatever the start.  Neighbours are
// fetched two at a time so that their loads are in flight together (the walk is a chain of dependent global loads); the
// code is kept small on purpose: in a team kernel every instruction line this rare path touches evicts one of the hot
// step's from the SM's instruction cache (measured: tools/team_timing.py, profiles/README.md round 2).
// The search runs in double whatever T is: the tie-breaking tilt (1e-7) is below float resolution, and without it a
// face lying flat on the table is a plateau on which the answer would depend on where the walk started.
template <typename T>
SO101_DEV int hull_support(const DevModel<T>& m, int g, const double (&d)[3], T (&v)[3], int start = -1) {
  const double* vert = reinterpret_cast<const double*>(hull_word_(m, 0));
  const int32_t* adj_start = reinterpret_cast<const int32_t*>(hull_word_(m, 1));
  const int32_t* adj = reinterpret_cast<const int32_t*>(hull_word_(m, 2));
  int cur = start;
  if (cur < 0) cur = hull_cube_start(m, g, d[0], d[1], d[2]);
  double best = dot3_(d[0], vert[3 * cur], d[1], vert[3 * cur + 1], d[2], vert[3 * cur + 2]);
#pragma unroll 1
  for (int it = 0; it < 4096; it++) {
    int nxt = -1;
    const int e0 = adj_start[cur], e1 = adj_start[cur + 1];
#pragma unroll 1
    for (int e = e0; e < e1; e += 2) {
      const int c0 = adj[e], c1 = adj[e + 1 < e1 ? e + 1 : e];
      const double v0 = dot3_(d[0], vert[3 * c0], d[1], vert[3 * c0 + 1], d[2], vert[3 * c0 + 2]);
      const double v1 = dot3_(d[0], vert[3 * c1], d[1], vert[3 * c1 + 1], d[2], vert[3 * c1 + 2]);
      if (v0 > best) { best = v0; nxt = c0; }
      if (v1 > best) { best = v1; nxt = c1; }
    }
    if (nxt < 0) break;
    cur = nxt;
  }
  v[0] = (T)vert[3 * cur]; v[1] = (T)vert[3 * cur + 1]; v[2] = (T)vert[3 * cur + 2];
  return cur;
}

template <typename T>
struct OneSided {        // one-sided rows (active when jar < 0): joint limits first, then 4 pyramid rows per contact
  int n;
  T J[MAXONE][NV], aref[MAXONE], D[MAXONE];
};

// cost, constraint force and Hessian activity of all rows at acceleration a (mj_constraintUpdate + Gauss term)
template <typename T>
SO101_DEV T contact_cost(const DevModel<T>& m, const T* aref_f, const OneSided<T>& os, const T* a, const T* Ma,
                         const T* fsm, const T* asm_, T* qc, T* hd, uint32_t& active) {
  T s = T(0);
#pragma unroll 1
  for (int i = 0; i < NV; i++) {
    const T jar = a[i] - aref_f[i];
    if (m.fr_f[i] == T(0)) { qc[i] = T(0); hd[i] = T(0); continue; }
    if (abs_(jar) >= m.fr_Rf[i]) {
      const T fs = copysign_(m.fr_f[i], jar);
      s += fs * jar - m.fr_hRff[i]; qc[i] = -fs; hd[i] = T(0);
    } else {
      const T Dj = m.fr_D[i] * jar;
      s += T(0.5) * Dj * jar; qc[i] = -Dj; hd[i] = m.fr_D[i];
    }
  }
  active = 0;
#pragma unroll 1
  for (int r = 0; r < os.n; r++) {
    T jar = -os.aref[r];
#pragma unroll 1
    for (int j = 0; j < NV; j++) jar += os.J[r][j] * a[j];
    if (jar < T(0)) {
      s += T(0.5) * os.D[r] * jar * jar;
      const T f = -os.D[r] * jar;
#pragma unroll 1
      for (int j = 0; j < NV; j++) qc[j] += os.J[r][j] * f;
      active |= 1u << r;
    }
  }
  T g = T(0);
#pragma unroll 1
  for (int i = 0; i < NV; i++) g += T(0.5) * (Ma[i] - fsm[i]) * (a[i] - asm_[i]);
  return s + g;
}

// exact line search over friction (Huber) and one-sided rows: the root of the piecewise-linear slope (see line_search)
template <typename T>
SO101_DEV T contact_line_search(const DevModel<T>& m, const T* aref_f, const OneSided<T>& os, const T* Mm, const T* a,
                                const T* Ma, const T* fsm, const T* sr, T* Mv, uint32_t& nev_total) {
  T ss = T(0);
#pragma unroll 1
  for (int i = 0; i < NV; i++) ss += sr[i] * sr[i];
  if (sqrt_(ss) < T(MJ_MINVAL)) return T(0);
#pragma unroll 1
  for (int i = 0; i < NV; i++) {
    T s = T(0);
#pragma unroll 1
    for (int j = 0; j < NV; j++) s += Mm[i >= j ? tri(i, j) : tri(j, i)] * sr[j];
    Mv[i] = s;
  }
  T G1 = T(0), G2 = T(0);
#pragma unroll 1
  for (int i = 0; i < NV; i++) { G1 += sr[i] * Ma[i] - fsm[i] * sr[i]; G2 += sr[i] * Mv[i]; }
  G2 *= T(0.5);
  T jar1[MAXONE], jv1[MAXONE];
#pragma unroll 1
  for (int r = 0; r < os.n; r++) {
    T x = -os.aref[r], v = T(0);
#pragma unroll 1
    for (int j = 0; j < NV; j++) { x += os.J[r][j] * a[j]; v += os.J[r][j] * sr[j]; }
    jar1[r] = x; jv1[r] = v;
  }
  T alpha = T(0), result = T(0);
  uint32_t nev = 0;
#pragma unroll 1
  for (int it = 0; it < 3 * NV + MAXONE + 2; it++) {
    T nb = inf_<T>();
#pragma unroll 1
    for (int i = 0; i < NV; i++) {
      if (sr[i] == T(0) || m.fr_f[i] == T(0)) continue;
      const T x0 = a[i] - aref_f[i];
      const T b0 = (-m.fr_Rf[i] - x0) / sr[i], b1 = (m.fr_Rf[i] - x0) / sr[i];
      if (b0 > alpha) nb = min_(nb, b0);
      if (b1 > alpha) nb = min_(nb, b1);
    }
#pragma unroll 1
    for (int r = 0; r < os.n; r++) {
      if (jv1[r] == T(0)) continue;
      const T b = -jar1[r] / jv1[r];
      if (b > alpha) nb = min_(nb, b);
    }
    const bool last = !(nb < inf_<T>());
    const T mid = last ? alpha + T(1) : T(0.5) * (alpha + nb);
    T d0 = G1 + T(2) * G2 * mid, d1 = T(2) * G2;
#pragma unroll 1
    for (int i = 0; i < NV; i++) {
      if (m.fr_f[i] == T(0)) continue;
      const T x = a[i] - aref_f[i] + mid * sr[i];
      const T Ds = m.fr_D[i] * sr[i];
      if (abs_(x) < m.fr_Rf[i]) { d0 += Ds * x; d1 += Ds * sr[i]; }
      else d0 += copysign_(m.fr_f[i], x) * sr[i];
    }
#pragma unroll 1
    for (int r = 0; r < os.n; r++) {
      const T x = jar1[r] + mid * jv1[r];
      if (x < T(0)) { d0 += os.D[r] * x * jv1[r]; d1 += os.D[r] * jv1[r] * jv1[r]; }
    }
    nev++;
    if (d1 <= T(0)) d1 = T(MJ_MINVAL);
    const T root = mid - d0 / d1;
    if (root <= nb || last) { result = max_(root, alpha); break; }
    alpha = nb;
    result = nb;
  }
  nev_total += nev;
  return result;
}

template <typename T>
struct Con3 {
  T Jn[NV], Jy[NV], Jx[NV];     // contact frame: normal +z, tangents (0,1,0) and (-1,0,0): Jx holds the x' = -x row
  T D, c0, vn, vy, vx;          // row weight, -K imp (dist - margin), J . qvel
};

// What the general solver reads and writes, copied by the caller: only these copies have their address taken, the hot
// path's own arrays stay in registers.
template <typename T>
struct ContactIO {
  T warm[NV], M[21], fsm[NV], aref_f[NV];
  T lim_side[NV], lim_aref[NV], lim_D[NV];
  uint32_t anylim;
  // out
  T a[NV], qc[NV];
  uint32_t flags, newton, lsevals;
};

// General solver over friction, limit and contact rows (the contacts as contact_geometry found them): the fallback of
// the inline active-set iteration.  io.a = qacc and io.qc = qfrc_constraint of the full problem.
template <typename T>
__device__ __noinline__ void contact_solve(const DevModel<T>& m, ContactIO<T>& io, const Con3<T>* con, int ncon) {
  const T* Mm = io.M; const T* fsm = io.fsm; const T* aref_f = io.aref_f;
  const T* lim_side = io.lim_side; const T* lim_aref = io.lim_aref; const T* lim_D = io.lim_D; const T* warm = io.warm;
  const bool anylim = io.anylim != 0;
  T* a = io.a; T* qc = io.qc;
  uint32_t& flags = io.flags;
  OneSided<T> os;
  os.n = 0;
  if (anylim) {
#pragma unroll 1
    for (int i = 0; i < NV; i++) {
      if (lim_side[i] == T(0)) continue;
      const int r = os.n++;
#pragma unroll 1
      for (int j = 0; j < NV; j++) os.J[r][j] = T(0);
      os.J[r][i] = lim_side[i];
      os.aref[r] = lim_aref[i];
      os.D[r] = lim_D[i];
    }
  }
  // contact frame: normal +z, tangents (0,1,0) and (-1,0,0) (mju_makeFrame); pyramid rows J_n +- mu J_t
#pragma unroll 1
  for (int c = 0; c < ncon; c++) {
#pragma unroll 1
    for (int e = 0; e < 4; e++) {
      const int r = os.n++;
      const T sg = (e & 1) ? -m.con_mu : m.con_mu;
#pragma unroll 1
      for (int j = 0; j < NV; j++) os.J[r][j] = con[c].Jn[j] + sg * (e < 2 ? con[c].Jy[j] : con[c].Jx[j]);
      os.D[r] = con[c].D;
      os.aref[r] = -m.con_B * (con[c].vn + sg * (e < 2 ? con[c].vy : con[c].vx)) + con[c].c0;
    }
  }

  // ---- direct active-set iteration ----------------------------------------------------------------------------------
  // The objective is strictly convex and piecewise quadratic.  On the piece where friction row i is in zone z_i
  // (0: quadratic, +-1: saturated) and the one-sided rows of the set A are active its minimiser solves
  //   (M + diag(D_i [z_i = 0]) + sum_A D_r J_r J_r') a = qfrc_smooth + D_i aref_i [z_i = 0] - z_i f_i + sum_A D_r aref_r J_r,
  // and a solution that lies in the piece it assumed is the global minimiser (KKT) - MuJoCo's converged Newton iterate
  // without cost evaluations or line searches.  Pieces: first the one qacc_warmstart lies in (a persisting contact keeps
  // its piece), then the one the rejected candidate lies in (a full Newton step), a few times; whatever is left takes
  // the safeguarded Newton iteration below.
  {
    T az[NV];
    uint32_t zq = 0, zp = 0, act = 0;      // friction zones (quadratic / saturated positive), active one-sided rows
#pragma unroll
    for (int i = 0; i < NV; i++) az[i] = warm[i];
    bool done = false;
#pragma unroll 1
    for (int attempt = 0; attempt <= ACTIVE_SET_ATTEMPTS && !done; attempt++) {
      uint32_t nzq = 0, nzp = 0, nact = 0;
      bool strict = true;
#pragma unroll
      for (int i = 0; i < NV; i++) {
        const T jar = az[i] - aref_f[i];
        if (m.fr_f[i] == T(0) || abs_(jar) < m.fr_Rf[i]) nzq |= 1u << i;
        else if (jar > T(0)) nzp |= 1u << i;
        strict &= m.fr_f[i] == T(0) || abs_(jar) != m.fr_Rf[i];
      }
#pragma unroll 1
      for (int r = 0; r < os.n; r++) {
        T jar = -os.aref[r];
#pragma unroll
        for (int j = 0; j < NV; j++) jar += os.J[r][j] * az[j];
        if (jar < T(0)) nact |= 1u << r;
      }
      if (attempt > 0 && nzq == zq && nzp == zp && nact == act && strict) {   // the candidate lies in its own piece
#pragma unroll
        for (int i = 0; i < NV; i++) {
          a[i] = az[i];
          const T jar = az[i] - aref_f[i];
          qc[i] = (zq >> i & 1u) ? -m.fr_D[i] * jar : ((zp >> i & 1u) ? -m.fr_f[i] : m.fr_f[i]);
          if (m.fr_f[i] == T(0)) qc[i] = T(0);
        }
#pragma unroll 1
        for (int r = 0; r < os.n; r++) {
          if (!(act >> r & 1u)) continue;
          T jar = -os.aref[r];
#pragma unroll
          for (int j = 0; j < NV; j++) jar += os.J[r][j] * az[j];
          const T f = -os.D[r] * jar;
#pragma unroll
          for (int j = 0; j < NV; j++) qc[j] += os.J[r][j] * f;
        }
        io.newton += attempt;
        done = true;
        break;
      }
      if (attempt == ACTIVE_SET_ATTEMPTS) break;
      if (attempt >= 3) active_set_unflip(zq, zp, nzq, nzp);
      zq = nzq; zp = nzp; act = nact;
      T H[21], rhs[NV];
#pragma unroll
      for (int i = 0; i < 21; i++) H[i] = Mm[i];
#pragma unroll
      for (int i = 0; i < NV; i++) {
        rhs[i] = fsm[i];
        if (m.fr_f[i] == T(0)) continue;
        if (zq >> i & 1u) { H[tri(i, i)] += m.fr_D[i]; rhs[i] += m.fr_D[i] * aref_f[i]; }
        else rhs[i] += (zp >> i & 1u) ? -m.fr_f[i] : m.fr_f[i];
      }
#pragma unroll 1
      for (int r = 0; r < os.n; r++) {
        if (!(act >> r & 1u)) continue;
        T Jr[NV];
#pragma unroll
        for (int j = 0; j < NV; j++) Jr[j] = os.J[r][j];
        const T Dr = os.D[r], Da = Dr * os.aref[r];
#pragma unroll
        for (int i = 0; i < NV; i++) {
          const T di = Dr * Jr[i];
          rhs[i] += Da * Jr[i];
#pragma unroll
          for (int j = 0; j <= i; j++) H[tri(i, j)] += di * Jr[j];
        }
      }
      ldl6_factor_solve(H, rhs);
#pragma unroll
      for (int i = 0; i < NV; i++) az[i] = rhs[i];
    }
    if (done) return;
  }

  // ---- Newton over all rows (mj_solPrimal, Newton flavour) ---------------------------------------------------------
  T asm_[NV], Ma[NV], hd[NV], cost;
  uint32_t active = 0;
  {
    T A[21];
#pragma unroll
    for (int i = 0; i < 21; i++) A[i] = Mm[i];
#pragma unroll
    for (int i = 0; i < NV; i++) asm_[i] = fsm[i];
    ldl6_factor_solve(A, asm_);
  }
  {  // warm start: the cheaper of qacc_warmstart and qacc_smooth
    T cbest = T(0);
#pragma unroll 1
    for (int c = 0; c < 2; c++) {
      T ca[NV], cMa[NV], cqc[NV], chd[NV];
      uint32_t cact;
#pragma unroll 1
      for (int i = 0; i < NV; i++) ca[i] = c ? asm_[i] : warm[i];
#pragma unroll 1
      for (int i = 0; i < NV; i++) {
        T s = T(0);
#pragma unroll 1
        for (int j = 0; j < NV; j++) s += Mm[i >= j ? tri(i, j) : tri(j, i)] * ca[j];
        cMa[i] = s;
      }
      const T cc = contact_cost(m, aref_f, os, ca, cMa, fsm, asm_, cqc, chd, cact);
      if (c == 0 || cbest > cc) {
#pragma unroll 1
        for (int i = 0; i < NV; i++) { a[i] = ca[i]; Ma[i] = cMa[i]; qc[i] = cqc[i]; hd[i] = chd[i]; }
        cost = cc; active = cact;
      }
      if (c == 0) cbest = cc;
    }
  }
  int iter = 0;
#pragma unroll 1
  while (iter < m.iterations) {
    T sr[NV], Mv[NV];
    {  // H = M + diag(friction quadratic) + sum_active D J'J ; search = -H^-1 grad
      T H[21];
#pragma unroll
      for (int i = 0; i < 21; i++) H[i] = Mm[i];
#pragma unroll
      for (int i = 0; i < NV; i++) { H[tri(i, i)] += hd[i]; sr[i] = Ma[i] - fsm[i] - qc[i]; }
#pragma unroll 1
      for (int r = 0; r < os.n; r++) {
        if (!(active >> r & 1u)) continue;
#pragma unroll 1
        for (int i = 0; i < NV; i++) {
          const T di = os.D[r] * os.J[r][i];
          if (di == T(0)) continue;
#pragma unroll 1
          for (int j = 0; j <= i; j++) H[tri(i, j)] += di * os.J[r][j];
        }
      }
      ldl6_factor_solve(H, sr);
#pragma unroll
      for (int i = 0; i < NV; i++) sr[i] = -sr[i];
    }
    const T alpha = contact_line_search(m, aref_f, os, Mm, a, Ma, fsm, sr, Mv, io.lsevals);
    if (alpha == T(0)) break;
#pragma unroll 1
    for (int i = 0; i < NV; i++) { a[i] += alpha * sr[i]; Ma[i] += alpha * Mv[i]; }
    const T oldcost = cost;
    cost = contact_cost(m, aref_f, os, a, Ma, fsm, asm_, qc, hd, active);
    T gg = T(0), nn = T(0);
#pragma unroll 1
    for (int i = 0; i < NV; i++) {
      const T g = Ma[i] - fsm[i] - qc[i];
      gg += g * g;
      if (Noise<T>::on) nn += fsm[i] * fsm[i] + qc[i] * qc[i];
    }
    T tol_i = m.tolerance, tol_g = m.tolerance;
    if (Noise<T>::on) {
      tol_i += T(Noise<T>::eps) * m.scale * abs_(cost);
      tol_g += T(Noise<T>::eps) * m.scale * sqrt_(nn);
    }
    iter++;
    if (m.scale * (oldcost - cost) < tol_i || m.scale * sqrt_(gg) < tol_g) break;
  }
  io.newton += iter;
  if (iter >= m.iterations) flags |= SO101_FLAG_MAXITER;
}


// ----------------------------------------------------------------------------------------------------------------------
// Fast path of the contact kernels (inline): the same problem as contact_solve, organised around the THREE Jacobian rows
// of each contact (normal, two tangents) instead of its four pyramid rows.  With rows J_n +- mu J_y, J_n +- mu J_x'
// (x' = -x) and activity bits a1..a4,
//   sum_active D J_r J_r' = D [ (a1+a2+a3+a4) nn' + mu (a1-a2)(ny' + yn') + mu (a3-a4)(nx'' + x'n') + mu^2 (a1+a2) yy'
//                              + mu^2 (a3+a4) x'x'' ],
// so one contact costs five 6x6 symmetric rank updates whatever its rows do, and a candidate's row residuals are three
// dot products.  The solve is the direct active-set iteration of contact_solve (pieces: qacc_warmstart's, then the
// rejected candidate's, up to six times); a step that does not settle returns 2 and takes contact_solve's safeguarded
// Newton.  0 = no hull touches the table.
// ----------------------------------------------------------------------------------------------------------------------

// Geometry half.  Every step in which a collision box is below the table top - mostly without a contact - needs only
// the exact hull test, and that needs no world frames: the height of a hull vertex is zo + zw . v with the world z axis
// zw and the origin height zo of the link frame, carried down the chain exactly as the tripwire does.  The hulls that do
// touch are noted in a HitList; only for them contact_rows forms world frames, contact points, Jacobian rows and row
// constants.  vcache (nullable): the support vertex each hull had when it was last tested - where the next walk starts.
template <typename T>
struct HitList {
  int n;
  int k[SO101_MAXTRIP], v[SO101_MAXTRIP];     // link and support vertex of the hulls that touch
  T dist[SO101_MAXTRIP], z[SO101_MAXTRIP];
};
template <typename T>
SO101_DEV void hull_test(const DevModel<T>& m, int k, int b, const T (&zw)[3], T zo, int32_t* vcache, HitList<T>& hl) {
  // "down" in the link frame, tilted by 1e-7 so that the vertices of an edge that lies parallel to the table are
  // ordered deterministically (same rule as the oracle's collision())
  const double d[3] = {sub_((double)m.con_tilt[k][0], (double)zw[0]), sub_((double)m.con_tilt[k][1], (double)zw[1]),
                       sub_((double)m.con_tilt[k][2], (double)zw[2])};
  const int g = m.trip_geom[k][b];
  T v[3];
  const int vi = hull_support(m, g, d, v, vcache ? vcache[g] : -1);
  if (vcache) vcache[g] = vi;
  const T zmin = add_(zo, dot3_(zw[0], v[0], zw[1], v[1], zw[2], v[2]));
  const T dist = sub_(zmin, m.trip_z);
  if (!(dist < m.con_margin)) return;
  hl.k[hl.n] = k; hl.v[hl.n] = vi; hl.dist[hl.n] = dist; hl.z[hl.n] = zmin;
  hl.n++;
}
// exact test of every box in `hits` (the one-warp kernels: the chain of (zw, zo) is walked again here)
template <typename T>
SO101_DEV void hull_tests(const DevModel<T>& m, const T* sn, const T* cs, int st, uint32_t hits, int32_t* vcache,
                          HitList<T>& hl) {
  T zw[3] = {T(0), T(0), T(1)}, zo = T(0);
#pragma unroll 1
  for (int k = 0; k < NV; k++) {
    if (!(hits >> (k * TRIP_PER_LINK))) break;            // no box further down the chain
    T R[9];
    make_R(m.E[k], cs[k * st], sn[k * st], R);
    tripwire_frame(R, m.r[k], zw, zo);
#pragma unroll 1
    for (int b = 0; b < m.trip_n[k]; b++)
      if (hits >> (k * TRIP_PER_LINK + b) & 1u) hull_test(m, k, b, zw, zo, vcache, hl);
  }
}
// the contacts of the hulls in the list -> con[], returns their number
template <typename T>
SO101_DEV int contact_rows(const DevModel<T>& m, const T* sn, const T* cs, int st, const T (&qd)[NV], const HitList<T>& hl,
                           Con3<T>* con, uint32_t& flags) {
  const int nhit = hl.n;
  const double* vert = reinterpret_cast<const double*>(hull_word_(m, 0));
  int ncon = 0, h = 0;
  T axw[NV][3], anw[NV][3];
  T Rw[9] = {T(1), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(1)}, ow[3] = {T(0), T(0), T(0)};
#pragma unroll 1
  for (int k = 0; k < NV && h < nhit; k++) {
    T R[9], Rn[9], o[3];
    make_R(m.E[k], cs[k * st], sn[k * st], R);
    rot(Rw, m.r[k], o);
    ow[0] = add_(ow[0], o[0]); ow[1] = add_(ow[1], o[1]); ow[2] = add_(ow[2], o[2]);
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) Rn[3 * i + j] = dot3_(Rw[3 * i], R[j], Rw[3 * i + 1], R[3 + j], Rw[3 * i + 2], R[6 + j]);
#pragma unroll
    for (int i = 0; i < 9; i++) Rw[i] = Rn[i];
    axw[k][0] = Rw[2]; axw[k][1] = Rw[5]; axw[k][2] = Rw[8];
    anw[k][0] = ow[0]; anw[k][1] = ow[1]; anw[k][2] = ow[2];
#pragma unroll 1
    for (; h < nhit && hl.k[h] == k; h++) {
      const T dist = hl.dist[h];
      const T v[3] = {(T)vert[3 * hl.v[h]], (T)vert[3 * hl.v[h] + 1], (T)vert[3 * hl.v[h] + 2]};
      T p[3];
      rot(Rw, v, p);
      p[0] = add_(p[0], ow[0]); p[1] = add_(p[1], ow[1]); p[2] = fma_(T(-0.5), dist, hl.z[h]);
      if (p[0] < m.con_box[0] || p[0] > m.con_box[1] || p[1] < m.con_box[2] || p[1] > m.con_box[3] || ncon == MAXCON) {
        flags |= SO101_FLAG_TRIP_TABLE;   // an edge of the table, or more contacts than rows: not simulated
        continue;
      }
      Con3<T>& c = con[ncon++];
      T vn = T(0), vy = T(0), vx = T(0);
#pragma unroll 1
      for (int j = 0; j < NV; j++) {
        T jn = T(0), jy = T(0), jx = T(0);
        if (j <= k) {
          const T rx = sub_(p[0], anw[j][0]), ry = sub_(p[1], anw[j][1]), rz = sub_(p[2], anw[j][2]);
          jx = -det2_(axw[j][1], rz, axw[j][2], ry);
          jy = det2_(axw[j][2], rx, axw[j][0], rz);
          jn = det2_(axw[j][0], ry, axw[j][1], rx);
        }
        c.Jn[j] = jn; c.Jy[j] = jy; c.Jx[j] = jx;
        vn = fma_(jn, qd[j], vn); vy = fma_(jy, qd[j], vy); vx = fma_(jx, qd[j], vx);
      }
      c.vn = vn; c.vy = vy; c.vx = vx;
      const T imp = limit_impedance(m.con_imp, dist, m.con_margin);
      const T mu = m.con_mu;
      const T tr = fma_(mul_(mu, mu), m.con_tran[k], m.con_tran[k]);
      const T R1 = max_(T(MJ_MINVAL), mul_(sub_(T(1), imp), tr) / imp);
      c.D = T(1) / mul_(mul_(T(2), mul_(mu, mu)), R1);
      c.c0 = mul_(mul_(-m.con_K, imp), sub_(dist, m.con_margin));
    }
  }
  return ncon;
}
template <typename T>
SO101_DEV int contact_geometry(const DevModel<T>& m, const T* sn, const T* cs, int st, const T (&qd)[NV], uint32_t hits,
                               Con3<T>* con, uint32_t& flags, int32_t* vcache = nullptr) {
  HitList<T> hl;
  hl.n = 0;
  hull_tests(m, sn, cs, st, hits, vcache, hl);
  if (hl.n == 0) return 0;
  return contact_rows(m, sn, cs, st, qd, hl, con, flags);
}

// Solve half: direct active-set iteration over friction, limit and contact rows.  1 = solved (a, qc set), 2 = did not
// settle (the caller takes contact_solve's safeguarded Newton).
template <typename T>
SO101_DEV int contact_active_set(const DevModel<T>& m, const Con3<T>* con, int ncon, const T (&start)[NV], bool prox,
                                 const T (&Mm)[21], const T (&fsm)[NV], const Rows<T>& rw, T (&a)[NV], T (&qc)[NV],
                                 Counters& cnt) {
  // First piece: the one-sided rows as they are at `start`; the friction rows likewise, or (prox: start = qacc_smooth,
  // possibly from a lagged factor of M) by the per-dof rule of active_set_guess.  The starting piece only decides the
  // number of solves, never the result.  Measured on the oracle's contact steps of the benchmark workload: from
  // qacc_warmstart 2.6 solves on average, from qacc_smooth with the per-dof friction guess 2.05.
  const T mu = m.con_mu;
  // pieces: friction zones (quadratic / saturated positive), limit rows, 4 bits per contact
  uint32_t zq = 0, zp = 0, zl = 0, zc = 0;
  T az[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) az[i] = start[i];
#pragma unroll 1
  for (int attempt = 0; attempt <= ACTIVE_SET_ATTEMPTS; attempt++) {
    uint32_t nzq = 0, nzp = 0, nzl = 0, nzc = 0;
    bool strict = true;
    const bool guess = prox && attempt == 0;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      T jar = az[i] - rw.aref_f[i], edge = m.fr_Rf[i];
      if (guess) { jar *= Mm[tri(i, i)]; edge = m.fr_Rf[i] * Mm[tri(i, i)] + m.fr_f[i]; }
      if (m.fr_f[i] == T(0) || abs_(jar) < edge) nzq |= 1u << i;
      else if (jar > T(0)) nzp |= 1u << i;
      strict &= m.fr_f[i] == T(0) || abs_(jar) != edge;
    }
    if (rw.anylim) {
#pragma unroll 1
      for (int i = 0; i < NV; i++)
        if (rw.side[i] != T(0) && rw.side[i] * az[i] - rw.aref_l[i] < T(0)) nzl |= 1u << i;
    }
#pragma unroll 1
    for (int c = 0; c < ncon; c++) {
      T jn = T(0), jy = T(0), jx = T(0);
#pragma unroll
      for (int j = 0; j < NV; j++) { jn += con[c].Jn[j] * az[j]; jy += con[c].Jy[j] * az[j]; jx += con[c].Jx[j] * az[j]; }
      // jar_r = J_r a - aref_r, aref_r = -B (J_r qd) + c0
      const T bn = jn + m.con_B * con[c].vn - con[c].c0, by = mu * (jy + m.con_B * con[c].vy), bx = mu * (jx + m.con_B * con[c].vx);
      if (bn + by < T(0)) nzc |= 1u << (4 * c);
      if (bn - by < T(0)) nzc |= 2u << (4 * c);
      if (bn + bx < T(0)) nzc |= 4u << (4 * c);
      if (bn - bx < T(0)) nzc |= 8u << (4 * c);
    }
    if (attempt > 0 && nzq == zq && nzp == zp && nzl == zl && nzc == zc && strict) {   // the candidate lies in its own piece
      // qfrc_constraint at the minimiser is M a - qfrc_smooth (stationarity of the piece): 36 multiply-adds instead of
      // every row's force again - and a quarter less of this rare code in the instruction cache (profiles/r2_team_timing.txt)
      T Ma[NV];
      symv6(Mm, az, Ma);
#pragma unroll
      for (int i = 0; i < NV; i++) { a[i] = az[i]; qc[i] = Ma[i] - fsm[i]; }
      cnt.newton += attempt;
      return 1;
    }
    if (attempt == ACTIVE_SET_ATTEMPTS) break;
    if (attempt >= 3) active_set_unflip(zq, zp, nzq, nzp);
    zq = nzq; zp = nzp; zl = nzl; zc = nzc;
    T H[21], rhs[NV];
#pragma unroll
    for (int i = 0; i < 21; i++) H[i] = Mm[i];
#pragma unroll
    for (int i = 0; i < NV; i++) {
      rhs[i] = fsm[i];
      if (m.fr_f[i] != T(0)) {
        if (zq >> i & 1u) { H[tri(i, i)] += m.fr_D[i]; rhs[i] += m.fr_D[i] * rw.aref_f[i]; }
        else rhs[i] += (zp >> i & 1u) ? -m.fr_f[i] : m.fr_f[i];
      }
    }
    if (rw.anylim) {
#pragma unroll 1
      for (int i = 0; i < NV; i++)
        if (zl >> i & 1u) { H[tri(i, i)] += rw.D_l[i]; rhs[i] += rw.D_l[i] * rw.aref_l[i] * rw.side[i]; }
    }
#pragma unroll 1
    for (int c = 0; c < ncon; c++) {
      const uint32_t bits = zc >> (4 * c) & 15u;
      if (!bits) continue;
      const T a1 = (bits & 1u) ? T(1) : T(0), a2 = (bits & 2u) ? T(1) : T(0), a3 = (bits & 4u) ? T(1) : T(0), a4 = (bits & 8u) ? T(1) : T(0);
      const T D = con[c].D;
      const T wnn = D * (a1 + a2 + a3 + a4), wny = D * mu * (a1 - a2), wnx = D * mu * (a3 - a4);
      const T wyy = D * mu * mu * (a1 + a2), wxx = D * mu * mu * (a3 + a4);
      // aref of the four rows
      const T an = -m.con_B * con[c].vn + con[c].c0, ay = -m.con_B * mu * con[c].vy, ax = -m.con_B * mu * con[c].vx;
      const T r1 = an + ay, r2 = an - ay, r3 = an + ax, r4 = an - ax;
      const T gn = D * (a1 * r1 + a2 * r2 + a3 * r3 + a4 * r4), gy = D * mu * (a1 * r1 - a2 * r2), gx = D * mu * (a3 * r3 - a4 * r4);
      T n_[NV], y_[NV], x_[NV];
#pragma unroll
      for (int j = 0; j < NV; j++) { n_[j] = con[c].Jn[j]; y_[j] = con[c].Jy[j]; x_[j] = con[c].Jx[j]; }
#pragma unroll
      for (int i = 0; i < NV; i++) {
        rhs[i] += gn * n_[i] + gy * y_[i] + gx * x_[i];
        const T pi = wnn * n_[i] + wny * y_[i] + wnx * x_[i];     // coefficient of n_j
        const T qi = wny * n_[i] + wyy * y_[i];                   // coefficient of y_j
        const T ri = wnx * n_[i] + wxx * x_[i];                   // coefficient of x_j
#pragma unroll
        for (int j = 0; j <= i; j++) H[tri(i, j)] += pi * n_[j] + qi * y_[j] + ri * x_[j];
      }
    }
    ldl6_factor_solve(H, rhs);
#pragma unroll
    for (int i = 0; i < NV; i++) az[i] = rhs[i];
  }
  return 2;
}

}  // namespace so101
