// so101_physics.cuh — one mj_step-equivalent physics step for one env, held by one thread.
//
// Pipeline (reference: mujoco.mj_step as called at SOARM101/SOARM101_Env.py:132; algorithm
// restated in oracle/so101_oracle.c):
//   checkPos/Vel -> smooth dynamics (RNEA bias + CRBA mass matrix, link-local frames)
//   -> passive + affine actuators -> qacc_smooth = M^-1 qfrc_smooth
//   -> friction-loss + joint-limit rows -> warm start pick -> Newton with exact line search
//   -> checkAcc -> Euler with implicit joint damping.
// Everything is fully unrolled over the 6 links so that table reads are constant-bank
// operands and all per-env quantities are registers (or L1-resident spills).
#pragma once
#include <cuda_runtime.h>

#include "so101_model.h"

namespace so101 {

#define SO101_DEV __device__ __forceinline__
#ifndef SO101_ONEWARP_ROLLED
#define SO101_ONEWARP_ROLLED 0
#endif
#ifndef SO101_RNEA_UNROLL
#define SO101_RNEA_UNROLL 1   // link-loop unroll factor of the compact RNEA (team kernels)
#endif
#ifndef SO101_CRBA_BRANCHLESS
#define SO101_CRBA_BRANCHLESS 1
#endif
#define MJ_MINVAL 1e-15
#define MJ_MAXVAL 1e10

// ------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------
SO101_DEV double rcp_(double x) {
  // MUFU.RCP64H seed (~20 bits) + two Newton steps: <= 1 ulp, ~7 instructions instead of the
  // ~20 of an IEEE division.  Arguments here are diagonal pivots / curvatures, never subnormal.
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  return y;
}
SO101_DEV float rcp_(float x) { return __frcp_rn(x); }
// sin/cos by Cody-Waite reduction to [-pi/4, pi/4] + fdlibm / Cephes kernels (<= 1 ulp for the
// joint angles of a limited hinge chain; arguments beyond 2^18 fall back to the library, which
// keeps its large-argument reduction out of the hot instruction stream).
static __device__ __noinline__ void sincos_slow_(double x, double* sp, double* cp) { sincos(x, sp, cp); }
static __device__ __noinline__ void sincos_slow_(float x, float* sp, float* cp) { sincosf(x, sp, cp); }
SO101_DEV void sincos_(double x, double* sp, double* cp) {
  if (!(fabs(x) < 262144.0)) { sincos_slow_(x, sp, cp); return; }
  double k = rint(x * 0.63661977236758134308);
  double r = fma(-k, 1.57079632679489655800e+00, x);
  r = fma(-k, 6.12323399573676603587e-17, r);
  double z = r * r;
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(z, ps, 2.75573137070700676789e-06);
  ps = fma(z, ps, -1.98412698298579493134e-04);
  ps = fma(z, ps, 8.33333333332248946124e-03);
  ps = fma(z, ps, -1.66666666666666324348e-01);
  double s = fma(r * z, ps, r);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(z, pc, -2.75573143513906633035e-07);
  pc = fma(z, pc, 2.48015872894767294178e-05);
  pc = fma(z, pc, -1.38888888888741095749e-03);
  pc = fma(z, pc, 4.16666666666666019037e-02);
  double c = fma(z * z, pc, fma(z, -0.5, 1.0));
  int q = (int)k;
  double ss = (q & 1) ? c : s, cc = (q & 1) ? s : c;
  *sp = (q & 2) ? -ss : ss;
  *cp = ((q + 1) & 2) ? -cc : cc;
}
SO101_DEV void sincos_(float x, float* sp, float* cp) {
  if (!(fabsf(x) < 8192.0f)) { sincos_slow_(x, sp, cp); return; }
  float k = rintf(x * 0.636619772f);
  float r = fmaf(-k, 1.5703125f, x);
  r = fmaf(-k, 4.837512969970703125e-4f, r);
  r = fmaf(-k, 7.54978995489188e-8f, r);
  float z = r * r;
  float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(z, ps, -1.6666654611e-1f);
  float s = fmaf(r * z, ps, r);
  float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(z, pc, 4.166664568298827e-2f);
  float c = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
  int q = (int)k;
  float ss = (q & 1) ? c : s, cc = (q & 1) ? s : c;
  *sp = (q & 2) ? -ss : ss;
  *cp = ((q + 1) & 2) ? -cc : cc;
}
SO101_DEV double sqrt_(double x) { return sqrt(x); }
SO101_DEV float sqrt_(float x) { return sqrtf(x); }
SO101_DEV double copysign_(double a, double b) { return copysign(a, b); }
SO101_DEV float copysign_(float a, float b) { return copysignf(a, b); }
SO101_DEV double abs_(double x) { return fabs(x); }
SO101_DEV float abs_(float x) { return fabsf(x); }
// relative noise floor added to the solver's stopping tests: 0 in fp64 (MuJoCo's tests verbatim);
// in fp32 the absolute tolerances 1e-8 sit below rounding noise and would never fire.
template <typename T> struct Noise { static constexpr bool on = false; static constexpr double eps = 0.0; };
template <> struct Noise<float> { static constexpr bool on = true; static constexpr double eps = 1e-5; };
template <typename T> SO101_DEV T min_(T a, T b) { return a < b ? a : b; }
template <typename T> SO101_DEV T max_(T a, T b) { return a > b ? a : b; }
template <typename T> SO101_DEV bool bad_(T x) { return !(x <= T(MJ_MAXVAL) && x >= T(-MJ_MAXVAL)); }

// index into a packed lower triangle, i >= j
__host__ __device__ constexpr int tri(int i, int j) { return i * (i + 1) / 2 + j; }

// ------------------------------------------------------------------------------------------
// Rounding-explicit arithmetic for the smooth dynamics.  The same link bodies are compiled in several
// contexts (fully unrolled in the one-warp kernels, rolled over the link index in the team kernels);
// every multiply-add below is written as an explicit fma / mul / add intrinsic, so that where the
// compiler contracts is not a per-context choice and all kernels produce the same bits.
// ------------------------------------------------------------------------------------------
SO101_DEV double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }
SO101_DEV float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SO101_DEV double mul_(double a, double b) { return __dmul_rn(a, b); }
SO101_DEV float mul_(float a, float b) { return __fmul_rn(a, b); }
SO101_DEV double add_(double a, double b) { return __dadd_rn(a, b); }
SO101_DEV float add_(float a, float b) { return __fadd_rn(a, b); }
SO101_DEV double sub_(double a, double b) { return __dsub_rn(a, b); }
SO101_DEV float sub_(float a, float b) { return __fsub_rn(a, b); }
template <typename T> SO101_DEV T dot3_(T a0, T b0, T a1, T b1, T a2, T b2) {   // (a0 b0 + a1 b1) + a2 b2
  return fma_(a2, b2, fma_(a1, b1, mul_(a0, b0)));
}
template <typename T> SO101_DEV T det2_(T a, T b, T c, T d) { return fma_(-c, d, mul_(a, b)); }   // a b - c d

// ------------------------------------------------------------------------------------------
// 3-vector / rotation helpers.  R (row-major 3x3) maps child coordinates to parent coordinates.
// ------------------------------------------------------------------------------------------
template <typename T> SO101_DEV void make_R(const T (&E)[9], T c, T s, T (&R)[9]) {
#pragma unroll
  for (int i = 0; i < 3; i++) {
    R[3 * i + 0] = fma_(s, E[3 * i + 1], mul_(c, E[3 * i]));
    R[3 * i + 1] = det2_(c, E[3 * i + 1], s, E[3 * i]);
    R[3 * i + 2] = E[3 * i + 2];
  }
}
template <typename T> SO101_DEV void rot(const T (&R)[9], const T* v, T* o) {  // o = R v
  T a = dot3_(R[0], v[0], R[1], v[1], R[2], v[2]);
  T b = dot3_(R[3], v[0], R[4], v[1], R[5], v[2]);
  T c = dot3_(R[6], v[0], R[7], v[1], R[8], v[2]);
  o[0] = a; o[1] = b; o[2] = c;
}
template <typename T> SO101_DEV void rotT(const T (&R)[9], const T* v, T* o) {  // o = R^T v
  T a = dot3_(R[0], v[0], R[3], v[1], R[6], v[2]);
  T b = dot3_(R[1], v[0], R[4], v[1], R[7], v[2]);
  T c = dot3_(R[2], v[0], R[5], v[1], R[8], v[2]);
  o[0] = a; o[1] = b; o[2] = c;
}
// spatial motion vector parent -> child: ang' = R^T ang, lin' = R^T (lin + ang x r)
template <typename T> SO101_DEV void xmotion(const T (&R)[9], const T (&r)[3], const T (&p)[6], T (&c)[6]) {
  T u[3] = {add_(p[3], det2_(p[1], r[2], p[2], r[1])), add_(p[4], det2_(p[2], r[0], p[0], r[2])),
            add_(p[5], det2_(p[0], r[1], p[1], r[0]))};
  rotT(R, &p[0], &c[0]);
  rotT(R, u, &c[3]);
}
// spatial force vector child -> parent: lin' = R lin, ang' = R ang + r x lin'
template <typename T> SO101_DEV void xforce(const T (&R)[9], const T (&r)[3], const T (&c)[6], T (&p)[6]) {
  T a[3], l[3];
  rot(R, &c[0], a);
  rot(R, &c[3], l);
  p[0] = add_(a[0], det2_(r[1], l[2], r[2], l[1]));
  p[1] = add_(a[1], det2_(r[2], l[0], r[0], l[2]));
  p[2] = add_(a[2], det2_(r[0], l[1], r[1], l[0]));
  p[3] = l[0]; p[4] = l[1]; p[5] = l[2];
}
// 10-parameter spatial inertia times motion vector (same layout as mju_mulInertVec)
template <typename T> SO101_DEV void inert_mul(const T (&i)[10], const T (&v)[6], T (&o)[6]) {
  o[0] = fma_(i[7], v[5], fma_(-i[8], v[4], dot3_(i[0], v[0], i[3], v[1], i[4], v[2])));
  o[1] = fma_(-i[6], v[5], fma_(i[8], v[3], dot3_(i[3], v[0], i[1], v[1], i[5], v[2])));
  o[2] = fma_(i[6], v[4], fma_(-i[7], v[3], dot3_(i[4], v[0], i[5], v[1], i[2], v[2])));
  o[3] = fma_(i[9], v[3], det2_(i[8], v[1], i[7], v[2]));
  o[4] = fma_(i[9], v[4], det2_(i[6], v[2], i[8], v[0]));
  o[5] = fma_(i[9], v[5], det2_(i[7], v[0], i[6], v[1]));
}
// composite inertia child -> parent, accumulated into the parent's
template <typename T> SO101_DEV void xinertia_add(const T (&R)[9], const T (&r)[3], const T (&c)[10], T (&p)[10]) {
  // rotate the symmetric 3x3 block: S = R * I * R^T
  T A[9];  // A = R * I
#pragma unroll
  for (int i = 0; i < 3; i++) {
    A[3 * i + 0] = dot3_(R[3 * i], c[0], R[3 * i + 1], c[3], R[3 * i + 2], c[4]);
    A[3 * i + 1] = dot3_(R[3 * i], c[3], R[3 * i + 1], c[1], R[3 * i + 2], c[5]);
    A[3 * i + 2] = dot3_(R[3 * i], c[4], R[3 * i + 1], c[5], R[3 * i + 2], c[2]);
  }
  T Sxx = dot3_(A[0], R[0], A[1], R[1], A[2], R[2]);
  T Syy = dot3_(A[3], R[3], A[4], R[4], A[5], R[5]);
  T Szz = dot3_(A[6], R[6], A[7], R[7], A[8], R[8]);
  T Sxy = dot3_(A[0], R[3], A[1], R[4], A[2], R[5]);
  T Sxz = dot3_(A[0], R[6], A[1], R[7], A[2], R[8]);
  T Syz = dot3_(A[3], R[6], A[4], R[7], A[5], R[8]);
  T h[3];
  rot(R, &c[6], h);
  // shift the reference point by r: I += 2(u.r)1 - (r u' + u r'), u = h + m r / 2
  T hm = mul_(T(0.5), c[9]);
  T u0 = fma_(hm, r[0], h[0]), u1 = fma_(hm, r[1], h[1]), u2 = fma_(hm, r[2], h[2]);
  T d0 = mul_(u0, r[0]), d1 = mul_(u1, r[1]), d2 = mul_(u2, r[2]);
  p[0] = add_(p[0], fma_(T(2), add_(d1, d2), Sxx));
  p[1] = add_(p[1], fma_(T(2), add_(d0, d2), Syy));
  p[2] = add_(p[2], fma_(T(2), add_(d0, d1), Szz));
  p[3] = add_(p[3], sub_(Sxy, fma_(u0, r[1], mul_(r[0], u1))));
  p[4] = add_(p[4], sub_(Sxz, fma_(u0, r[2], mul_(r[0], u2))));
  p[5] = add_(p[5], sub_(Syz, fma_(u1, r[2], mul_(r[1], u2))));
  p[6] = add_(p[6], fma_(c[9], r[0], h[0]));
  p[7] = add_(p[7], fma_(c[9], r[1], h[1]));
  p[8] = add_(p[8], fma_(c[9], r[2], h[2]));
  p[9] = add_(p[9], c[9]);
}
// RNEA forward step of one link: v, a (parent frame, in/out: this link's frame), joint rate w -> link force f
template <typename T>
SO101_DEV void rnea_link(const T (&R)[9], const T (&r)[3], const T (&I)[10], T w, T (&v)[6], T (&a)[6], T* f) {
  T vc[6], ac[6];
  xmotion(R, r, v, vc);
  xmotion(R, r, a, ac);
  vc[2] = add_(vc[2], w);
  // a += v x (S qd), S = [0 0 1 | 0 0 0]
  ac[0] = fma_(w, vc[1], ac[0]);
  ac[1] = fma_(-w, vc[0], ac[1]);
  ac[3] = fma_(w, vc[4], ac[3]);
  ac[4] = fma_(-w, vc[3], ac[4]);
  T Ia[6], Iv[6];
  inert_mul(I, ac, Ia);
  inert_mul(I, vc, Iv);
  // f = I a + v x* (I v)
  f[0] = add_(add_(Ia[0], det2_(vc[1], Iv[2], vc[2], Iv[1])), det2_(vc[4], Iv[5], vc[5], Iv[4]));
  f[1] = add_(add_(Ia[1], det2_(vc[2], Iv[0], vc[0], Iv[2])), det2_(vc[5], Iv[3], vc[3], Iv[5]));
  f[2] = add_(add_(Ia[2], det2_(vc[0], Iv[1], vc[1], Iv[0])), det2_(vc[3], Iv[4], vc[4], Iv[3]));
  f[3] = add_(Ia[3], det2_(vc[1], Iv[5], vc[2], Iv[4]));
  f[4] = add_(Ia[4], det2_(vc[2], Iv[3], vc[0], Iv[5]));
  f[5] = add_(Ia[5], det2_(vc[0], Iv[4], vc[1], Iv[3]));
#pragma unroll
  for (int c = 0; c < 6; c++) { v[c] = vc[c]; a[c] = ac[c]; }
}
// contact tripwire of one link: zw = world z axis in the link frame, zo = world height of its origin.
// A box that dips below the table top sets bit k * TRIP_PER_LINK + b of `hits`: the caller either runs the contact
// path on exactly those hulls (so101_contact.cuh) or, without hull data, raises SO101_FLAG_TRIP_TABLE.
template <typename T, typename BOX>
SO101_DEV void tripwire_box(const BOX& m, int k, int b, const T (&zw)[3], T zo, uint32_t& hits) {
  T zc0 = add_(zo, dot3_(zw[0], m.trip_c[k][b][0], zw[1], m.trip_c[k][b][1], zw[2], m.trip_c[k][b][2]));
  if (sub_(zc0, m.trip_rad[k][b]) >= m.trip_z) return;   // bounding sphere clears the plane: box does too
  T ext = T(0);
#pragma unroll
  for (int ax = 0; ax < 3; ax++)
    ext = fma_(abs_(dot3_(zw[0], m.trip_ax[k][b][3 * ax], zw[1], m.trip_ax[k][b][3 * ax + 1], zw[2],
                          m.trip_ax[k][b][3 * ax + 2])), m.trip_half[k][b][ax], ext);
  if (sub_(zc0, ext) < m.trip_z) hits |= 1u << (k * TRIP_PER_LINK + b);
}
template <typename T> SO101_DEV void tripwire_frame(const T (&R)[9], const T (&r)[3], T (&zw)[3], T& zo) {
  zo = add_(zo, dot3_(zw[0], r[0], zw[1], r[1], zw[2], r[2]));
  T zc[3];
  rotT(R, zw, zc);
  zw[0] = zc[0]; zw[1] = zc[1]; zw[2] = zc[2];
}

// Self-collision flag.  The joint box trip_qlo..trip_qhi is the fast accept (inside it no pair of non-adjacent link hulls
// intersected in the host's sampling, tripwire.py); a pose outside it (outside_self_box, at the start of a step) gets the test below:
// do the oriented boxes of two colliding geoms on non-adjacent links overlap?  Separating-axis test, exact for the boxes
// and conservative for the hulls inside them (box contains hull contains mesh), restated in tripwire.self_overlap_numpy.  Rare, out
// of line and called where little is live (a call in the middle of the step cost the float32 kernels 40 %: 128 registers),
// in double for both dtypes; world frames of the links from the joint sines / cosines.
template <typename T> SO101_DEV bool outside_self_box(const DevModel<T>& m, const T (&q)[NV]) {
  bool out = false;
#pragma unroll
  for (int k = 0; k < NV; k++) out |= (q[k] < m.trip_qlo[k]) | (q[k] > m.trip_qhi[k]);
  return out;
}
// constants of box g of the flat list: a tripwire box of its link (slot sl >= 0) or one of the extra boxes.  Indexed reads of
// the kernel parameter (constant bank); pointers into it would turn every read into a generic load.
template <typename T> struct SelfBox {
  const DevModel<T>& m; int k, sl;
  SO101_DEV double c(int i) const { return sl >= 0 ? (double)m.trip_c[k][sl][i] : (double)m.sx_c[-1 - sl][i]; }
  SO101_DEV double ax(int i) const { return sl >= 0 ? (double)m.trip_ax[k][sl][i] : (double)m.sx_ax[-1 - sl][i]; }
  SO101_DEV double half(int i) const { return sl >= 0 ? (double)m.trip_half[k][sl][i] : (double)m.sx_half[-1 - sl][i]; }
  SO101_DEV double rad() const { return sl >= 0 ? (double)m.trip_rad[k][sl] : (double)m.sx_rad[-1 - sl]; }
};
template <typename T> SO101_DEV SelfBox<T> self_box(const DevModel<T>& m, int g) { return SelfBox<T>{m, m.sb_link[g], m.sb_slot[g]}; }
// The double instantiation keeps everything in registers (fully unrolled: ~4 k instructions when a handful of pairs pass the
// sphere pre-check); the float one rolls its loops over local arrays: what a callee needs in registers is taken from its
// caller, and the float32 one-warp kernels sit at 128 registers (unrolled they spilled 80 % more and lost 35 % on every
// workload; rolled the test itself is ~10x slower, on the few steps that run it).
// Returns < 0 when two such boxes overlap.  Otherwise it leaves, in the env's hull-cache words, lower bounds in radians of joint
// travel of how far the arm is from two boxes touching: per pair, gap / lever arm with gap = their distance along a separating
// unit axis and lever arm = the largest distance of a point of the outer box from the axes of the joints between the two
// links (only those joints move one box against the other) - the smallest of them with its pair (the critical pair), and the
// smallest over all other pairs.  The caller spends h * sum |qvel| of both per step: while the second one lasts only the
// critical pair can have closed, and `only` = that pair tests it alone (a sixth of the work); the full test runs again when
// the second budget is gone, i.e. every few dozen steps for an arm that is not about to fold onto itself.
template <typename T>
__device__ __noinline__ double self_boxes_overlap(const DevModel<T>& m, T s0, T s1, T s2, T s3, T s4, T s5, T c0, T c1, T c2, T c3,
                                                T c4, T c5, int32_t* vcache, int only) {
  const int oa = only >> 4, ob = only & 15;      // only < 0: every pair
  constexpr int UR3 = sizeof(T) == 8 ? 3 : 1, UR9 = sizeof(T) == 8 ? 9 : 1, URNV = sizeof(T) == 8 ? NV : 1;
  double Rw[NV][9], cw[SO101_MAXTRIP][3];
  {
    const double sn[NV] = {(double)s0, (double)s1, (double)s2, (double)s3, (double)s4, (double)s5};
    const double cs[NV] = {(double)c0, (double)c1, (double)c2, (double)c3, (double)c4, (double)c5};
    double Rp[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, pp[3] = {0, 0, 0};
#pragma unroll(URNV)
    for (int k = 0; k < NV; k++) {
      const double c = cs[k], s_ = sn[k];
      double R[9], Rn[9];     // v_{k-1} = E_k Rz(theta_k) v_k
#pragma unroll(UR3)
      for (int i = 0; i < 3; i++) {
        const double e0 = (double)m.E[k][3 * i], e1 = (double)m.E[k][3 * i + 1];
        R[3 * i] = c * e0 + s_ * e1; R[3 * i + 1] = c * e1 - s_ * e0; R[3 * i + 2] = (double)m.E[k][3 * i + 2];
      }
#pragma unroll(UR3)
      for (int i = 0; i < 3; i++) pp[i] += Rp[3 * i] * (double)m.r[k][0] + Rp[3 * i + 1] * (double)m.r[k][1] + Rp[3 * i + 2] * (double)m.r[k][2];
#pragma unroll(UR3)
      for (int i = 0; i < 3; i++)
#pragma unroll(UR3)
        for (int j = 0; j < 3; j++) Rn[3 * i + j] = Rp[3 * i] * R[j] + Rp[3 * i + 1] * R[3 + j] + Rp[3 * i + 2] * R[6 + j];
#pragma unroll(UR9)
      for (int i = 0; i < 9; i++) { Rp[i] = Rn[i]; Rw[k][i] = Rn[i]; }
#pragma unroll 1
      for (int g = 0; g < m.sb_n; g++) {
        if (m.sb_link[g] != k || (only >= 0 && g != oa && g != ob)) continue;
        const SelfBox<T> bx = self_box(m, g);
#pragma unroll(UR3)
        for (int i = 0; i < 3; i++)
          cw[g][i] = pp[i] + Rn[3 * i] * bx.c(0) + Rn[3 * i + 1] * bx.c(1) + Rn[3 * i + 2] * bx.c(2);
      }
    }
  }
  double sep = 1.0, sep2 = 1.0;   // radians: smallest (its pair: cp) and second smallest
  int cp = -1;
#pragma unroll 1
  for (int a = only >= 0 ? oa : 0; a + 1 < m.sb_n && (only < 0 || a == oa); a++) {
    const int ka = m.sb_link[a];
    const SelfBox<T> ba = self_box(m, a);
#pragma unroll 1
    for (int b = only >= 0 ? ob : a + 1; b < m.sb_n && (only < 0 || b == ob); b++) {
      const int kb = m.sb_link[b];
      if (kb - ka < 2 && ka - kb < 2) continue;               // same or adjacent links never collide (filterparent)
      double ps;                                              // this pair's bound
      const SelfBox<T> bb = self_box(m, b);
      // lever arm of the joints between the two links on the box of the outer link
      double rho;
      {
        const SelfBox<T>& bo = kb > ka ? bb : ba;
        const int klo = kb > ka ? ka : kb, khi = kb > ka ? kb : ka;
        rho = sqrt(bo.c(0) * bo.c(0) + bo.c(1) * bo.c(1) + bo.c(2) * bo.c(2)) + bo.rad();
#pragma unroll 1
        for (int k = klo + 2; k <= khi; k++) rho += (double)m.self_rlen[k];
      }
      double d[3], Aa[9], Ab[9], R[9], AR[9], tv[3], ha[3], hb[3];
      double dd = 0.0;
#pragma unroll(UR3)
      for (int i = 0; i < 3; i++) { d[i] = cw[b][i] - cw[a][i]; dd += d[i] * d[i]; }
      const double rr = ba.rad() + bb.rad();
      if (dd > rr * rr) {                                     // bounding spheres apart
        ps = (sqrt(dd) - rr) / rho;
        if (ps < sep) { sep2 = sep; sep = ps; cp = 16 * a + b; } else sep2 = fmin(sep2, ps);
        continue;
      }
#pragma unroll(UR3)
      for (int i = 0; i < 3; i++) {                           // box axes (rows) in the world frame
        ha[i] = ba.half(i); hb[i] = bb.half(i);
#pragma unroll(UR3)
        for (int c = 0; c < 3; c++) {
          Aa[3 * i + c] = Rw[ka][3 * c] * ba.ax(3 * i) + Rw[ka][3 * c + 1] * ba.ax(3 * i + 1) + Rw[ka][3 * c + 2] * ba.ax(3 * i + 2);
          Ab[3 * i + c] = Rw[kb][3 * c] * bb.ax(3 * i) + Rw[kb][3 * c + 1] * bb.ax(3 * i + 1) + Rw[kb][3 * c + 2] * bb.ax(3 * i + 2);
        }
      }
      bool apart = false;
      double gap = 0.0;                                       // along a face normal (unit axis): a true distance bound
#pragma unroll(UR3)
      for (int i = 0; i < 3; i++) {                           // rotation b -> a, translation in a's axes
        tv[i] = Aa[3 * i] * d[0] + Aa[3 * i + 1] * d[1] + Aa[3 * i + 2] * d[2];
#pragma unroll(UR3)
        for (int j = 0; j < 3; j++) {
          R[3 * i + j] = Aa[3 * i] * Ab[3 * j] + Aa[3 * i + 1] * Ab[3 * j + 1] + Aa[3 * i + 2] * Ab[3 * j + 2];
          AR[3 * i + j] = fabs(R[3 * i + j]) + 1e-12;
        }
      }
#pragma unroll(UR3)
      for (int i = 0; i < 3; i++) gap = fmax(gap, fabs(tv[i]) - (ha[i] + AR[3 * i] * hb[0] + AR[3 * i + 1] * hb[1] + AR[3 * i + 2] * hb[2]));
#pragma unroll(UR3)
      for (int j = 0; j < 3; j++)
        gap = fmax(gap, fabs(tv[0] * R[j] + tv[1] * R[3 + j] + tv[2] * R[6 + j]) - (ha[0] * AR[j] + ha[1] * AR[3 + j] + ha[2] * AR[6 + j] + hb[j]));
      if (gap > 0.0) {                                        // a face normal separates (the common case)
        ps = gap / rho;
        if (ps < sep) { sep2 = sep; sep = ps; cp = 16 * a + b; } else sep2 = fmin(sep2, ps);
        continue;
      }
#pragma unroll(UR3)
      for (int i = 0; i < 3; i++) {
        const int i1 = i == 2 ? 0 : i + 1, i2 = i == 0 ? 2 : i - 1;
#pragma unroll(UR3)
        for (int j = 0; j < 3; j++) {
          const int j1 = j == 2 ? 0 : j + 1, j2 = j == 0 ? 2 : j - 1;
          const double ra = ha[i1] * AR[3 * i2 + j] + ha[i2] * AR[3 * i1 + j];
          const double rb = hb[j1] * AR[3 * i + j2] + hb[j2] * AR[3 * i + j1];
          apart |= fabs(tv[i2] * R[3 * i1 + j] - tv[i1] * R[3 * i2 + j]) > ra + rb;
        }
      }
      if (!apart) return -1.0;
      sep2 = sep; sep = 0.0; cp = 16 * a + b;                 // only an edge-edge axis separates: no distance bound, test again next step
    }
  }
  vcache[SELF_BUDGET_SLOT] = __float_as_int(sep > 0.0 ? 0.999f * (float)sep : 0.f);
  if (only < 0) {
    vcache[SELF_REST_SLOT] = __float_as_int(0.999f * (float)sep2);
    vcache[SELF_PAIR_SLOT] = cp;
  }
  return sep;
}

// ------------------------------------------------------------------------------------------
// dense 6x6 LDL^T on a packed lower triangle (A = L D L^T, unit L), and solve
// ------------------------------------------------------------------------------------------
// solve (A) x = b in place for a packed lower triangle A (A = L D L^T, unit L)
template <typename T> SO101_DEV void ldl6_factor_solve(const T (&A)[21], T (&x)[NV]) {
  T L[21], Dinv[NV], D[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) {
    T W[NV];
#pragma unroll
    for (int k = 0; k < j; k++) W[k] = L[tri(j, k)] * D[k];
    T d = A[tri(j, j)];
#pragma unroll
    for (int k = 0; k < j; k++) d -= L[tri(j, k)] * W[k];
    d = max_(d, T(MJ_MINVAL));
    D[j] = d;
    Dinv[j] = rcp_(d);
#pragma unroll
    for (int i = j + 1; i < NV; i++) {
      T s = A[tri(i, j)];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[tri(i, k)] * W[k];
      L[tri(i, j)] = s * Dinv[j];
    }
  }
#pragma unroll
  for (int i = 1; i < NV; i++) {
#pragma unroll
    for (int k = 0; k < i; k++) x[i] -= L[tri(i, k)] * x[k];
  }
#pragma unroll
  for (int i = 0; i < NV; i++) x[i] *= Dinv[i];
#pragma unroll
  for (int i = NV - 2; i >= 0; i--) {
#pragma unroll
    for (int k = i + 1; k < NV; k++) x[i] -= L[tri(k, i)] * x[k];
  }
}
// factor only: A + diag(dd) = L D L^T (strictly-lower L packed in 15, 1/D in 6)
template <typename T> SO101_DEV void ldl6_factor(const T (&A)[21], const T (&dd)[NV], T (&Ls)[15], T (&Dinv)[NV]) {
  T L[21], D[NV];
#pragma unroll
  for (int j = 0; j < NV; j++) {
    T W[NV];
#pragma unroll
    for (int k = 0; k < j; k++) W[k] = L[tri(j, k)] * D[k];
    T d = A[tri(j, j)] + dd[j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= L[tri(j, k)] * W[k];
    d = max_(d, T(MJ_MINVAL));
    D[j] = d;
    Dinv[j] = rcp_(d);
#pragma unroll
    for (int i = j + 1; i < NV; i++) {
      T s = A[tri(i, j)];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[tri(i, k)] * W[k];
      L[tri(i, j)] = s * Dinv[j];
    }
  }
#pragma unroll
  for (int i = 1; i < NV; i++) {
#pragma unroll
    for (int k = 0; k < i; k++) Ls[i * (i - 1) / 2 + k] = L[tri(i, k)];
  }
}
template <typename T> SO101_DEV void ldl6_solve(const T (&Ls)[15], const T (&Dinv)[NV], T (&x)[NV]) {
#pragma unroll
  for (int i = 1; i < NV; i++) {
#pragma unroll
    for (int k = 0; k < i; k++) x[i] -= Ls[i * (i - 1) / 2 + k] * x[k];
  }
#pragma unroll
  for (int i = 0; i < NV; i++) x[i] *= Dinv[i];
#pragma unroll
  for (int i = NV - 2; i >= 0; i--) {
#pragma unroll
    for (int k = i + 1; k < NV; k++) x[i] -= Ls[k * (k - 1) / 2 + i] * x[k];
  }
}
template <typename T> SO101_DEV void symv6(const T (&M)[21], const T (&x)[NV], T (&y)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; i++) {
    T s = T(0);
#pragma unroll
    for (int j = 0; j < NV; j++) s += M[i >= j ? tri(i, j) : tri(j, i)] * x[j];
    y[i] = s;
  }
}

// ------------------------------------------------------------------------------------------
// per-thread environment and solver counters
// ------------------------------------------------------------------------------------------
template <typename T>
struct Env {
  T q[NV], qd[NV], warm[NV], fa[NV];  // qpos, qvel, qacc_warmstart, qfrc_applied
  T time;
  uint32_t flags;
};
struct Counters {
  uint32_t steps, newton, lsevals, limsteps;
};

// ------------------------------------------------------------------------------------------
// smooth dynamics: qfrc_bias (RNEA, flg_acc = 0) and M (CRBA) in link-local frames.
// Optionally the world position of the observation site and the contact tripwire.
// ------------------------------------------------------------------------------------------
// sin/cos of the joint angles (relative to the reference pose the frames were folded at)
template <typename T> SO101_DEV void joint_sincos(const DevModel<T>& m, const T (&q)[NV], T (&sn)[NV], T (&cs)[NV]) {
#pragma unroll
  for (int k = 0; k < NV; k++) sincos_(q[k] - m.qpos0[k], &sn[k], &cs[k]);
}

// contact tripwire of link k, unrolled form
template <typename T>
SO101_DEV void tripwire_link(const DevModel<T>& m, int k, const T (&R)[9], T qk, T (&zw)[3], T& zo, uint32_t& flags,
                             uint32_t& hits) {
  tripwire_frame(R, m.r[k], zw, zo);
#pragma unroll
  for (int b = 0; b < TRIP_PER_LINK; b++) {
    if (m.trip_n[k] > b) tripwire_box(m, k, b, zw, zo, hits);
  }
}

template <typename T, bool WANT_M, bool WANT_BIAS = true>
SO101_DEV void smooth_dynamics(const DevModel<T>& m, const T (&q)[NV], const T (&qd)[NV], const T (&sn)[NV],
                               const T (&cs)[NV], T (&M)[21], T (&bias)[NV], bool want_site, T (&site)[3], bool trip,
                               uint32_t& flags, uint32_t& hits) {
  // ---- forward pass: velocities, accelerations, link forces --------------------------------
  T f[NV][6];
  if (WANT_BIAS) {
    T v[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    T a[6] = {T(0), T(0), T(0), m.accg[0], m.accg[1], m.accg[2]};
    T zw[3] = {T(0), T(0), T(1)};  // world z axis in the current frame (tripwire)
    T zo = T(0);                   // world height of the current frame origin
#pragma unroll
    for (int k = 0; k < NV; k++) {
      T R[9];
      make_R(m.E[k], cs[k], sn[k], R);
      rnea_link(R, m.r[k], m.I[k], qd[k], v, a, f[k]);
      if (trip) tripwire_link(m, k, R, q[k], zw, zo, flags, hits);
    }
  }

  // ---- backward pass: bias forces, composite inertias, mass matrix ---------------------------
  T fs[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};   // accumulated force in the current frame
  T Ic[10] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
  T F[NV][6];                                        // F[j] = Ic_j * S propagated to the current frame
  T p[3] = {T(0), T(0), T(0)};                       // site position in the current frame
#pragma unroll
  for (int k = NV - 1; k >= 0; k--) {
    if (WANT_BIAS) {
#pragma unroll
      for (int c = 0; c < 6; c++) fs[c] = add_(fs[c], f[k][c]);
      bias[k] = fs[2];
    }
    if (WANT_M) {
#pragma unroll
      for (int c = 0; c < 10; c++) Ic[c] = add_(Ic[c], m.I[k][c]);
      F[k][0] = Ic[4]; F[k][1] = Ic[5]; F[k][2] = Ic[2];
      F[k][3] = -Ic[7]; F[k][4] = Ic[6]; F[k][5] = T(0);
      M[tri(k, k)] = add_(Ic[2], m.armature[k]);
    }
    if (want_site && k == m.site_link) { p[0] = m.site[0]; p[1] = m.site[1]; p[2] = m.site[2]; }
    if ((k > 0 && (WANT_M || WANT_BIAS)) || want_site) {
      T R[9];
      make_R(m.E[k], cs[k], sn[k], R);
      if (want_site && k <= m.site_link) {
        T o[3];
        rot(R, p, o);
        p[0] = add_(o[0], m.r[k][0]); p[1] = add_(o[1], m.r[k][1]); p[2] = add_(o[2], m.r[k][2]);
      }
      if (k > 0) {
        T t6[6];
        if (WANT_BIAS) {
          xforce(R, m.r[k], fs, t6);
#pragma unroll
          for (int c = 0; c < 6; c++) fs[c] = t6[c];
        }
        if (WANT_M) {
#pragma unroll
          for (int j = NV - 1; j >= k; j--) {
            xforce(R, m.r[k], F[j], t6);
#pragma unroll
            for (int c = 0; c < 6; c++) F[j][c] = t6[c];
            M[tri(j, k - 1)] = t6[2];
          }
          T Ip[10] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
          xinertia_add(R, m.r[k], Ic, Ip);
#pragma unroll
          for (int c = 0; c < 10; c++) Ic[c] = Ip[c];
        }
      }
    }
  }
  if (want_site) { site[0] = p[0]; site[1] = p[1]; site[2] = p[2]; }
}

// ------------------------------------------------------------------------------------------
// Compact ("rolled") smooth dynamics.  The unrolled smooth_dynamics above is ~4 k SASS instructions (64 KB);
// a physics step then exceeds the 32 KB L1.5 instruction cache several times over and every block streams
// its code from L2 once per step.  ncu: with 128 blocks doing that `stall_no_instruction` is 2.8 per issue
// (0.6 with 16 blocks), i.e. L2 instruction fetch is the shared bottleneck of the small-batch configuration,
// and it costs the large-batch one a barrier per phase to keep warps on one fetch stream.
// These versions loop over the links with ONE copy of each link body: link constants come from the
// constant bank through uniform loads (the link index is warp-uniform), per-thread link arrays (sin/cos,
// link forces f_k) are indexed dynamically (local memory or shared memory, L1 resident).
// `st` = element stride of the per-thread arrays (1: thread-local array, 32: shared memory [k][lane]).
// ------------------------------------------------------------------------------------------
template <typename T>
SO101_DEV void joint_sincos_range(const DevModel<T>& m, const T* q, int qst, T* sn, T* cs, int st, int k0, int k1) {
#pragma unroll 1
  for (int k = k0; k < k1; k++) {
    T s_, c_;
    sincos_(q[k * qst] - m.qpos0[k], &s_, &c_);
    sn[k * st] = s_; cs[k * st] = c_;
  }
}

constexpr int RNEA_UNROLL = SO101_RNEA_UNROLL;
// qfrc_bias by RNEA (flg_acc = 0): forward pass over the links (velocity, bias acceleration, link force),
// backward pass accumulating the forces down the chain
template <typename T>
SO101_DEV void rnea_bias(const DevModel<T>& m, const T* sn, const T* cs, int st, const T* qd, int qst, T* bias) {
  T f[NV][6];
  {
    T v[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
    T a[6] = {T(0), T(0), T(0), m.accg[0], m.accg[1], m.accg[2]};
#pragma unroll (RNEA_UNROLL)
    for (int k = 0; k < NV; k++) {
      T R[9];
      make_R(m.E[k], cs[k * st], sn[k * st], R);
      rnea_link(R, m.r[k], m.I[k], qd[k * qst], v, a, f[k]);
    }
  }
  T fs[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};   // accumulated force in the current frame
#pragma unroll (RNEA_UNROLL)
  for (int k = NV - 1; k >= 0; k--) {
#pragma unroll
    for (int c = 0; c < 6; c++) fs[c] = add_(fs[c], f[k][c]);
    bias[k] = fs[2];
    if (k > 0) {
      T R[9], t6[6];
      make_R(m.E[k], cs[k * st], sn[k * st], R);
      xforce(R, m.r[k], fs, t6);
#pragma unroll
      for (int c = 0; c < 6; c++) fs[c] = t6[c];
    }
  }
}

// mass matrix by CRBA + armature, packed lower triangle written to M[tri(i,j) * mst].  The link loop is
// rolled; the column loop is unrolled under warp-uniform branches (F_j stay in registers, no work for j < k).
template <typename T>
SO101_DEV void crba_mass(const DevModel<T>& m, const T* sn, const T* cs, int st, T* M, int mst) {
  T Ic[10] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
  T F[NV][6];                                        // F[j] = Ic_j * S propagated to the current frame
#pragma unroll
  for (int j = 0; j < NV; j++) {
#pragma unroll
    for (int c = 0; c < 6; c++) F[j][c] = T(0);
  }
#pragma unroll 1
  for (int k = NV - 1; k >= 0; k--) {
#pragma unroll
    for (int c = 0; c < 10; c++) Ic[c] = add_(Ic[c], m.I[k][c]);
    M[tri(k, k) * mst] = add_(Ic[2], m.armature[k]);
    if (k > 0) {
      T R[9];
      make_R(m.E[k], cs[k * st], sn[k * st], R);
#if SO101_CRBA_BRANCHLESS
      // all five columns in one basic block (columns j < k are still zero and stay zero): 25 instead of 15 force
      // transforms, but five independent chains the scheduler can interleave
#pragma unroll
      for (int j = NV - 1; j >= 1; j--) {
        const bool init = j == k;
        F[j][0] = init ? Ic[4] : F[j][0]; F[j][1] = init ? Ic[5] : F[j][1]; F[j][2] = init ? Ic[2] : F[j][2];
        F[j][3] = init ? -Ic[7] : F[j][3]; F[j][4] = init ? Ic[6] : F[j][4]; F[j][5] = init ? T(0) : F[j][5];
        T t6[6];
        xforce(R, m.r[k], F[j], t6);
#pragma unroll
        for (int c = 0; c < 6; c++) F[j][c] = t6[c];
        if (j >= k) M[(j * (j + 1) / 2 + k - 1) * mst] = t6[2];
      }
#else
#pragma unroll
      for (int j = NV - 1; j >= 1; j--) {
        if (j >= k) {                                // warp-uniform
          if (j == k) {
            F[j][0] = Ic[4]; F[j][1] = Ic[5]; F[j][2] = Ic[2];
            F[j][3] = -Ic[7]; F[j][4] = Ic[6]; F[j][5] = T(0);
          }
          T t6[6];
          xforce(R, m.r[k], F[j], t6);
#pragma unroll
          for (int c = 0; c < 6; c++) F[j][c] = t6[c];
          M[(j * (j + 1) / 2 + k - 1) * mst] = t6[2];
        }
      }
#endif
      T Ip[10] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)};
      xinertia_add(R, m.r[k], Ic, Ip);
#pragma unroll
      for (int c = 0; c < 10; c++) Ic[c] = Ip[c];
    }
  }
}

// contact tripwire over all links (flags only)
template <typename T>
SO101_DEV void tripwire_all(const DevModel<T>& m, const T* sn, const T* cs, int st, const T* q, int qst,
                            uint32_t& flags, uint32_t& hits) {
  T zw[3] = {T(0), T(0), T(1)};  // world z axis in the current frame
  T zo = T(0);                   // world height of the current frame origin
#pragma unroll 1
  for (int k = 0; k < NV; k++) {
    T R[9];
    make_R(m.E[k], cs[k * st], sn[k * st], R);
    tripwire_frame(R, m.r[k], zw, zo);
#pragma unroll 1
    for (int b = 0; b < m.trip_n[k]; b++) tripwire_box(m, k, b, zw, zo, hits);
  }
}

// world position of the observation site from the joint sin/cos
template <typename T>
SO101_DEV void site_from_trig(const DevModel<T>& m, const T* sn, const T* cs, int st, T (&site)[3]) {
  T p[3] = {m.site[0], m.site[1], m.site[2]};
#pragma unroll 1
  for (int k = m.site_link; k >= 0; k--) {
    T R[9], o[3];
    make_R(m.E[k], cs[k * st], sn[k * st], R);
    rot(R, p, o);
    p[0] = add_(o[0], m.r[k][0]); p[1] = add_(o[1], m.r[k][1]); p[2] = add_(o[2], m.r[k][2]);
  }
  site[0] = p[0]; site[1] = p[1]; site[2] = p[2];
}

// forward kinematics of the observation site only (reset / mj_forward observations)
template <typename T> SO101_DEV void site_fk(const DevModel<T>& m, const T (&q)[NV], T (&site)[3]) {
  T p[3] = {T(0), T(0), T(0)};
#pragma unroll
  for (int k = NV - 1; k >= 0; k--) {
    if (k == m.site_link) { p[0] = m.site[0]; p[1] = m.site[1]; p[2] = m.site[2]; }
    if (k <= m.site_link) {
      T s, c, R[9], o[3];
      sincos_(q[k] - m.qpos0[k], &s, &c);
      make_R(m.E[k], c, s, R);
      rot(R, p, o);
      p[0] = add_(o[0], m.r[k][0]); p[1] = add_(o[1], m.r[k][1]); p[2] = add_(o[2], m.r[k][2]);
    }
  }
  site[0] = p[0]; site[1] = p[1]; site[2] = p[2];
}

// ------------------------------------------------------------------------------------------
// constraint rows.  Row order follows mj_makeConstraint: friction rows (dof order), then the
// active joint-limit row of each joint.  All Jacobian rows are +-e_i, so J is never formed.
// Limit rows are rare: their arrays are indexed in rolled loops (local memory, tiny code).
// ------------------------------------------------------------------------------------------
template <typename T>
struct Rows {
  T aref_f[NV];                      // friction rows: jar = a_i - aref_f
  T side[NV], aref_l[NV], D_l[NV];   // limit rows: jar = side*a_i - aref_l ; side = 0: inactive
  bool anylim;
};

// getimpedance (engine_core_constraint.c) for a limit row; solimp power 1 or 2 (checked at
// model creation: MuJoCo's default is 2)
template <typename T> SO101_DEV T limit_impedance(const T* si, T pos, T margin) {
  if (si[0] == si[1] || si[2] <= T(MJ_MINVAL)) return T(0.5) * (si[0] + si[1]);
  T x = abs_(sub_(pos, margin) / si[2]);
  if (x >= T(1)) return si[1];
  if (x <= T(0)) return si[0];
  T y;
  if (si[4] == T(1)) y = x;
  else if (x <= si[3]) y = mul_(x, x) / si[3];
  else y = sub_(T(1), mul_(sub_(T(1), x), sub_(T(1), x)) / sub_(T(1), si[3]));
  return fma_(y, sub_(si[1], si[0]), si[0]);   // rounding-explicit: evaluated by the lookout warp and by the one-warp kernels
}

// mj_constraintUpdate + Gauss term: total cost at acceleration a, qfrc_constraint and the diagonal
// that the quadratic rows add to the Hessian.  Friction row i (Huber): quadratic for |jar| < R f,
// linear outside; sign handling by copysign keeps the row at ~10 instructions.
template <typename T>
SO101_DEV T cost_update(const DevModel<T>& m, const Rows<T>& rw, const T (&a)[NV], const T (&Ma)[NV],
                        const T (&fsm)[NV], const T (&asm_)[NV], T (&qc)[NV], T (&hd)[NV]) {
  T s = T(0);
#pragma unroll
  for (int i = 0; i < NV; i++) {
    T jar = a[i] - rw.aref_f[i];
    bool lin = abs_(jar) >= m.fr_Rf[i];
    T fs = copysign_(m.fr_f[i], jar);
    T Dj = m.fr_D[i] * jar;
    if (lin) { s += fs * jar - m.fr_hRff[i]; qc[i] = -fs; hd[i] = T(0); }
    else { s += T(0.5) * Dj * jar; qc[i] = -Dj; hd[i] = m.fr_D[i]; }
  }
  if (rw.anylim) {
#pragma unroll 1
    for (int i = 0; i < NV; i++) {
      if (rw.side[i] != T(0)) {
        T jar = rw.side[i] * a[i] - rw.aref_l[i];
        if (jar < T(0)) {
          s += T(0.5) * rw.D_l[i] * jar * jar;
          qc[i] += rw.side[i] * (-rw.D_l[i] * jar);
          hd[i] += rw.D_l[i];
        }
      }
    }
  }
  T g = T(0);
#pragma unroll
  for (int i = 0; i < NV; i++) g += T(0.5) * (Ma[i] - fsm[i]) * (a[i] - asm_[i]);
  return s + g;
}

// ------------------------------------------------------------------------------------------
// Line search.  Along a search direction the solver's objective is a convex piecewise quadratic:
//   f'(alpha) = G1 + 2 G2 alpha + sum_i D_i clamp(jar_i + alpha jv_i, -R_i f_i, R_i f_i) jv_i   (friction rows)
//                               + sum_l D_l min(jar_l + alpha jv_l, 0) jv_l                      (limit rows)
// is continuous, piecewise linear and non-decreasing.  MuJoCo's PrimalSearch (restated literally
// in oracle/so101_oracle.c) locates its root with safeguarded Newton steps and bracketing until
// |f'| < tolerance * ls_tolerance * |search| / scale, i.e. it converges to the exact root; it needs
// 2 evaluations when no row changes zone and 9-12 when the root sits in a steep piece.  Here the
// root is computed directly by walking the pieces: find the next zone breakpoint beyond alpha,
// evaluate slope and curvature of that piece (at its midpoint, so zone membership is never
// decided on a boundary), stop if the piece's own root lies inside it.  One iteration per piece
// crossed (mean 1.7, at most one per breakpoint).  Against the literal search the step result differs by
// <= 3e-16 (qpos), <= 1.4e-13 (qvel), <= 3e-12 (qacc) relative with identical Newton iteration
// counts (CPU test `test_exact_line_search_is_equivalent` under tests/), below the CUDA-vs-CPU
// differences of the smooth dynamics.
// ------------------------------------------------------------------------------------------
// ctrlrange clamp of mj_fwdActuation, hoisted out of the sub-step loop: the control is constant over a control step
template <typename T> SO101_DEV void clamp_ctrl(const DevModel<T>& m, T (&ctrl)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; i++)
    if (m.ctrllim_mask >> i & 1) ctrl[i] = max_(m.ctrl_lo[i], min_(m.ctrl_hi[i], ctrl[i]));
}

// mj_passive, mj_fwdActuation, right-hand side of mj_fwdAcceleration, and the constraint rows of this step
// (mj_instantiateLimit + mj_makeImpedance + mj_referenceConstraint for the limit rows)
template <typename T>
SO101_DEV void build_rows(const DevModel<T>& m, Env<T>& e, const T (&ctrl)[NV], const T (&bias)[NV], T (&fsm)[NV],
                          Rows<T>& rw, Counters& cnt) {
  rw.anylim = false;
  uint32_t near = 0;   // joints inside the margin of one of their limits
#pragma unroll
  for (int i = 0; i < NV; i++) {
    T passive = -m.damping[i] * e.qd[i];
    if (m.any_stiffness) passive = -m.stiffness[i] * (e.q[i] - m.qspring[i]) - m.damping[i] * e.qd[i];
    const T c = ctrl[i];   // already clamped to ctrlrange (clamp_ctrl, once per control step)
    T force = m.act_gain[i] * c + m.act_b0[i] + m.act_b1[i] * (m.act_gear[i] * e.q[i]) +
              m.act_b2[i] * (m.act_gear[i] * e.qd[i]);
    if (m.frclim_mask >> i & 1) force = max_(m.frc_lo[i], min_(m.frc_hi[i], force));
    fsm[i] = passive - bias[i] + e.fa[i] + m.act_gear[i] * force;
    rw.aref_f[i] = -m.fr_B[i] * e.qd[i];
    if ((m.limited_mask >> i & 1) &&
        (e.q[i] - m.lim_lo[i] < m.lim_margin[i] || m.lim_hi[i] - e.q[i] < m.lim_margin[i])) near |= 1u << i;
  }
  if (near) {   // rare: the row arrays are only ever read when rw.anylim is set
#pragma unroll 1
    for (int i = 0; i < NV; i++) {
      rw.side[i] = T(0); rw.aref_l[i] = T(0); rw.D_l[i] = T(0);
      if (!(near >> i & 1)) continue;
      T dlo = e.q[i] - m.lim_lo[i], dhi = m.lim_hi[i] - e.q[i];
      bool lo = dlo < m.lim_margin[i];
      T side = lo ? T(1) : T(-1), pos = lo ? dlo : dhi;
      T imp = limit_impedance(m.lim_imp[i], pos, m.lim_margin[i]);
      T R = max_(T(MJ_MINVAL), (T(1) - imp) * m.lim_invw[i] / imp);
      rw.side[i] = side;
      rw.D_l[i] = T(1) / R;
      rw.aref_l[i] = -m.lim_B[i] * (side * e.qd[i]) - m.lim_K[i] * imp * (pos - m.lim_margin[i]);
    }
    rw.anylim = true;
    e.flags |= SO101_FLAG_LIMIT;
    cnt.limsteps++;
  }
}

template <typename T> SO101_DEV T inf_();
template <> SO101_DEV double inf_<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <> SO101_DEV float inf_<float>() { return __int_as_float(0x7f800000); }

// Returns alpha; Mv = M * sr out; exact = the step ends inside the piece it started in (see newton_exact_finish).
template <typename T>
SO101_DEV T line_search(const DevModel<T>& m, const Rows<T>& rw, const T (&Mm)[21], const T (&a)[NV],
                        const T (&Ma)[NV], const T (&fsm)[NV], const T (&sr)[NV], T (&Mv)[NV],
                        uint32_t& nev_total, bool& exact) {
  exact = false;
  T ss = T(0);
#pragma unroll
  for (int i = 0; i < NV; i++) ss += sr[i] * sr[i];
  if (sqrt_(ss) < T(MJ_MINVAL)) return T(0);
  symv6(Mm, sr, Mv);
  T G1, G2;
  {
    T g1a = T(0), g1b = T(0), g2 = T(0);
#pragma unroll
    for (int i = 0; i < NV; i++) { g1a += sr[i] * Ma[i]; g1b += fsm[i] * sr[i]; g2 += sr[i] * Mv[i]; }
    G1 = g1a - g1b;
    G2 = T(0.5) * g2;
  }
  T jar0[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) jar0[i] = a[i] - rw.aref_f[i];
  // Fast path (99 % of the searches after a prox start): if the root of the FIRST piece - the zones the rows
  // are in at alpha = 0 - is reached before any row changes zone, it is the answer.  Zones are intervals and
  // jar moves monotonically with alpha, so "same zone at alpha = root" means no breakpoint in between; a row
  // sitting exactly on a zone boundary takes the general path.
  if (!rw.anylim) {
    T d0 = G1, d1 = T(2) * G2;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const T Ds = m.fr_D[i] * sr[i];
      if (abs_(jar0[i]) < m.fr_Rf[i]) { d0 += Ds * jar0[i]; d1 += Ds * sr[i]; }
      else d0 += copysign_(m.fr_f[i], jar0[i]) * sr[i];
    }
    if (d1 <= T(0)) d1 = T(MJ_MINVAL);
    const T root = -d0 * rcp_(d1);
    bool same = root > T(0);
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const T x0 = jar0[i], x1 = jar0[i] + root * sr[i];
      const bool q0 = abs_(x0) < m.fr_Rf[i], q1 = abs_(x1) < m.fr_Rf[i];
      const bool lin_same = abs_(x0) > m.fr_Rf[i] && abs_(x1) >= m.fr_Rf[i] && (x0 > T(0)) == (x1 > T(0));
      same &= (q0 && q1) || (!q0 && lin_same);
    }
    if (same) { nev_total += 1; exact = true; return root; }
  }
  // friction row i changes zone where jar0_i + alpha sr_i = -+ R_i f_i
  T blo[NV], bhi[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const bool moving = sr[i] != T(0);
    const T inv = rcp_(moving ? sr[i] : T(1));
    blo[i] = moving ? (-m.fr_Rf[i] - jar0[i]) * inv : inf_<T>();
    bhi[i] = moving ? (m.fr_Rf[i] - jar0[i]) * inv : inf_<T>();
  }
  T alpha = T(0), result = T(0);
  uint32_t nev = 0;
#pragma unroll 1
  for (int it = 0; it < 2 * NV + NV + 2; it++) {   // one piece per iteration, <= 2 NV + NV breakpoints
    T nb = inf_<T>();
#pragma unroll
    for (int i = 0; i < NV; i++) {
      if (blo[i] > alpha) nb = min_(nb, blo[i]);
      if (bhi[i] > alpha) nb = min_(nb, bhi[i]);
    }
    if (rw.anylim) {
#pragma unroll 1
      for (int i = 0; i < NV; i++) {
        if (rw.side[i] != T(0)) {
          const T jv = rw.side[i] * sr[i], jl = rw.side[i] * a[i] - rw.aref_l[i];
          if (jv != T(0)) { const T bl = -jl / jv; if (bl > alpha) nb = min_(nb, bl); }
        }
      }
    }
    const bool last = !(nb < inf_<T>());
    const T mid = last ? alpha + T(1) : T(0.5) * (alpha + nb);
    T d0 = G1 + T(2) * G2 * mid, d1 = T(2) * G2;
#pragma unroll
    for (int i = 0; i < NV; i++) {
      const T x = jar0[i] + mid * sr[i];
      const T Ds = m.fr_D[i] * sr[i];
      if (abs_(x) < m.fr_Rf[i]) { d0 += Ds * x; d1 += Ds * sr[i]; }
      else d0 += copysign_(m.fr_f[i], x) * sr[i];
    }
    if (rw.anylim) {
#pragma unroll 1
      for (int i = 0; i < NV; i++) {
        if (rw.side[i] != T(0)) {
          const T jv = rw.side[i] * sr[i], x = rw.side[i] * a[i] - rw.aref_l[i] + mid * jv;
          if (x < T(0)) { d0 += rw.D_l[i] * x * jv; d1 += rw.D_l[i] * jv * jv; }
        }
      }
    }
    nev++;
    if (d1 <= T(0)) d1 = T(MJ_MINVAL);
    const T root = mid - d0 * rcp_(d1);
    if (root <= nb || last) { result = max_(root, alpha); break; }
    alpha = nb;
    result = nb;
  }
  nev_total += nev;
  return result;
}

// ------------------------------------------------------------------------------------------
// Direct active-set solve (friction rows only).  The solver's objective is strictly convex and piecewise quadratic:
// on the piece where row i is in zone z_i (0: quadratic, |jar| < R f; +-1: saturated at -+f) its minimiser solves
//   (M + diag(D_i [z_i = 0])) a = qfrc_smooth + D_i aref_i [z_i = 0] - z_i f_i [z_i != 0],
// and a point that solves this AND lies in the piece it assumed is the global minimiser (KKT).  The zones are guessed
// from the per-dof problem 0.5 M_ii (a - qacc_smooth_i)^2 + huber_i(a - aref_i), whose minimiser is in the
// quadratic zone iff |M_ii (qacc_smooth_i - aref_i)| < R f M_ii + f and saturated with the sign of that quantity
// otherwise (no division).  M is armature dominated: the guess is right in 99.8 % of the steps; then ONE factorisation
// gives MuJoCo's converged Newton iterate without cost evaluations, gradients or a line search.  A rejected candidate
// gets one more attempt with the zones it lies in (99.97 % accepted after that; a warp of 32 envs: 93 % -> 99 %);
// what is left takes the general Newton path below from the prox point.
// ------------------------------------------------------------------------------------------
template <typename T>
SO101_DEV void active_set_guess(const DevModel<T>& m, const Rows<T>& rw, const T (&M)[21], const T (&asm_)[NV],
                                T (&zone)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const T Mii = M[tri(i, i)];
    const T t = Mii * (asm_[i] - rw.aref_f[i]);
    const bool quad = abs_(t) < m.fr_Rf[i] * Mii + m.fr_f[i];
    zone[i] = quad ? T(0) : copysign_(T(1), t);
  }
}
// second attempt: the zones the rejected candidate lies in (one step of an active-set iteration)
template <typename T>
SO101_DEV void active_set_zones_at(const DevModel<T>& m, const Rows<T>& rw, const T (&a)[NV], T (&zone)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const T jar = a[i] - rw.aref_f[i];
    zone[i] = abs_(jar) < m.fr_Rf[i] ? T(0) : copysign_(T(1), jar);
  }
}
// right-hand side and Hessian diagonal of the piece given by the zones
template <typename T>
SO101_DEV void active_set_system(const DevModel<T>& m, const Rows<T>& rw, const T (&zone)[NV], const T (&fsm)[NV],
                                 T (&rhs)[NV], T (&hd)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const bool quad = zone[i] == T(0);
    hd[i] = quad ? m.fr_D[i] : T(0);
    rhs[i] = fsm[i] + (quad ? m.fr_D[i] * rw.aref_f[i] : -zone[i] * m.fr_f[i]);
  }
}
// KKT check of the direct solve at a, and the constraint force there (same values as cost_update's)
template <typename T>
SO101_DEV bool active_set_accept(const DevModel<T>& m, const Rows<T>& rw, const T (&zone)[NV], const T (&a)[NV],
                                 T (&qc)[NV]) {
  bool ok = true;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    const T jar = a[i] - rw.aref_f[i];
    const bool quad = zone[i] == T(0);
    ok &= quad ? (abs_(jar) < m.fr_Rf[i]) : (zone[i] * jar > m.fr_Rf[i] || m.fr_f[i] == T(0));
    qc[i] = quad ? -m.fr_D[i] * jar : -zone[i] * m.fr_f[i];
  }
  return ok;
}

// Minimiser of the per-dof problem 0.5 M_ii (a - as)^2 + huber_i(a - ar) (the Newton start, see physics_step).
// Quadratic zone |a - ar| < R f: a = (M_ii as + D ar) / (M_ii + D), i.e. a - ar = M_ii (as - ar) / (M_ii + D), so the
// zone test needs no division; linear zones: a = as -+ f / M_ii if that lands beyond the zone, else the zone edge.
// One reciprocal per dof, no divergent branch.
template <typename T> SO101_DEV T prox_point(const DevModel<T>& m, int i, T Mii, T as, T ar) {
  const T d = as - ar;
  const bool quad = abs_(Mii * d) < m.fr_Rf[i] * (Mii + m.fr_D[i]);
  const T inv = rcp_(quad ? Mii + m.fr_D[i] : Mii);
  const T aq = (Mii * as + m.fr_D[i] * ar) * inv;
  const T df = m.fr_f[i] * inv;
  const T ap = as - df, an = as + df;
  const T edge = ar + copysign_(m.fr_Rf[i], d);
  const T al = (ap - ar >= m.fr_Rf[i]) ? ap : ((an - ar <= -m.fr_Rf[i]) ? an : edge);
  return quad ? aq : al;
}

// A Newton step whose line search ended inside the quadratic piece it started in (line_search fast path) lands on
// the exact minimiser of that piece, and the cost is convex: it is the solution.  MuJoCo would evaluate the cost once
// more, find improvement / gradient below tolerance and stop with the same iterate; here only the constraint force
// of the new point is formed (same expressions as cost_update) and the solver stops - the second cost evaluation,
// M a and the gradient norm are skipped.
template <typename T>
SO101_DEV void newton_exact_finish(const DevModel<T>& m, const Rows<T>& rw, const T (&a)[NV], T (&qc)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; i++) {
    T jar = a[i] - rw.aref_f[i];
    bool lin = abs_(jar) >= m.fr_Rf[i];
    T fs = copysign_(m.fr_f[i], jar);
    T Dj = m.fr_D[i] * jar;
    qc[i] = lin ? -fs : -Dj;
  }
}

}  // namespace so101
#include "so101_contact.cuh"
namespace so101 {

// the caller's side of the contact path.  sn / cs: joint sines / cosines (stride st); con / ncon: the contacts if the
// geometry half has already run elsewhere (the team's lookout warp), else ncon < 0 and it runs here.  Falls back to the
// out-of-line general solver (copies in, call, copies out: see ContactIO) when the active-set iteration does not settle.
template <typename T>
SO101_DEV bool contact_branch(const DevModel<T>& m, Env<T>& e, uint32_t hits, const T* sn, const T* cs, int st,
                              Con3<T>* con, int ncon, const T (&M)[21], const T (&fsm)[NV], const Rows<T>& rw,
                              const T (&asm_)[NV], T (&a)[NV], T (&qc)[NV], Counters& cnt, int32_t* vcache) {
  if (ncon < 0) {
    uint32_t fl = 0;
    ncon = contact_geometry<T>(m, sn, cs, st, e.qd, hits, con, fl, vcache);
    e.flags |= fl;
  }
  if (ncon == 0) return false;
  e.flags |= SO101_FLAG_CONTACT;
  // start of the active-set iteration: qacc_smooth (`asm_`) with the per-dof friction guess; with an active joint-limit
  // row (no per-dof rule for it) qacc_warmstart
  if ((rw.anylim ? contact_active_set<T>(m, con, ncon, e.warm, false, M, fsm, rw, a, qc, cnt)
                 : contact_active_set<T>(m, con, ncon, asm_, true, M, fsm, rw, a, qc, cnt)) == 1) return true;
  ContactIO<T> io;
#pragma unroll
  for (int i = 0; i < NV; i++) {
    io.warm[i] = e.warm[i]; io.fsm[i] = fsm[i]; io.aref_f[i] = rw.aref_f[i];
    io.lim_side[i] = rw.anylim ? rw.side[i] : T(0);
    io.lim_aref[i] = rw.anylim ? rw.aref_l[i] : T(0);
    io.lim_D[i] = rw.anylim ? rw.D_l[i] : T(0);
  }
#pragma unroll
  for (int i = 0; i < 21; i++) io.M[i] = M[i];
  io.anylim = rw.anylim ? 1u : 0u;
  io.flags = 0; io.newton = 0; io.lsevals = 0;
  contact_solve<T>(m, io, con, ncon);
  e.flags |= io.flags;
#pragma unroll
  for (int i = 0; i < NV; i++) { a[i] = io.a[i]; qc[i] = io.qc[i]; }
  cnt.newton += io.newton;
  cnt.lsevals += io.lsevals;
  return true;
}

// ------------------------------------------------------------------------------------------
// one physics step.  After the smooth dynamics every lane runs the same small phase machine
//   SMOOTH (qacc_smooth = M^-1 qfrc_smooth) -> NEWTON x n -> EULER ((M + hB)^-1 ...)
// whose three phases share ONE factor-and-solve of M + diag(dd): the code of the 6x6 LDL^T
// exists once, and lanes in different phases execute it together.  SYNC: the block re-converges
// on a barrier at the phase boundaries so that its warps stream the same instructions through
// the instruction cache together (the step is ~50 KB of SASS; ncu: stall_no_instruction).
// ------------------------------------------------------------------------------------------
// rare: box-box test of the non-adjacent links, right after the joint sines / cosines (little else is live yet).  The budget
// (float bits in the env's hull cache word SELF_BUDGET_SLOT) is spent by every step, whether the pose is inside the fast-accept
// box or not, so it is always a valid bound for the pose it is compared at.
// -> bit 0: the critical pair's budget is gone, bit 1: the budget of the other pairs is gone
template <typename T> SO101_DEV int self_budget_spend(const DevModel<T>& m, const T (&qd)[NV], int32_t* vcache) {
  float mv = 0.f;
#pragma unroll
  for (int k = 1; k < NV; k++) mv += fabsf((float)qd[k]);
  mv *= 1.001f * (float)m.h;
  const float b = __int_as_float(vcache[SELF_BUDGET_SLOT]) - mv, r = __int_as_float(vcache[SELF_REST_SLOT]) - mv;
  vcache[SELF_BUDGET_SLOT] = __float_as_int(b);
  vcache[SELF_REST_SLOT] = __float_as_int(r);
  return (b <= 0.f ? 1 : 0) | (r <= 0.f ? 2 : 0);
}
// When one env of a warp needs the full test, every env of the warp that is outside the fast-accept box takes it with it and
// refreshes its own budgets: the lanes pay for the code together.
#define SO101_SELF_TEST_ONEWARP                                                                                            \
  if (self_cand && (self_full || self_pair)) {                                                                             \
    const double sep_ = !m.self_boxes ? -1.0 :                                                                             \
                        self_boxes_overlap<T>(m, sn[0], sn[1], sn[2], sn[3], sn[4], sn[5], cs[0], cs[1], cs[2], cs[3],     \
                                              cs[4], cs[5], vcache, self_full ? -1 : vcache[SELF_PAIR_SLOT]);              \
    if (sep_ < 0.0) e.flags |= SO101_FLAG_TRIP_SELF;                                                                       \
  }
template <typename T, bool SYNC>
SO101_DEV void physics_step(const DevModel<T>& m, Env<T>& e, const T (&ctrl)[NV], bool gravcomp_capture,
                            bool want_site, T (&site)[3], bool trip, Counters& cnt, int32_t* vcache) {
  bool self_cand = false, self_full = false, self_pair = false;
  if (trip && vcache) {
    const int gone = self_budget_spend(m, e.qd, vcache);
    self_cand = !(e.flags & SO101_FLAG_TRIP_SELF) && outside_self_box(m, e.q);
    self_full = __any_sync(__activemask(), self_cand && (gone & 2));
    self_pair = self_cand && (gone & 1);
  }
  if (SYNC) __syncthreads();
  // mj_checkPos / mj_checkVel
  {
    bool bad = false;
#pragma unroll
    for (int i = 0; i < NV; i++) bad |= bad_(e.q[i]) | bad_(e.qd[i]);
    if (bad) {
#pragma unroll
      for (int i = 0; i < NV; i++) { e.q[i] = m.qpos0[i]; e.qd[i] = T(0); e.warm[i] = T(0); e.fa[i] = T(0); }
      e.time = T(0);
      e.flags |= SO101_FLAG_BADSTATE;
    }
  }
  T M[21], bias[NV];
  T sn[NV], cs[NV];    // joint sines / cosines (the contact path reads them again)
  uint32_t hits = 0;   // tripwire boxes below the table top
  {
#if SO101_ONEWARP_ROLLED   // experiment: the compact link loops of the team kernels in the one-warp kernels (see profiles/README.md)
  {
    T lq[NV], lqd[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) { lq[i] = e.q[i]; lqd[i] = e.qd[i]; }
    joint_sincos_range(m, lq, 1, sn, cs, 1, 0, NV);
    SO101_SELF_TEST_ONEWARP
    rnea_bias(m, sn, cs, 1, lqd, 1, bias);
    crba_mass(m, sn, cs, 1, M, 1);
    if (trip) tripwire_all(m, sn, cs, 1, lq, 1, e.flags, hits);
    if (want_site) site_from_trig(m, sn, cs, 1, site);
  }
#else
  {
    joint_sincos(m, e.q, sn, cs);
    SO101_SELF_TEST_ONEWARP
    smooth_dynamics<T, true>(m, e.q, e.qd, sn, cs, M, bias, want_site, site, trip, e.flags, hits);
  }
#endif
  }
  if (SYNC) __syncthreads();
  if (gravcomp_capture) {
#pragma unroll
    for (int i = 0; i < NV; i++) e.fa[i] = bias[i];
  }
  T fsm[NV], x[NV], dd[NV];
  Rows<T> rw;
  build_rows(m, e, ctrl, bias, fsm, rw, cnt);
#pragma unroll
  for (int i = 0; i < NV; i++) { x[i] = fsm[i]; dd[i] = T(0); }
  const bool constrained = m.nfriction != 0 || rw.anylim;

  enum { PH_SMOOTH, PH_DIRECT, PH_NEWTON, PH_POLISH, PH_EULER, PH_DONE };
  int phase = PH_SMOOTH, iter = 0;
  T asm_[NV], a[NV], Ma[NV], qc[NV], hd[NV], sr[NV], zone[NV];
  T cost = T(0);
  bool need_setup = false;
  while (phase != PH_DONE) {
    {
      T A[21];
#pragma unroll
      for (int i = 0; i < 21; i++) A[i] = M[i];
#pragma unroll
      for (int i = 0; i < NV; i++) A[tri(i, i)] += dd[i];
      ldl6_factor_solve(A, x);
    }
    bool to_euler = false;
    const int ph = phase;   // the phase whose solve has just been done
    if (ph == PH_SMOOTH) {
#pragma unroll
      for (int i = 0; i < NV; i++) { asm_[i] = x[i]; a[i] = x[i]; qc[i] = T(0); hd[i] = T(0); }
      bool in_contact = false;
      if (hits) {
        // a collision box is below the table top: exact hull test and, if a hull does touch, the full constraint solve
        // with contact rows (so101_contact.cuh); the phase machine then only runs its Euler solve
        if (m.con_enabled) {
          Con3<T> con[MAXCON];
          in_contact = contact_branch<T>(m, e, hits, sn, cs, 1, con, -1, M, fsm, rw, asm_, a, qc, cnt, vcache);
        } else {
          e.flags |= SO101_FLAG_TRIP_TABLE;
        }
      }
      if (in_contact) {
        to_euler = true;
      } else if (!constrained) {
        to_euler = true;  // nefc == 0: qacc = qacc_smooth
      } else if (!rw.anylim) {
        active_set_guess(m, rw, M, asm_, zone);
        active_set_system(m, rw, zone, fsm, x, dd);   // next pass of the loop solves (M + diag(dd)) x = rhs
        phase = PH_DIRECT;
      } else {
        need_setup = true;
      }
    } else if (ph == PH_DIRECT) {
      iter++;
      if (active_set_accept(m, rw, zone, x, qc)) {
#pragma unroll
        for (int i = 0; i < NV; i++) a[i] = x[i];
        cnt.newton += iter;
        to_euler = true;
      } else if (iter == 1) {
        active_set_zones_at(m, rw, x, zone);
        active_set_system(m, rw, zone, fsm, x, dd);
      } else {
        iter = 0;
        need_setup = true;
      }
    }
    if (need_setup) {
      need_setup = false;
      {
        // Starting point.  MuJoCo starts Newton from the cheaper of qacc_warmstart and qacc_smooth and then
        // needs 3 iterations in ~43 % of this scene's steps.  The cost is strictly convex (unique minimiser), so
        // the start only decides how many iterations are needed: start from the closed-form minimiser of the
        // per-dof problem 0.5 M_ii (a - qacc_smooth_i)^2 + huber_i(a - aref_i).  M is armature dominated, this
        // identifies the active set of the coupled problem in > 99.5 % of the steps and the first Newton step
        // is exact (CPU test `test_prox_start_is_equivalent`: same result to ~1e-13, 1.0 instead of 1.9-2.3
        // iterations; warp-level maximum 1.06 instead of 3.1).  qacc_warmstart is still written every step
        // (it is part of the state the API exposes) and read only when a joint-limit row is active.
        if (!rw.anylim) {
#pragma unroll
          for (int i = 0; i < NV; i++) {
            a[i] = prox_point(m, i, M[tri(i, i)], asm_[i], rw.aref_f[i]);
          }
          symv6(M, a, Ma);
          cost = cost_update(m, rw, a, Ma, fsm, asm_, qc, hd);
        } else {
          // A joint-limit row is active (rare): stiff rows make MuJoCo stop at its tolerance rather than at
          // rounding, so follow its own warm start here - the cheaper of qacc_warmstart and qacc_smooth - to
          // stay on its iterates.
          T cost0 = T(0);
#pragma unroll 1
          for (int c = 0; c < 2; c++) {
            T ca[NV], cMa[NV], cqc[NV], chd[NV];
#pragma unroll
            for (int i = 0; i < NV; i++) ca[i] = c ? asm_[i] : e.warm[i];
            symv6(M, ca, cMa);
            T cc = cost_update(m, rw, ca, cMa, fsm, asm_, cqc, chd);
            if (c == 0 || cost0 > cc) {
#pragma unroll
              for (int i = 0; i < NV; i++) { a[i] = ca[i]; Ma[i] = cMa[i]; qc[i] = cqc[i]; hd[i] = chd[i]; }
              cost = cc;
            }
            if (c == 0) cost0 = cc;
          }
        }
#pragma unroll
        for (int i = 0; i < NV; i++) { x[i] = Ma[i] - fsm[i] - qc[i]; dd[i] = hd[i]; }
        phase = PH_NEWTON;
      }
    } else if (ph == PH_NEWTON) {
#pragma unroll
      for (int i = 0; i < NV; i++) sr[i] = -x[i];
      bool stop = iter >= m.iterations;
      if (!stop) {
        T Mv[NV];
        bool exact;
        T alpha = line_search(m, rw, M, a, Ma, fsm, sr, Mv, cnt.lsevals, exact);
        if (alpha == T(0)) {
          stop = true;
        } else if (exact) {
#pragma unroll
          for (int i = 0; i < NV; i++) a[i] += alpha * sr[i];
          newton_exact_finish(m, rw, a, qc);
          iter++;
          stop = true;
        } else {
#pragma unroll
          for (int i = 0; i < NV; i++) { a[i] += alpha * sr[i]; Ma[i] += alpha * Mv[i]; }
          T oldcost = cost;
          cost = cost_update(m, rw, a, Ma, fsm, asm_, qc, hd);
          T gg = T(0), nn = T(0);
#pragma unroll
          for (int i = 0; i < NV; i++) {
            x[i] = Ma[i] - fsm[i] - qc[i];
            dd[i] = hd[i];
            gg += x[i] * x[i];
            if (Noise<T>::on) nn += fsm[i] * fsm[i] + qc[i] * qc[i];
          }
          T improvement = m.scale * (oldcost - cost);
          T gradnorm = m.scale * sqrt_(gg);
          T tol_i = m.tolerance, tol_g = m.tolerance;
          if (Noise<T>::on) {
            tol_i += T(Noise<T>::eps) * m.scale * abs_(cost);
            tol_g += T(Noise<T>::eps) * m.scale * sqrt_(nn);
          }
          iter++;
          if (improvement < tol_i || gradnorm < tol_g) stop = true;
        }
      }
      if (stop) {
        cnt.newton += iter;
        if (iter >= m.iterations) e.flags |= SO101_FLAG_MAXITER;
        if (!rw.anylim) {
          // Polish: Newton stopped on its tolerance; solve the piece its iterate lies in exactly, so that a step
          // that needed the general path ends on the same bits as one the direct solve accepted - whichever zone
          // guess a kernel family used (the team kernels guess with a lagged factor of M).
          active_set_zones_at(m, rw, a, zone);
          active_set_system(m, rw, zone, fsm, x, dd);
          phase = PH_POLISH;
        } else {
          to_euler = true;
        }
      }
    } else if (ph == PH_POLISH) {
      T pq[NV];
      if (active_set_accept(m, rw, zone, x, pq)) {
#pragma unroll
        for (int i = 0; i < NV; i++) { a[i] = x[i]; qc[i] = pq[i]; }
      }
      to_euler = true;
    } else if (ph == PH_EULER) {  // x = (M + h B)^-1 (qfrc_smooth + qfrc_constraint)
#pragma unroll
      for (int i = 0; i < NV; i++) {
        e.qd[i] += m.h * x[i];
        e.q[i] += m.h * e.qd[i];
        e.warm[i] = a[i];
      }
      e.time += m.h;
      phase = PH_DONE;
    }
    if (to_euler) {
      // mj_checkAcc
      bool bad = false;
#pragma unroll
      for (int i = 0; i < NV; i++) bad |= bad_(a[i]);
      if (bad) {
#pragma unroll
        for (int i = 0; i < NV; i++) { e.q[i] = m.qpos0[i]; e.qd[i] = T(0); e.warm[i] = T(0); e.fa[i] = T(0); }
        e.time = T(0);
        e.flags |= SO101_FLAG_BADSTATE;
        phase = PH_DONE;
      } else if (m.any_damping) {
        // mj_Euler, implicit in joint damping
#pragma unroll
        for (int i = 0; i < NV; i++) { dd[i] = m.h * m.damping[i]; x[i] = fsm[i] + qc[i]; }
        phase = PH_EULER;
      } else {
#pragma unroll
        for (int i = 0; i < NV; i++) {
          e.qd[i] += m.h * a[i];
          e.q[i] += m.h * e.qd[i];
          e.warm[i] = a[i];
        }
        e.time += m.h;
        phase = PH_DONE;
      }
    }
  }
  cnt.steps++;
}


// ------------------------------------------------------------------------------------------
// Small batches: one physics step shared by a TEAM of three warps working on the same 32 envs.
// A lone warp per SM is latency bound (ncu: 25 % of the issue slots of its scheduler, 4.4 cycles per
// instruction) and a step is one serial chain of ~5.6 k instructions.  The chain has parallel parts that do
// not need each other until the constraint solve:
//   warp 0 "dynamics" : sincos(q0,q1) | RNEA bias forces, actuation, constraint rows | Newton solve, Euler step
//   warp 1 "geometry" : sincos(q2,q3) | CRBA mass matrix, LDL^T of M (qacc_smooth) and of M + h B (implicit Euler)
//   warp 2 "lookout"  : sincos(q4,q5) | contact tripwire, observation site
// Exchange through shared memory at four block barriers per step: (S) sin/cos, (A) M + its factors + flags + site,
// (E) factors of M + h B, (B) the new (qpos, qvel).  Every value is produced by the same expression as in physics_step, so the result is
// bit-identical to the one-warp path (GPU test `test_split_team_is_bitwise_identical`): which kernel a batch size
// selects never changes a trajectory.
// ------------------------------------------------------------------------------------------
#ifdef SO101_TIMING   // debug build: where the dynamics warp of a team spends its cycles, by kind of step (tools/team_timing.py)
__device__ unsigned long long g_timing[16];   // [0..14]: dynamics warp by kind of step; [15]: unused
__device__ unsigned long long g_timing_helpers[8];
__device__ unsigned long long g_timing_con[4];   // geometry: steps, cycles to (A), (A)..(E); lookout: steps, cycles to (A)
#define SO101_TICK(var) const long long var = clock64()
#else
#define SO101_TICK(var)
#endif
constexpr int TEAM_WARPS = 3;
template <typename T>
struct SplitXch {      // shared memory of one team, structure-of-arrays over the 32 lanes
  T sn[NV][32], cs[NV][32];
  T M[21][32];
  T L1[2][15][32], D1inv[2][NV][32];   // M = L1 D1 L1^T, double buffered: step n reads the factor of M_(n-1) (see split_geometry_step)
  T L2[15][32], D2inv[NV][32];   // M + h B = L2 D2 L2^T
  T site[3][32];
  uint32_t trip[32], hits[32];   // TRIP_SELF flag / tripwire boxes below the table top (lookout warp)
  // table contact: the lookout warp does the exact hull test of every tripped box beside the dynamics warp's RNEA and hands
  // over the hulls that touch (link, support vertex, distance, height of the lowest point); the dynamics warp forms the
  // contact rows itself, inside its contact branch.  (Round 2, late: when the lookout warp formed the rows too, its rare
  // code ran just before barrier (A) and the dynamics warp paid ~5.5 k cycles of instruction-cache misses right after it,
  // profiles/r2_team_timing.txt.)
  uint32_t ncon[32];             // number of hulls that touch
  int32_t hit_k[SO101_MAXTRIP][32], hit_v[SO101_MAXTRIP][32];
  T hit_dist[SO101_MAXTRIP][32], hit_z[SO101_MAXTRIP][32];
  T q[NV][32], qd[NV][32];       // state handed back by the dynamics warp
};

// every role takes the same mj_checkPos / mj_checkVel decision on its own copy of (qpos, qvel)
template <typename T> SO101_DEV bool team_check_state(const DevModel<T>& m, T (&q)[NV], T (&qd)[NV]) {
  bool bad = false;
#pragma unroll
  for (int i = 0; i < NV; i++) bad |= bad_(q[i]) | bad_(qd[i]);
  if (bad) {
#pragma unroll
    for (int i = 0; i < NV; i++) { q[i] = m.qpos0[i]; qd[i] = T(0); }
  }
  return bad;
}
// sin/cos of the two joints of this role -> shared, barrier (S)
template <typename T>
SO101_DEV void team_sincos(const DevModel<T>& m, SplitXch<T>& x, int lane, int role, const T (&q)[NV]) {
#pragma unroll
  for (int r = 0; r < TEAM_WARPS; r++) {
    if (r == role) {
#pragma unroll
      for (int k = 2 * r; k < 2 * r + 2; k++) {
        T s_, c_;
        sincos_(q[k] - m.qpos0[k], &s_, &c_);
        x.sn[k][lane] = s_; x.cs[k][lane] = c_;
      }
    }
  }
  __syncthreads();   // (S)
}

// f64 teams guess the zones with the factor of the previous step's M (off the critical path).  In f32 a row lands
// within rounding of a zone boundary often enough (a few times in 10^7 row-steps) that a different guess can end on a
// neighbouring piece and on different bits than the one-warp kernels: f32 teams keep the exact, up-front factor.
template <typename T> struct LaggedGuess { static constexpr bool value = true; };
template <> struct LaggedGuess<float> { static constexpr bool value = false; };

// geometry warp: M(q) and the two factorisations -> shared memory
template <typename T>
SO101_DEV void split_geometry_step(const DevModel<T>& m, SplitXch<T>& x, int lane, T (&q)[NV], T (&qd)[NV],
                                    int64_t n) {
  SO101_TICK(tg0);
  team_check_state(m, q, qd);
  team_sincos(m, x, lane, 1, q);
  crba_mass(m, &x.sn[0][lane], &x.cs[0][lane], 32, &x.M[0][lane], 32);
  T M[21], zero[NV], hB[NV], Ls[15], Dinv[NV];
#pragma unroll
  for (int i = 0; i < NV; i++) { zero[i] = T(0); hB[i] = m.h * m.damping[i]; }
#pragma unroll
  for (int i = 0; i < 21; i++) M[i] = x.M[i][lane];
  // The factor of M only feeds the dynamics warp's zone guess (active_set_guess), for which the factor of the
  // PREVIOUS step's M is as good (M moves by ~1e-3 per step): it is computed after barrier (A), off the path the
  // dynamics warp waits on, into the buffer the next step reads.  Only the first step of a launch factors up front.
  if (n == 0 || !LaggedGuess<T>::value) {
    ldl6_factor(M, zero, Ls, Dinv);
#pragma unroll
    for (int i = 0; i < 15; i++) x.L1[0][i][lane] = Ls[i];
#pragma unroll
    for (int i = 0; i < NV; i++) x.D1inv[0][i][lane] = Dinv[i];
  }
#ifdef SO101_TIMING
  const long long tg1 = clock64();
#endif
  __syncthreads();   // (A) M published
#ifdef SO101_TIMING
  const long long tg2 = clock64();
#endif
  if (LaggedGuess<T>::value) {
    const int nb = (int)((n + 1) & 1);
    ldl6_factor(M, zero, Ls, Dinv);
#pragma unroll
    for (int i = 0; i < 15; i++) x.L1[nb][i][lane] = Ls[i];
#pragma unroll
    for (int i = 0; i < NV; i++) x.D1inv[nb][i][lane] = Dinv[i];
  }
  ldl6_factor(M, hB, Ls, Dinv);      // needed only by the Euler step: overlaps the constraint solve
#pragma unroll
  for (int i = 0; i < 15; i++) x.L2[i][lane] = Ls[i];
#pragma unroll
  for (int i = 0; i < NV; i++) x.D2inv[i][lane] = Dinv[i];
#ifdef SO101_TIMING
  if (lane == 0) {
    atomicAdd(&g_timing_helpers[0], 1ull);
    atomicAdd(&g_timing_helpers[1], (unsigned long long)(tg1 - tg0));
    atomicAdd(&g_timing_helpers[2], (unsigned long long)(clock64() - tg2));
    atomicAdd(&g_timing_helpers[3], (unsigned long long)(tg2 - tg1));
  }
#endif
  __syncthreads();   // (E) factors of M + h B published
  __syncthreads();   // (B) new state published by the dynamics warp
#pragma unroll
  for (int i = 0; i < NV; i++) { q[i] = x.q[i][lane]; qd[i] = x.qd[i][lane]; }
}

// contact tripwire over all links with the exact hull test of every box that trips done on the spot, from the same
// (zw, zo): one walk down the chain serves both (team lookout warp)
template <typename T>
SO101_DEV void tripwire_all_tests(const DevModel<T>& m, const T* sn, const T* cs, int st, const T* q, int qst,
                                  uint32_t& flags, uint32_t& hits, int32_t* vcache, HitList<T>& hl) {
  T zw[3] = {T(0), T(0), T(1)};  // world z axis in the current frame
  T zo = T(0);                   // world height of the current frame origin
#pragma unroll 1
  for (int k = 0; k < NV; k++) {
    T R[9];
    make_R(m.E[k], cs[k * st], sn[k * st], R);
    tripwire_frame(R, m.r[k], zw, zo);
#pragma unroll 1
    for (int b = 0; b < m.trip_n[k]; b++) {
      const uint32_t before = hits;
      tripwire_box(m, k, b, zw, zo, hits);
      if (hits != before) hull_test(m, k, b, zw, zo, vcache, hl);
    }
  }
}

// lookout warp: contact tripwire and (on the last substep of a control step) the observation site
template <typename T>
SO101_DEV void split_lookout_step(const DevModel<T>& m, SplitXch<T>& x, int lane, T (&q)[NV], T (&qd)[NV],
                                  bool want_site, bool trip, int32_t* vcache) {
  SO101_TICK(tl0);
  team_check_state(m, q, qd);
  team_sincos(m, x, lane, 2, q);
  uint32_t fl = 0, hits = 0, nc = 0;
  if (trip) {
    T lq[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) lq[i] = q[i];
    if (m.con_enabled) {
      HitList<T> hl;
      hl.n = 0;
      tripwire_all_tests(m, &x.sn[0][lane], &x.cs[0][lane], 32, lq, 1, fl, hits, vcache, hl);
      nc = (uint32_t)hl.n;
#pragma unroll 1
      for (int h = 0; h < hl.n; h++) {
        x.hit_k[h][lane] = hl.k[h]; x.hit_v[h][lane] = hl.v[h]; x.hit_dist[h][lane] = hl.dist[h]; x.hit_z[h][lane] = hl.z[h];
      }
    } else {
      tripwire_all(m, &x.sn[0][lane], &x.cs[0][lane], 32, lq, 1, fl, hits);
    }
  }
  if (trip) {   // see physics_step; x.trip still holds this lane's flags of the previous step (sticky within the launch)
    const int gone = self_budget_spend(m, qd, vcache);
    const bool cand = !(x.trip[lane] & SO101_FLAG_TRIP_SELF) && outside_self_box(m, q);
    const bool full = __any_sync(__activemask(), cand && (gone & 2));        // the warp's candidates take the full test together
    if (x.trip[lane] & SO101_FLAG_TRIP_SELF) fl |= SO101_FLAG_TRIP_SELF;
    else if (cand && (full || (gone & 1))) {
      const double sep = !m.self_boxes ? -1.0 : self_boxes_overlap<T>(m, x.sn[0][lane], x.sn[1][lane], x.sn[2][lane], x.sn[3][lane], x.sn[4][lane], x.sn[5][lane],
                                               x.cs[0][lane], x.cs[1][lane], x.cs[2][lane], x.cs[3][lane], x.cs[4][lane], x.cs[5][lane],
                                               vcache, full ? -1 : vcache[SELF_PAIR_SLOT]);
      if (sep < 0.0) fl |= SO101_FLAG_TRIP_SELF;
    }
  }
  x.ncon[lane] = nc;
  x.trip[lane] = fl;
  x.hits[lane] = hits;
  if (want_site) {
    T p[3];
    site_from_trig(m, &x.sn[0][lane], &x.cs[0][lane], 32, p);
#pragma unroll
    for (int c = 0; c < 3; c++) x.site[c][lane] = p[c];
  }
#ifdef SO101_TIMING
  const long long tl1 = clock64();
#endif
  __syncthreads();   // (A)
#ifdef SO101_TIMING
  if (lane == 0) { atomicAdd(&g_timing_helpers[4], 1ull); atomicAdd(&g_timing_helpers[5], (unsigned long long)(tl1 - tl0));
                   atomicAdd(&g_timing_helpers[6], (unsigned long long)(clock64() - tl1)); }
#endif
  __syncthreads();   // (E)
  __syncthreads();   // (B)
#pragma unroll
  for (int i = 0; i < NV; i++) { q[i] = x.q[i][lane]; qd[i] = x.qd[i][lane]; }
}

// dynamics warp: everything else of the step.  nstep = index of this step within the launch (same for the whole block:
// selects the buffer of the lagged factor the geometry warp wrote).  A frozen lane (see physics_step) skips the
// arithmetic but walks through the same barriers and re-publishes its unchanged state.
template <typename T>
SO101_DEV void split_dynamics_step(const DevModel<T>& m, SplitXch<T>& x, int lane, Env<T>& e, const T (&ctrl)[NV],
                                   bool gravcomp_capture, bool want_site, T (&site)[3], bool trip, Counters& cnt,
                                   int64_t nstep) {
  SO101_TICK(tk0);
  if (team_check_state(m, e.q, e.qd)) {
#pragma unroll
    for (int i = 0; i < NV; i++) { e.warm[i] = T(0); e.fa[i] = T(0); }
    e.time = T(0);
    e.flags |= SO101_FLAG_BADSTATE;
  }
  T M[21], bias[NV];
  team_sincos(m, x, lane, 0, e.q);
  T fsm[NV], asm_[NV];
  Rows<T> rw;
  rw.anylim = false;
  bool constrained = false;
  {
    T lqd[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) lqd[i] = e.qd[i];
    rnea_bias(m, &x.sn[0][lane], &x.cs[0][lane], 32, lqd, 1, bias);
    if (gravcomp_capture) {
#pragma unroll
      for (int i = 0; i < NV; i++) e.fa[i] = bias[i];
    }
    build_rows(m, e, ctrl, bias, fsm, rw, cnt);
#pragma unroll
    for (int i = 0; i < NV; i++) asm_[i] = fsm[i];
    constrained = m.nfriction != 0 || rw.anylim;
  }
  SO101_TICK(tk1);
  __syncthreads();   // (A) wait for the geometry and lookout warps
  SO101_TICK(tk2);
  T Ls[15], Dinv[NV];
  T a[NV], Ma[NV], qc[NV], hd[NV];
  bool solved = false, had_contact = false;
  {
  if (trip) e.flags |= x.trip[lane];
  if (want_site) {
#pragma unroll
    for (int c = 0; c < 3; c++) site[c] = x.site[c][lane];
  }
#pragma unroll
  for (int i = 0; i < 21; i++) M[i] = x.M[i][lane];
  if (constrained && !rw.anylim) {
    // qacc_smooth approximated with the factor of the previous step's M (split_geometry_step): feeds the zone guess of
    // the direct solve and the first piece of the contact solve
    const int rb = LaggedGuess<T>::value ? (int)(nstep & 1) : 0;
#pragma unroll
    for (int i = 0; i < 15; i++) Ls[i] = x.L1[rb][i][lane];
#pragma unroll
    for (int i = 0; i < NV; i++) Dinv[i] = x.D1inv[rb][i][lane];
    ldl6_solve(Ls, Dinv, asm_);
  }
  // The lanes without a contact solve FIRST: their code is the hot code of every step and still sits in the instruction
  // cache; the contact lane's solver (rare, ~1400 instructions of its own) runs after them instead of displacing it before
  // they get to it (profiles/r2_team_timing.txt).
  SO101_TICK(tp0);
  const bool clane = trip && x.hits[lane] && m.con_enabled && x.ncon[lane] != 0;
  if (!clane && constrained && !rw.anylim) {   // direct active-set solve (see active_set_guess)
    T zone[NV], xs[NV], dh[NV];
    active_set_guess(m, rw, M, asm_, zone);
#pragma unroll 1
    for (int pass = 1; pass <= 2; pass++) {
      active_set_system(m, rw, zone, fsm, xs, dh);
      {
        T A[21];
#pragma unroll
        for (int i = 0; i < 21; i++) A[i] = M[i];
#pragma unroll
        for (int i = 0; i < NV; i++) A[tri(i, i)] += dh[i];
        ldl6_factor_solve(A, xs);
      }
      if (active_set_accept(m, rw, zone, xs, qc)) {
#pragma unroll
        for (int i = 0; i < NV; i++) a[i] = xs[i];
        cnt.newton += pass;
        solved = true;
        break;
      }
      active_set_zones_at(m, rw, xs, zone);
    }
  }
  __syncwarp();
  SO101_TICK(tp1);
  if (trip && x.hits[lane]) {        // table contact (see physics_step)
    if (m.con_enabled) {
      const uint32_t nc = x.ncon[lane];
      if (nc) {
        Con3<T> con[MAXCON];
        HitList<T> hl;
        hl.n = (int)nc;
#pragma unroll 1
        for (int h = 0; h < hl.n; h++) {
          hl.k[h] = x.hit_k[h][lane]; hl.v[h] = x.hit_v[h][lane]; hl.dist[h] = x.hit_dist[h][lane]; hl.z[h] = x.hit_z[h][lane];
        }
        uint32_t fl = 0;
        const int ncon = contact_rows<T>(m, &x.sn[0][lane], &x.cs[0][lane], 32, e.qd, hl, con, fl);
        e.flags |= fl;
        if (ncon)
          solved = contact_branch<T>(m, e, x.hits[lane], &x.sn[0][lane], &x.cs[0][lane], 32, con, ncon, M, fsm, rw, asm_, a, qc,
                                     cnt, nullptr);
        had_contact = solved;
      }
    } else {
      e.flags |= SO101_FLAG_TRIP_TABLE;
    }
  }
  __syncwarp();   // the lanes that took the contact branch rejoin here
#ifdef SO101_TIMING
  {
    const long long tp2 = clock64();
    if (__ballot_sync(0xffffffffu, had_contact) && lane == 0) {
      atomicAdd(&g_timing_helpers[7], 1ull);
      atomicAdd(&g_timing[15], (unsigned long long)(tp0 - tk2));          // M load + lagged qacc_smooth
      atomicAdd(&g_timing_con[0], (unsigned long long)(tp1 - tp0));       // direct solve of the lanes without contact
      atomicAdd(&g_timing_con[1], (unsigned long long)(tp2 - tp1));       // contact block
    }
  }
#endif
  if (!solved) {
    // general path (no friction rows, an active limit row, or the direct solve was rejected): exact qacc_smooth
#pragma unroll
    for (int i = 0; i < NV; i++) asm_[i] = fsm[i];
    ldl6_factor_solve(M, asm_);
#pragma unroll
    for (int i = 0; i < NV; i++) { a[i] = asm_[i]; qc[i] = T(0); hd[i] = T(0); }
  }
  if (!solved && constrained) {
    T cost;
    if (!rw.anylim) {   // prox start (see physics_step)
#pragma unroll
      for (int i = 0; i < NV; i++) {
        a[i] = prox_point(m, i, M[tri(i, i)], asm_[i], rw.aref_f[i]);
      }
      symv6(M, a, Ma);
      cost = cost_update(m, rw, a, Ma, fsm, asm_, qc, hd);
    } else {            // MuJoCo's warm start when a limit row is active
      T cost0 = T(0);
      cost = T(0);
#pragma unroll 1
      for (int c = 0; c < 2; c++) {
        T ca[NV], cMa[NV], cqc[NV], chd[NV];
#pragma unroll
        for (int i = 0; i < NV; i++) ca[i] = c ? asm_[i] : e.warm[i];
        symv6(M, ca, cMa);
        T cc = cost_update(m, rw, ca, cMa, fsm, asm_, cqc, chd);
        if (c == 0 || cost0 > cc) {
#pragma unroll
          for (int i = 0; i < NV; i++) { a[i] = ca[i]; Ma[i] = cMa[i]; qc[i] = cqc[i]; hd[i] = chd[i]; }
          cost = cc;
        }
        if (c == 0) cost0 = cc;
      }
    }
    int iter = 0;
    while (iter < m.iterations) {
      T sr[NV], Mv[NV];
#pragma unroll
      for (int i = 0; i < NV; i++) sr[i] = Ma[i] - fsm[i] - qc[i];
      {
        T A[21];
#pragma unroll
        for (int i = 0; i < 21; i++) A[i] = M[i];
#pragma unroll
        for (int i = 0; i < NV; i++) A[tri(i, i)] += hd[i];
        ldl6_factor_solve(A, sr);
      }
#pragma unroll
      for (int i = 0; i < NV; i++) sr[i] = -sr[i];
      bool exact;
      T alpha = line_search(m, rw, M, a, Ma, fsm, sr, Mv, cnt.lsevals, exact);
      if (alpha == T(0)) break;
      if (exact) {
#pragma unroll
        for (int i = 0; i < NV; i++) a[i] += alpha * sr[i];
        newton_exact_finish(m, rw, a, qc);
        iter++;
        break;
      }
#pragma unroll
      for (int i = 0; i < NV; i++) { a[i] += alpha * sr[i]; Ma[i] += alpha * Mv[i]; }
      T oldcost = cost;
      cost = cost_update(m, rw, a, Ma, fsm, asm_, qc, hd);
      T gg = T(0), nn = T(0);
#pragma unroll
      for (int i = 0; i < NV; i++) {
        T g = Ma[i] - fsm[i] - qc[i];
        gg += g * g;
        if (Noise<T>::on) nn += fsm[i] * fsm[i] + qc[i] * qc[i];
      }
      T improvement = m.scale * (oldcost - cost);
      T gradnorm = m.scale * sqrt_(gg);
      T tol_i = m.tolerance, tol_g = m.tolerance;
      if (Noise<T>::on) {
        tol_i += T(Noise<T>::eps) * m.scale * abs_(cost);
        tol_g += T(Noise<T>::eps) * m.scale * sqrt_(nn);
      }
      iter++;
      if (improvement < tol_i || gradnorm < tol_g) break;
    }
    cnt.newton += iter;
    if (iter >= m.iterations) e.flags |= SO101_FLAG_MAXITER;
    if (!rw.anylim) {   // polish (see physics_step): the exact minimiser of the piece the Newton iterate lies in
      T zone[NV], xs[NV], dh[NV], pq[NV];
      active_set_zones_at(m, rw, a, zone);
      active_set_system(m, rw, zone, fsm, xs, dh);
      {
        T A[21];
#pragma unroll
        for (int i = 0; i < 21; i++) A[i] = M[i];
#pragma unroll
        for (int i = 0; i < NV; i++) A[tri(i, i)] += dh[i];
        ldl6_factor_solve(A, xs);
      }
      if (active_set_accept(m, rw, zone, xs, pq)) {
#pragma unroll
        for (int i = 0; i < NV; i++) { a[i] = xs[i]; qc[i] = pq[i]; }
      }
    }
  }
  }
  SO101_TICK(tk3);
  __syncthreads();   // (E) factors of M + h B
  SO101_TICK(tk4);
  {
  // mj_checkAcc, mj_Euler
  bool bad = false;
#pragma unroll
  for (int i = 0; i < NV; i++) bad |= bad_(a[i]);
  if (bad) {
#pragma unroll
    for (int i = 0; i < NV; i++) { e.q[i] = m.qpos0[i]; e.qd[i] = T(0); e.warm[i] = T(0); e.fa[i] = T(0); }
    e.time = T(0);
    e.flags |= SO101_FLAG_BADSTATE;
  } else {
    T acc[NV];
    if (m.any_damping) {
#pragma unroll
      for (int i = 0; i < 15; i++) Ls[i] = x.L2[i][lane];
#pragma unroll
      for (int i = 0; i < NV; i++) { Dinv[i] = x.D2inv[i][lane]; acc[i] = fsm[i] + qc[i]; }
      ldl6_solve(Ls, Dinv, acc);
    } else {
#pragma unroll
      for (int i = 0; i < NV; i++) acc[i] = a[i];
    }
#pragma unroll
    for (int i = 0; i < NV; i++) {
      e.qd[i] += m.h * acc[i];
      e.q[i] += m.h * e.qd[i];
      e.warm[i] = a[i];
    }
    e.time += m.h;
  }
  cnt.steps++;
  }
#pragma unroll
  for (int i = 0; i < NV; i++) { x.q[i][lane] = e.q[i]; x.qd[i][lane] = e.qd[i]; }
  __syncthreads();   // (B) new state published
#ifdef SO101_TIMING
  {
    SO101_TICK(tk5);
    const uint32_t anyhit = __ballot_sync(0xffffffffu, trip && x.hits[lane] != 0), anycon = __ballot_sync(0xffffffffu, had_contact);
    if (lane == 0) {
      const int kind = anycon ? 2 : (anyhit ? 1 : 0);     // 0: plain step, 1: a box tripped, 2: a contact was solved
      atomicAdd(&g_timing[kind * 5 + 0], 1ull);
      atomicAdd(&g_timing[kind * 5 + 1], (unsigned long long)(tk1 - tk0));   // own work before (A)
      atomicAdd(&g_timing[kind * 5 + 2], (unsigned long long)(tk2 - tk1));   // waiting at (A)
      atomicAdd(&g_timing[kind * 5 + 3], (unsigned long long)(tk3 - tk2));   // solve
      atomicAdd(&g_timing[kind * 5 + 4], (unsigned long long)(tk5 - tk3));   // (E) .. end of step
    }
  }
#endif
}

}  // namespace so101
