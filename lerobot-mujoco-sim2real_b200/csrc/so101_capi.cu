// so101_capi.cu — kernels and the C ABI of include/so101_b200.h (sm_100a).
//
// One thread owns one environment for a whole launch: its state (qpos, qvel, qacc_warmstart,
// qfrc_applied: 25 scalars) is loaded once from the structure-of-arrays buffer, stepped
// n_substeps (step) or T*frame_skip (rollout, shoot) times entirely in registers, and stored
// once.  Loads/stores are coalesced (lane i <-> env i, consecutive addresses per SoA row).
// The path is FP64/FP32-pipe bound (about 160 FLOP per byte of state traffic); tensor cores
// and TMA are deliberately unused (no dense contraction, 100 bytes of state per thread).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "so101_physics.cuh"

using namespace so101;

// ==========================================================================================
// device helpers
// ==========================================================================================
constexpr int ROW_Q = 0, ROW_QD = 6, ROW_WARM = 12, ROW_FA = 18, ROW_TIME = 24, NROWS = 25;

template <typename T>
struct StateView {
  T* base;
  uint32_t* flags;
  int64_t n;
};

// thread -> env for the stepping kernels.  Tail threads are clamped to a valid env (they shadow it and never
// store) so that every thread reaches the block barriers.
// SPLIT kernels (small batches): a 96-thread block is a TEAM of three warps on the same 32 envs, warp 0 = dynamics
// role (owns the env, loads and stores), warps 1, 2 = geometry and lookout roles (see so101_physics.cuh, SplitXch).
template <typename T> SO101_DEV int64_t env_of_pair(const StateView<T>& s, bool& active) {
  const int64_t i = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  active = threadIdx.x < 32 && i < s.n;
  return i < s.n ? i : s.n - 1;
}
template <typename T, bool SPLIT> struct XchStorage { char unused; };
template <typename T> struct XchStorage<T, true> { SplitXch<T> x; };
template <typename T> SO101_DEV SplitXch<T>& xch_of(XchStorage<T, true>& st) { return st.x; }
template <typename T> SO101_DEV SplitXch<T>& xch_of(XchStorage<T, false>& st) { return *reinterpret_cast<SplitXch<T>*>(&st); }

// helper roles of a SPLIT kernel: pick up the initial state, then shadow `nsteps` physics steps
template <typename T, int CM>
SO101_DEV void helper_role(const DevModel<T>& m, SplitXch<T>& x, int64_t nsteps, int frame_skip) {
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const bool trip = m.ntrip > 0;
  T q[NV], qd[NV];
  __syncthreads();   // (0) initial state published
#pragma unroll
  for (int k = 0; k < NV; k++) { q[k] = x.q[k][lane]; qd[k] = x.qd[k][lane]; }
  if (role == 1) {
#pragma unroll 1
    for (int64_t n = 0; n < nsteps; n++) split_geometry_step(m, x, lane, q, qd, n);
  } else {
    int ss = 0;
#pragma unroll 1
    for (int64_t n = 0; n < nsteps; n++) {
      split_lookout_step<T, CM>(m, x, lane, q, qd, ss == frame_skip - 1, trip);
      if (++ss == frame_skip) ss = 0;
    }
  }
}
template <typename T> SO101_DEV void publish_state(SplitXch<T>& x, const Env<T>& e) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NV; k++) { x.q[k][lane] = e.q[k]; x.qd[k][lane] = e.qd[k]; }
  __syncthreads();   // (0)
}
template <typename T, bool SPLIT, int CM>
SO101_DEV void step_env(const DevModel<T>& m, SplitXch<T>& x, Env<T>& e, const T (&ctrl)[NV], bool gravcomp_capture,
                        bool want_site, T (&site)[3], bool trip, Counters& cnt, int64_t nstep) {
  if (SPLIT) split_dynamics_step<T, CM>(m, x, threadIdx.x & 31, e, ctrl, gravcomp_capture, want_site, site, trip, cnt, nstep);
  else physics_step<T, true, CM>(m, e, ctrl, gravcomp_capture, want_site, site, trip, cnt);
}

// ---- freeze / resume (table contact, see physics_step) ---------------------------------------------------------------
// The fast kernels (RESUME = false) freeze an env the moment one of its collision boxes dips below the table top, note
// where it stopped (at[i] = index of the physics step to execute next) and append it to `list`; the contact kernels
// (RESUME = true: the same kernel bodies with the contact path compiled in, one thread / team lane per LIST ENTRY)
// carry the listed envs on from there.  Frozen envs keep SO101_FLAG_FROZEN in their stored flags until the next API
// call clears it, so that later time chunks of the same call leave them to the contact kernels.
struct Frz {
  int32_t* at;
  int32_t* list;
  int32_t* count;        // entries appended so far
  const int32_t* upto;   // RESUME: process entries [0, *upto)
  int32_t clear;         // fast kernels: first launch of an API call - stale FROZEN bits are dropped at load
  int32_t min_n, max_n;  // RESUME: this launch handles lists with min_n < *upto <= max_n (short lists go to the
                         // three-warp team version of the contact kernel, long ones to the one-warp version)
};
template <typename T, bool SPLIT, bool RESUME>
SO101_DEV int64_t env_slot(const StateView<T>& s, const Frz& fz, bool& active, bool& exit_block) {
  const int64_t n_eff = RESUME ? (int64_t)*fz.upto : s.n;
  const int64_t j = SPLIT ? (int64_t)blockIdx.x * 32 + (threadIdx.x & 31) : (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t first = SPLIT ? (int64_t)blockIdx.x * 32 : (int64_t)blockIdx.x * blockDim.x;
  exit_block = first >= n_eff;                       // uniform over the block
  if (RESUME) exit_block |= n_eff <= fz.min_n || n_eff > fz.max_n;
  if (exit_block) { active = false; return 0; }
  active = (!SPLIT || threadIdx.x < 32) && j < n_eff;
  const int64_t jc = j < n_eff ? j : n_eff - 1;      // tail threads shadow a valid entry and never store
  return RESUME ? (int64_t)fz.list[jc] : jc;
}
// end of a fast kernel: a newly frozen env records where it stopped and joins the list
SO101_DEV void frz_append(const Frz& fz, int64_t i, int32_t idx) {
  fz.at[i] = idx;
  fz.list[atomicAdd(fz.count, 1)] = (int32_t)i;
}
__global__ void k_snapshot(const int32_t* count, int32_t* snap) { *snap = *count; }
__global__ void k_mask_flags(uint32_t* f, int64_t n, uint32_t mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) f[i] &= mask;
}

template <typename T> SO101_DEV int64_t env_of_thread(const StateView<T>& s, bool& active) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  active = i < s.n;
  return i < s.n ? i : s.n - 1;
}

template <typename T> SO101_DEV void load_env(const StateView<T>& s, int64_t i, Env<T>& e) {
#pragma unroll
  for (int k = 0; k < NV; k++) {
    e.q[k] = s.base[(ROW_Q + k) * s.n + i];
    e.qd[k] = s.base[(ROW_QD + k) * s.n + i];
    e.warm[k] = s.base[(ROW_WARM + k) * s.n + i];
    e.fa[k] = s.base[(ROW_FA + k) * s.n + i];
  }
  e.time = s.base[ROW_TIME * s.n + i];
  e.flags = s.flags[i];
}
template <typename T> SO101_DEV void store_env(const StateView<T>& s, int64_t i, const Env<T>& e) {
#pragma unroll
  for (int k = 0; k < NV; k++) {
    s.base[(ROW_Q + k) * s.n + i] = e.q[k];
    s.base[(ROW_QD + k) * s.n + i] = e.qd[k];
    s.base[(ROW_WARM + k) * s.n + i] = e.warm[k];
    s.base[(ROW_FA + k) * s.n + i] = e.fa[k];
  }
  s.base[ROW_TIME * s.n + i] = e.time;
  s.flags[i] = e.flags;
}
template <typename T> SO101_DEV void reset_env(const DevModel<T>& m, Env<T>& e) {  // mj_resetData
#pragma unroll
  for (int k = 0; k < NV; k++) { e.q[k] = m.qpos0[k]; e.qd[k] = T(0); e.warm[k] = T(0); e.fa[k] = T(0); }
  e.time = T(0);
  e.flags = 0;
}

SO101_DEV void add_stats(unsigned long long* stats, const Counters& c) {
  // warp-reduce, one atomic per warp and counter
  uint32_t v[4] = {c.steps, c.newton, c.lsevals, c.limsteps};
  const unsigned mask = __activemask();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    unsigned long long x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(mask, x, o);
    // with a partial warp shfl_down reads inactive lanes as undefined; fall back to per-lane atomics
    if (mask == 0xffffffffu) {
      if ((threadIdx.x & 31) == 0) atomicAdd(&stats[k], x);
    } else {
      atomicAdd(&stats[k], (unsigned long long)v[k]);
    }
  }
}

// Philox4x32-10 (Salmon et al. 2011).  Stream layout specified in DESIGN.md ("control RNG"):
// key = seed, counter = (env_lo, env_hi, step, 2*stream + block); 32-bit lanes -> [0,1).
SO101_DEV void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll 1
  for (int r = 0; r < 10; r++) {
    uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
    uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
    uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
enum { STREAM_RESET = 0, STREAM_CTRL = 1, STREAM_FREQ = 2, STREAM_AMP = 3, STREAM_PHASE = 4 };
// first 5 of the 8 uniforms of (seed, env, step, stream)
__device__ __noinline__ void uniform5(uint64_t seed, int64_t env, uint32_t step, uint32_t stream, double* out) {
  uint32_t c[4] = {(uint32_t)env, (uint32_t)((uint64_t)env >> 32), step, stream * 2u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = (double)c[k] * (1.0 / 4294967296.0);
  uint32_t d[4] = {(uint32_t)env, (uint32_t)((uint64_t)env >> 32), step, stream * 2u + 1u};
  philox4x32_10(d, (uint32_t)seed, (uint32_t)(seed >> 32));
  out[4] = (double)d[0] * (1.0 / 4294967296.0);
}
// a + b*c without contraction: the control stream must be bit-identical to the CPU restatement
SO101_DEV double muladd_rn(double a, double b, double c) { return __dadd_rn(a, __dmul_rn(b, c)); }

struct CtrlGen {  // per-env generator state for SO101_CTRL_SIN / CHIRP
  double freq[5], amp[5], phase[5];
};
struct DevSpec {
  int32_t kind, t_total;
  uint64_t seed;
  int64_t env_offset;
  double amp, freq_lo, freq_hi, reset_lo, reset_hi;
  const void* u;
};

SO101_DEV void ctrl_init(const DevSpec& s, int64_t env, CtrlGen& g) {
  if (s.kind == SO101_CTRL_SIN || s.kind == SO101_CTRL_CHIRP) {
    double r[5];
    uniform5(s.seed, env, 0, STREAM_FREQ, r);
#pragma unroll
    for (int k = 0; k < 5; k++) g.freq[k] = muladd_rn(s.freq_lo, s.freq_hi - s.freq_lo, r[k]);
    uniform5(s.seed, env, 0, STREAM_AMP, r);
#pragma unroll
    for (int k = 0; k < 5; k++) g.amp[k] = muladd_rn(-s.amp, 2 * s.amp, r[k]);
    uniform5(s.seed, env, 0, STREAM_PHASE, r);
#pragma unroll
    for (int k = 0; k < 5; k++) g.phase[k] = __dmul_rn(2 * 3.14159265358979323846, r[k]);
  }
}
// u_t  [REF SOARM101_DataCollection.py:57-74 (sin/chirp), :115,132 (random)]
template <typename T>
SO101_DEV void ctrl_gen(const DevSpec& s, const CtrlGen& g, int64_t env, int64_t local, int64_t n, int t,
                        double (&u)[5]) {
  if (s.kind == SO101_CTRL_RANDOM) {
    double r[5];
    uniform5(s.seed, env, (uint32_t)t, STREAM_CTRL, r);
#pragma unroll
    for (int k = 0; k < 5; k++) u[k] = __dmul_rn(__dmul_rn(__dadd_rn(r[k], -0.5), 2.0), s.amp);
  } else if (s.kind == SO101_CTRL_TENSOR) {
    const T* ut = static_cast<const T*>(s.u);
#pragma unroll
    for (int k = 0; k < 5; k++) u[k] = (double)ut[((int64_t)t * 5 + k) * n + local];
  } else {
#pragma unroll 1
    for (int k = 0; k < 5; k++) {
      double f = g.freq[k];
      if (s.kind == SO101_CTRL_CHIRP) f = muladd_rn(g.freq[k], s.freq_hi - s.freq_lo, (double)t / (double)s.t_total);
      double arg = __dadd_rn(__dmul_rn(__dmul_rn(2 * 3.14159265358979323846, f), (double)t), g.phase[k]);
      u[k] = __dmul_rn(g.amp[k], sin(arg));
    }
  }
}

// ==========================================================================================
// kernels
// ==========================================================================================
// Launch bounds, from the ncu / timing experiments in profiles/README.md.  f64 needs all 255 registers
// (one 256-thread block = 8 warps per SM; 168- or 128-register builds lose 20-30 % to spills).  The f32
// instantiation fits 128 registers with few spills: 16 warps per SM as ONE 512-thread block, so that all
// of them share the instruction stream between the block barriers (+5-16 % for 16 warps, +17 % more
// for the single block; stall_no_instruction was the top f32 stall with two independent blocks).
#ifndef SO101_F64_THREADS
#define SO101_F64_THREADS 256
#endif
#ifndef SO101_F32_THREADS
#define SO101_F32_THREADS 512
#endif
template <typename T> struct LBThreads { static constexpr int value = SO101_F64_THREADS; };
template <> struct LBThreads<float> { static constexpr int value = SO101_F32_THREADS; };
#define SO101_LB_THREADS LBThreads<T>::value
#define SO101_LB_BLOCKS 1
// Resident teams per SM that the register allocation of the team kernels must allow.  f64: 1 (255 registers, two
// teams fit; a 168-register build is 12 % slower at 4096 envs).  f32: 3 (161 registers, no spills, as fast as the
// unconstrained 194-register build and four teams fit per SM).
template <typename T> struct TeamMinBlocks { static constexpr int value = 1; };
template <> struct TeamMinBlocks<float> { static constexpr int value = 3; };
#define SO101_TEAM_MINBLOCKS TeamMinBlocks<T>::value
#define SO101_KERNEL(T) template <typename T> __global__ void __launch_bounds__(SO101_LB_THREADS, SO101_LB_BLOCKS)
#define SO101_STEP_KERNEL(T) \
  template <typename T, bool SPLIT> __global__ void __launch_bounds__(SPLIT ? 32 * TEAM_WARPS : SO101_LB_THREADS, SPLIT ? SO101_TEAM_MINBLOCKS : SO101_LB_BLOCKS)

// reset: mj_resetData + qpos/qvel write + (observation part of) mj_forward
//   mode 0: qpos0/qvel0 [6][N] (nullable)   mode 1: qpos[0:5] ~ U(lo,hi) from Philox
SO101_KERNEL(T)
k_reset(const __grid_constant__ DevModel<T> m, StateView<T> s, const T* qpos0, const T* qvel0, int mode,
        uint64_t seed, int64_t env_offset, double lo, double hi, float* obs) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.n) return;
  Env<T> e;
  reset_env(m, e);
  if (mode == 1) {
    double r[5];
    uniform5(seed, env_offset + i, 0, STREAM_RESET, r);
#pragma unroll
    for (int k = 0; k < 5; k++) e.q[k] = (T)muladd_rn(lo, hi - lo, r[k]);
  } else {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      if (qpos0) e.q[k] = qpos0[k * s.n + i];
      if (qvel0) e.qd[k] = qvel0[k * s.n + i];
    }
  }
  store_env(s, i, e);
  if (obs) {
    T site[3];
    site_fk(m, e.q, site);
#pragma unroll
    for (int k = 0; k < 3; k++) obs[k * s.n + i] = (float)site[k];
#pragma unroll
    for (int k = 0; k < 5; k++) obs[(3 + k) * s.n + i] = (float)e.q[k];
  }
}

// mj_forward outputs the Env shims read: observation and qfrc_bias
SO101_KERNEL(T)
k_forward(const __grid_constant__ DevModel<T> m, StateView<T> s, float* obs, T* qfrc_bias) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= s.n) return;
  Env<T> e;
  load_env(s, i, e);
  T M[21], bias[NV], site[3];
  uint32_t fl = 0;
  T sn[NV], cs[NV];
  joint_sincos_range(m, e.q, 1, sn, cs, 1, 0, NV);
  rnea_bias(m, sn, cs, 1, e.qd, 1, bias);
  site_from_trig(m, sn, cs, 1, site);
  if (obs) {
#pragma unroll
    for (int k = 0; k < 3; k++) obs[k * s.n + i] = (float)site[k];
#pragma unroll
    for (int k = 0; k < 5; k++) obs[(3 + k) * s.n + i] = (float)e.q[k];
  }
  if (qfrc_bias) {
#pragma unroll
    for (int k = 0; k < NV; k++) qfrc_bias[k * s.n + i] = bias[k];
  }
}

// SOARM101Env.step: ctrl rows [n_ctrl][N] (missing rows = 0), nsub x mj_step, observation
// MODE of a stepping kernel: how table contact is handled (see struct Frz)
//   MODE_FREEZE  large batches, pass 1: no contact code in the instruction stream; envs that reach the table are frozen
//   MODE_RESUME  large batches, pass 2: one thread per LIST ENTRY, contact path compiled in
//   MODE_INLINE  small batches (team kernels): contact path compiled in, nothing is frozen.  A team kernel is as long as
//                its slowest block whatever it does, so handing its few contact envs to a second launch only adds that
//                launch's latency.
enum { MODE_FREEZE = 0, MODE_RESUME = 1, MODE_INLINE = 2 };
#define SO101_STEP_KERNEL2(T) \
  template <typename T, bool SPLIT, int MODE> __global__ void __launch_bounds__(SPLIT ? 32 * TEAM_WARPS : SO101_LB_THREADS, SPLIT ? SO101_TEAM_MINBLOCKS : SO101_LB_BLOCKS)
SO101_STEP_KERNEL2(T)
k_step(const __grid_constant__ DevModel<T> m, StateView<T> s, const T* ctrl, int n_ctrl, int nsub, float* obs,
       unsigned long long* stats, Frz fz) {
  constexpr bool RESUME = MODE == MODE_RESUME;
  constexpr int CM = MODE == MODE_FREEZE ? CM_FREEZE : CM_SOLVE;
  __shared__ XchStorage<T, SPLIT> xst;
  SplitXch<T>& xch = xch_of(xst);
  bool active, exit_block;
  const int64_t i = env_slot<T, SPLIT, RESUME>(s, fz, active, exit_block);
  if (exit_block) return;
  if (SPLIT && threadIdx.x >= 32) { helper_role<T, CM>(m, xch, nsub, nsub); return; }
  Env<T> e;
  load_env(s, i, e);
  if (!RESUME && fz.clear) e.flags &= ~SO101_FLAG_FROZEN;
  const bool was_frozen = !RESUME && (e.flags & SO101_FLAG_FROZEN);
  const int start = RESUME ? fz.at[i] : 0;     // RESUME: the sub-step at which this env froze
  if (SPLIT) publish_state(xch, e);
  T u[NV], site[3] = {T(0), T(0), T(0)};
#pragma unroll
  for (int k = 0; k < NV; k++) u[k] = (ctrl && k < n_ctrl) ? ctrl[k * s.n + i] : T(0);
  clamp_ctrl(m, u);
  Counters cnt = {0, 0, 0, 0};
  const bool trip = m.ntrip > 0;
  int frozen_idx = -1;
#pragma unroll 1
  for (int ss = 0; ss < nsub; ss++) {
    if (RESUME && ss == start) e.flags &= ~SO101_FLAG_FROZEN;
    step_env<T, SPLIT, CM>(m, xch, e, u, false, ss == nsub - 1, site, trip, cnt, ss);
    if (!RESUME && frozen_idx < 0 && !was_frozen && (e.flags & SO101_FLAG_FROZEN)) { frozen_idx = ss; if (active) frz_append(fz, i, frozen_idx); }
  }
  if (nsub == 0) site_fk(m, e.q, site);
  if (active && !was_frozen) {
    if (RESUME) { e.flags |= SO101_FLAG_FROZEN; fz.at[i] = nsub; }
    store_env(s, i, e);
    if (obs && (RESUME || frozen_idx < 0)) {
#pragma unroll
      for (int k = 0; k < 3; k++) obs[k * s.n + i] = (float)site[k];
#pragma unroll
      for (int k = 0; k < 5; k++) obs[(3 + k) * s.n + i] = (float)e.q[k];
    }
  } else {
    cnt = {0, 0, 0, 0};
  }
  add_stats(stats, cnt);
}

// SOARM101DataGenerator.generate_physics_based_data, one env per thread:
// rows[N][T+1][13] = [u_t(5) | float32(ee_pos)(3) | float32(qpos[0:5])(5)]
template <typename T, typename ROW, bool SPLIT, int MODE>
__global__ void __launch_bounds__(SPLIT ? 32 * TEAM_WARPS : SO101_LB_THREADS, SPLIT ? SO101_TEAM_MINBLOCKS : SO101_LB_BLOCKS)
k_rollout(const __grid_constant__ DevModel<T> m, StateView<T> s, DevSpec spec, int t0, int t1, int Tn, int frame_skip,
          ROW* rows, uint32_t rflags, unsigned long long* stats, Frz fz) {
  // control steps (t0, t1] of a rollout of Tn steps; t0 > 0 continues a previous launch (row t0 is already written,
  // u_t0 is regenerated: the control stream is a pure function of (seed, env, t)).  Physics step index of sub-step ss
  // of loop iteration t: t * frame_skip + ss (what Frz.at holds).
  constexpr bool RESUME = MODE == MODE_RESUME;
  constexpr int CM = MODE == MODE_FREEZE ? CM_FREEZE : CM_SOLVE;
  __shared__ XchStorage<T, SPLIT> xst;
  SplitXch<T>& xch = xch_of(xst);
  bool active, exit_block;
  const int64_t i = env_slot<T, SPLIT, RESUME>(s, fz, active, exit_block);
  if (exit_block) return;
  if (RESUME) {
    // the list is ordered by freeze time: skip the iterations before the first env of this block takes over (iteration
    // t0' regenerates the control u_t0' and nothing else, exactly like iteration t0 of any launch)
    __shared__ int s_first;
    if (threadIdx.x == 0) s_first = 0x7fffffff;
    __syncthreads();
    if (active) atomicMin(&s_first, fz.at[i]);
    __syncthreads();
    const int tb = s_first / frame_skip - 1;
    if (tb > t0) t0 = tb;
  }
  if (SPLIT && threadIdx.x >= 32) { helper_role<T, CM>(m, xch, (int64_t)(t1 - t0) * frame_skip, frame_skip); return; }
  const int64_t env = spec.env_offset + i;
  Env<T> e;
  if (RESUME || (rflags & SO101_ROLL_NO_RESET)) {
    load_env(s, i, e);
    if (!RESUME && fz.clear) e.flags &= ~SO101_FLAG_FROZEN;
  } else {
    reset_env(m, e);
    double r[5];
    uniform5(spec.seed, env, 0, STREAM_RESET, r);
#pragma unroll
    for (int k = 0; k < 5; k++) e.q[k] = (T)muladd_rn(spec.reset_lo, spec.reset_hi - spec.reset_lo, r[k]);
  }
  const bool was_frozen = !RESUME && (e.flags & SO101_FLAG_FROZEN);
  const int start = RESUME ? fz.at[i] : 0;
  CtrlGen g;
  ctrl_init(spec, env, g);
  Counters cnt = {0, 0, 0, 0};
  const bool trip = m.ntrip > 0;
  const bool hold = rflags & SO101_ROLL_GRAVCOMP_HOLD;
  T site[3];
  site_fk(m, e.q, site);
  if (SPLIT) publish_state(xch, e);
  double u[5];
  T uc[NV] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  int frozen_idx = -1;
  int64_t nstep = 0;
#pragma unroll 1
  for (int t = t0; t <= t1; t++) {
    if (t > t0) {
#pragma unroll 1
      for (int ss = 0; ss < frame_skip; ss++, nstep++) {
        if (RESUME && t * frame_skip + ss == start) e.flags &= ~SO101_FLAG_FROZEN;
        step_env<T, SPLIT, CM>(m, xch, e, uc, hold && ss == 0, ss == frame_skip - 1, site, trip, cnt, nstep);
        if (!RESUME && frozen_idx < 0 && !was_frozen && (e.flags & SO101_FLAG_FROZEN)) {
          frozen_idx = t * frame_skip + ss;      // appended now: the list ends up ordered by freeze time, so the lanes of a
          if (active) frz_append(fz, i, frozen_idx);   // contact-kernel warp start at about the same step
        }
      }
    }
    ctrl_gen<T>(spec, g, env, i, s.n, t, u);
#pragma unroll
    for (int k = 0; k < 5; k++) uc[k] = (T)u[k];
    clamp_ctrl(m, uc);   // rows keep the unclamped u, as the reference's dataset does
    // row t belongs to whoever executed the last sub-step of iteration t: the fast kernel while the env is not frozen,
    // the contact kernel from the iteration in which it took over
    const bool mine = RESUME ? (t > t0 && !(e.flags & SO101_FLAG_FROZEN)) : (!(e.flags & SO101_FLAG_FROZEN) && (t > t0 || t0 == 0));
    if (rows && active && mine) {
      ROW* row = rows + ((int64_t)i * (Tn + 1) + t) * SO101_ROW;
#pragma unroll
      for (int k = 0; k < 5; k++) row[k] = (ROW)u[k];
#pragma unroll
      for (int k = 0; k < 3; k++) row[5 + k] = (ROW)(float)site[k];
#pragma unroll
      for (int k = 0; k < 5; k++) row[8 + k] = (ROW)(float)e.q[k];
    }
  }
  if (active && !was_frozen) {
    if (RESUME) { e.flags |= SO101_FLAG_FROZEN; fz.at[i] = (t1 + 1) * frame_skip; }
    store_env(s, i, e);
  }
  if (!active || was_frozen) cnt = {0, 0, 0, 0};
  add_stats(stats, cnt);
}

struct State0 { double v[18]; };

// B control sequences U[H][5][B] from one shared state; X[B][H+1][8] float32 observations
SO101_STEP_KERNEL2(T)
k_shoot(const __grid_constant__ DevModel<T> m, StateView<T> s, const __grid_constant__ State0 s0, const T* U, int H,
        int frame_skip, float* X, uint32_t rflags, unsigned long long* stats, Frz fz) {
  constexpr bool RESUME = MODE == MODE_RESUME;
  constexpr int CM = MODE == MODE_FREEZE ? CM_FREEZE : CM_SOLVE;
  __shared__ XchStorage<T, SPLIT> xst;
  SplitXch<T>& xch = xch_of(xst);
  bool active, exit_block;
  const int64_t i = env_slot<T, SPLIT, RESUME>(s, fz, active, exit_block);
  if (exit_block) return;
  if (SPLIT && threadIdx.x >= 32) { helper_role<T, CM>(m, xch, (int64_t)H * frame_skip, frame_skip); return; }
  Env<T> e;
  if (RESUME) {
    load_env(s, i, e);
  } else {
    reset_env(m, e);
#pragma unroll
    for (int k = 0; k < NV; k++) { e.q[k] = (T)s0.v[k]; e.qd[k] = (T)s0.v[6 + k]; e.warm[k] = (T)s0.v[12 + k]; }
  }
  const int start = RESUME ? fz.at[i] : 0;
  Counters cnt = {0, 0, 0, 0};
  const bool trip = m.ntrip > 0;
  const bool hold = rflags & SO101_ROLL_GRAVCOMP_HOLD;
  T site[3];
  site_fk(m, e.q, site);
  if (SPLIT) publish_state(xch, e);
  T uc[NV] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  int frozen_idx = -1;
  int64_t nstep = 0;
#pragma unroll 1
  for (int t = 0; t <= H; t++) {
    if (t > 0) {
#pragma unroll
      for (int k = 0; k < 5; k++) uc[k] = U[((int64_t)(t - 1) * 5 + k) * s.n + i];
      clamp_ctrl(m, uc);
#pragma unroll 1
      for (int ss = 0; ss < frame_skip; ss++, nstep++) {
        if (RESUME && t * frame_skip + ss == start) e.flags &= ~SO101_FLAG_FROZEN;
        step_env<T, SPLIT, CM>(m, xch, e, uc, hold && ss == 0, ss == frame_skip - 1, site, trip, cnt, nstep);
        if (!RESUME && frozen_idx < 0 && (e.flags & SO101_FLAG_FROZEN)) { frozen_idx = t * frame_skip + ss; if (active) frz_append(fz, i, frozen_idx); }
      }
    }
    const bool mine = RESUME ? (t > 0 && !(e.flags & SO101_FLAG_FROZEN)) : !(e.flags & SO101_FLAG_FROZEN);
    if (active && mine) {
      float* x = X + ((int64_t)i * (H + 1) + t) * SO101_NOBS;
#pragma unroll
      for (int k = 0; k < 3; k++) x[k] = (float)site[k];
#pragma unroll
      for (int k = 0; k < 5; k++) x[3 + k] = (float)e.q[k];
    }
  }
  if (active) {
    if (RESUME) { e.flags |= SO101_FLAG_FROZEN; fz.at[i] = (H + 1) * frame_skip; }
    store_env(s, i, e);
  } else {
    cnt = {0, 0, 0, 0};
  }
  add_stats(stats, cnt);
}

// register-resident FMA loop for the roofline denominator ("of measured")
template <typename T> __global__ void k_fma_peak(T* out, int iters, T a, T b) {
  T x[8];
#pragma unroll
  for (int k = 0; k < 8; k++) x[k] = T(threadIdx.x + k) * T(1e-3);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = x[k] * a + b;
  }
  T s = T(0);
#pragma unroll
  for (int k = 0; k < 8; k++) s += x[k];
  if (s == T(-12345.678)) out[0] = s;  // never true; keeps the loop alive
}

// ==========================================================================================
// host side
// ==========================================================================================
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t err__ = (expr);                                                                     \
    if (err__ != cudaSuccess)                                                                       \
      return fail(SO101_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));              \
  } while (0)

struct So101Model {
  So101Tables tables;
  DevModel<double> d;
  DevModel<float> f;
  // convex hulls of the colliding geoms (so101_model_set_hulls), vertices already in link frames; empty: none
  std::vector<double> hull_vert;
  std::vector<int32_t> hull_vstart, hull_adj_start, hull_adj, hull_cube;
  int hull_res = 0;
};

struct So101Batch {
  const So101Model* model;
  int64_t n;
  int dtype, device;
  void* state;        // NROWS*n elements + n uint32 flags
  bool owns_state;
  DevModel<double> dm_d;      // the model's constants with this batch's device pointers patched in (hull data)
  DevModel<float> dm_f;
  void* hull_dev[4];          // vert, adj_start, adj, cube on b->device (null: no hulls -> tripwire flags only)
  // freeze / resume bookkeeping (struct Frz): at[n], list[n], count, snapshots of count per time chunk
  int32_t* frz_at; int32_t* frz_list; int32_t* frz_count; int32_t* frz_snap;
  unsigned long long* stats;  // device, 4 counters
  void* ctrl_stage;   // [6][n] batch dtype (host variants)
  float* obs_stage;   // [8][n]
  void* init_stage;   // [12][n] (reset_host)
  void* u_stage;      // [T+1][5][n] control tensor (rollout_host)
  size_t u_stage_bytes;
  void* rows_stage;   // [n][T+1][13] dataset rows (rollout_host)
  size_t rows_stage_bytes;
  // rollout_host pipeline: upload / download streams and per-chunk events
  static constexpr int MAXCHUNK = 12;
  bool pipe_ready;
  cudaStream_t s_up, s_down;
  cudaEvent_t ev_up[MAXCHUNK], ev_k[MAXCHUNK], ev_start;
  // explicit experiment options (so101_batch_set_option); 0 = automatic.  Nothing on this path reads the environment.
  int opt_family, opt_block, opt_host_chunks, opt_host_even, opt_contact;
};

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

static size_t elem_size(int dtype) { return dtype == SO101_F64 ? sizeof(double) : sizeof(float); }

// block size: big batches use 256-thread blocks (8 warps share the instruction stream of a step,
// see physics_step SYNC); small batches use smaller blocks so that every SM gets work
static int pick_block(int64_t n) {
  if (n >= (int64_t)148 * 256) return 256;
  if (n >= (int64_t)148 * 128) return 128;
  if (n >= (int64_t)148 * 64) return 64;
  return 32;
}
static unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

template <typename T> static StateView<T> view(const So101Batch* b) {
  StateView<T> v;
  v.base = static_cast<T*>(b->state);
  v.flags = reinterpret_cast<uint32_t*>(static_cast<char*>(b->state) + (size_t)NROWS * b->n * sizeof(T));
  v.n = b->n;
  return v;
}
// Small batches run the SPLIT kernels: three warps per 32 envs, 96-thread blocks - f64 up to two teams per SM
// (148 x 64 envs), f32 up to four (148 x 128 envs).  Measured on B200 (tools/split_probe.py, ms per 100 control steps,
// one-warp -> team): f64 1024 envs 11.3 -> 7.0, 4096 envs 11.9 -> 7.3, 9472 envs 13.1 -> 9.5, 14208 envs (three teams
// per SM do not fit 255 registers) 16.3 -> 17.6; f32 4096 envs 9.8 -> 5.9, 9472 envs 10.8 -> 6.7, 18944 envs
// 11.4 -> 7.7, 23680 envs 13.6 -> 14.5.  SO101_OPT_KERNEL_FAMILY forces one family (experiments, bitwise-equality test).
static bool pick_split(const So101Batch* b, bool f32) {
  if (b->opt_family == SO101_FAMILY_ONEWARP) return false;
  if (b->opt_family == SO101_FAMILY_TEAM) return true;
  return b->n <= (int64_t)148 * (f32 ? 128 : 64);
}
template <typename T> static StateView<T> step_view(const So101Batch* b, int& blk, unsigned& grid, bool& split) {
  StateView<T> v = view<T>(b);
  split = pick_split(b, sizeof(T) == 4);
  if (split) {
    blk = 32 * TEAM_WARPS;
    grid = (unsigned)((b->n + 31) / 32);
    return v;
  }
  blk = pick_block(b->n);
  if (sizeof(T) == 4 && blk == 256 && b->n >= (int64_t)148 * 512) blk = 512;
  if (b->opt_block) blk = b->opt_block;                 // validated against the launch bounds in so101_batch_set_option
  const int64_t warps = (b->n + 31) / 32;
  grid = (unsigned)((warps * 32 + blk - 1) / blk);
  return v;
}

// ---- freeze / resume launches ----------------------------------------------------------------------------------------
static bool has_contact(const So101Batch* b) { return b->dm_d.con_enabled != 0; }
// one-warp kernels: contact inside the stepping kernel (default) or freeze-and-resume (SO101_OPT_CONTACT_MODE = 2); the
// team kernels always handle it inside.  Measured on B200 (tools/contact_perf.py, profiles/README.md): in-kernel wins
// wherever more than a handful of envs touch the table, because the contact kernel of freeze-and-resume is a small,
// latency-bound launch that cannot overlap the fast kernel (which fills every SM's register file).
static bool contact_inline(const So101Batch* b) { return b->opt_contact != 2; }
// Frz of a fast launch (chunk = index of the time chunk within the API call; chunk 0 clears stale FROZEN bits)
static Frz frz_fast(const So101Batch* b, int chunk) {
  Frz f;
  f.at = b->frz_at; f.list = b->frz_list; f.count = b->frz_count; f.upto = nullptr; f.clear = chunk == 0;
  f.min_n = 0; f.max_n = 0x7fffffff;
  return f;
}
// lists up to this long are latency bound (less than one team per SM ... two per SM): the team version finishes them
// about twice as fast per step; beyond it the one-warp version has the throughput
constexpr int TEAM_LIST_MAX = 148 * 64;
static Frz frz_resume(const So101Batch* b, int chunk, bool team) {
  Frz f = frz_fast(b, chunk);
  f.upto = b->frz_snap + chunk;
  f.clear = 0;
  f.min_n = team ? 0 : TEAM_LIST_MAX;
  f.max_n = team ? TEAM_LIST_MAX : 0x7fffffff;
  return f;
}
// launch shape of a contact kernel: one thread (team lane) per list entry, worst case every env (blocks beyond the
// list's end exit at once)
static void resume_shape(const So101Batch* b, bool team, int& blk, unsigned& grid) {
  if (team) {
    const int64_t n = b->n < TEAM_LIST_MAX ? b->n : TEAM_LIST_MAX;
    blk = 32 * TEAM_WARPS;
    grid = (unsigned)((n + 31) / 32);
  } else {
    blk = 64;
    grid = (unsigned)((b->n + 63) / 64);
  }
}
// before the first fast launch of an API call: the list starts empty
static cudaError_t frz_begin(const So101Batch* b, int chunk, cudaStream_t st) {
  if (!has_contact(b) || chunk != 0) return cudaSuccess;
  return cudaMemsetAsync(b->frz_count, 0, sizeof(int32_t), st);
}

extern "C" {

const char* so101_last_error(void) { return g_err.c_str(); }
int so101_abi_version(void) { return SO101_ABI_VERSION; }
size_t so101_tables_sizeof(void) { return sizeof(So101Tables); }
int so101_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int so101_model_create(const So101Tables* tables, So101Model** out) {
  if (!tables || !out) return fail(SO101_EINVAL, "null argument");
  So101Model* m = new So101Model();
  m->tables = *tables;
  std::string why = hostbuild::build(*tables, m->d);
  if (!why.empty()) { delete m; return fail(SO101_EMODEL, "unsupported model: " + why); }
  hostbuild::convert<float>(m->d, m->f);
  *out = m;
  return SO101_OK;
}
void so101_model_destroy(So101Model* m) { delete m; }

int so101_model_set_hulls(So101Model* m, const So101Hulls* h) {
  if (!m || !h) return fail(SO101_EINVAL, "null argument");
  const So101Tables& t = m->tables;
  if (!t.con_enabled) return fail(SO101_EMODEL, "the tables carry no contact parameters (con_enabled = 0)");
  if (h->ngeom != t.ntrip || h->nvert <= 0 || h->nadj <= 0 || h->cube_res <= 0 || !h->vert_start || !h->vert ||
      !h->adj_start || !h->adj || !h->cube)
    return fail(SO101_EINVAL, "hulls: need one hull per tripwire box and all five arrays");
  if (h->vert_start[0] != 0 || h->vert_start[h->ngeom] != h->nvert || h->adj_start[0] != 0 || h->adj_start[h->nvert] != h->nadj)
    return fail(SO101_EINVAL, "hulls: inconsistent offsets");
  std::vector<double> vert((size_t)3 * h->nvert);
  for (int g = 0; g < h->ngeom; g++) {
    if (h->vert_start[g + 1] <= h->vert_start[g]) return fail(SO101_EINVAL, "hulls: empty hull");
    for (int i = h->vert_start[g]; i < h->vert_start[g + 1]; i++) {
      if (!hostbuild::hull_to_link(t, t.trip_body[g], h->vert + 3 * i, &vert[3 * (size_t)i]))
        return fail(SO101_EMODEL, "hulls: geom on a body without a joint");
      for (int e = h->adj_start[i]; e < h->adj_start[i + 1]; e++)
        if (h->adj[e] < h->vert_start[g] || h->adj[e] >= h->vert_start[g + 1])
          return fail(SO101_EINVAL, "hulls: an edge leaves its hull");
    }
  }
  const size_t ncube = (size_t)h->ngeom * 6 * h->cube_res * h->cube_res;
  for (size_t i = 0; i < ncube; i++) {
    const int g = (int)(i / ((size_t)6 * h->cube_res * h->cube_res));
    if (h->cube[i] < h->vert_start[g] || h->cube[i] >= h->vert_start[g + 1]) return fail(SO101_EINVAL, "hulls: cube map entry outside its hull");
  }
  m->hull_vert.swap(vert);
  m->hull_vstart.assign(h->vert_start, h->vert_start + h->ngeom + 1);
  m->hull_adj_start.assign(h->adj_start, h->adj_start + h->nvert + 1);
  m->hull_adj.assign(h->adj, h->adj + h->nadj);
  m->hull_cube.assign(h->cube, h->cube + ncube);
  m->hull_res = h->cube_res;
  return SO101_OK;
}

size_t so101_batch_state_bytes(int64_t n, int dtype) {
  if (n <= 0) return 0;
  return (size_t)NROWS * n * elem_size(dtype) + (size_t)n * sizeof(uint32_t);
}

int so101_batch_create(const So101Model* model, int64_t n, int dtype, int device, void* state_buf, So101Batch** out) {
  if (!model || !out || n <= 0) return fail(SO101_EINVAL, "bad argument");
  if (dtype != SO101_F64 && dtype != SO101_F32) return fail(SO101_EINVAL, "dtype must be SO101_F64 or SO101_F32");
  int ndev = so101_device_count();
  if (ndev <= 0) return fail(SO101_ENODEVICE, "no CUDA device visible: this library has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(SO101_EINVAL, "device index out of range");
  DeviceGuard g(device);
  if (!g.ok) return fail(SO101_ECUDA, "cudaSetDevice failed");
  So101Batch* b = new So101Batch();
  std::memset(b, 0, sizeof *b);
  b->model = model; b->n = n; b->dtype = dtype; b->device = device;
  size_t bytes = so101_batch_state_bytes(n, dtype);
  if (state_buf) { b->state = state_buf; b->owns_state = false; }
  else {
    cudaError_t e = cudaMalloc(&b->state, bytes);
    if (e != cudaSuccess) { delete b; return fail(SO101_ECUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    b->owns_state = true;
  }
  cudaError_t e = cudaMalloc(&b->stats, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(b->stats, 0, 4 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(b->state, 0, bytes);
  // the model's constants, with the hull data of this device patched in (contact path on) or contact off
  b->dm_d = model->d;
  b->dm_d.con_enabled = 0;
  if (e == cudaSuccess && model->tables.con_enabled && !model->hull_vert.empty()) {
    const void* src[4] = {model->hull_vert.data(), model->hull_adj_start.data(), model->hull_adj.data(), model->hull_cube.data()};
    const size_t nb[4] = {model->hull_vert.size() * sizeof(double), model->hull_adj_start.size() * sizeof(int32_t),
                          model->hull_adj.size() * sizeof(int32_t), model->hull_cube.size() * sizeof(int32_t)};
    for (int k = 0; k < 4 && e == cudaSuccess; k++) {
      e = cudaMalloc(&b->hull_dev[k], nb[k]);
      if (e == cudaSuccess) e = cudaMemcpy(b->hull_dev[k], src[k], nb[k], cudaMemcpyHostToDevice);
      const uint64_t w = (uint64_t)(uintptr_t)b->hull_dev[k];
      b->dm_d.hull_ptr[2 * k] = (int32_t)(uint32_t)(w & 0xffffffffu);
      b->dm_d.hull_ptr[2 * k + 1] = (int32_t)(uint32_t)(w >> 32);
    }
    b->dm_d.hull_res = model->hull_res;
    b->dm_d.con_enabled = 1;
    if (e == cudaSuccess) e = cudaMalloc(&b->frz_at, (size_t)n * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&b->frz_list, (size_t)n * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&b->frz_count, (1 + So101Batch::MAXCHUNK + 1) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMemset(b->frz_count, 0, (1 + So101Batch::MAXCHUNK + 1) * sizeof(int32_t));
    b->frz_snap = b->frz_count + 1;
  }
  hostbuild::convert<float>(b->dm_d, b->dm_f);
  if (e != cudaSuccess) {
    if (b->owns_state) cudaFree(b->state);
    cudaFree(b->stats);
    for (int k = 0; k < 4; k++) cudaFree(b->hull_dev[k]);
    cudaFree(b->frz_at); cudaFree(b->frz_list); cudaFree(b->frz_count);
    delete b;
    return fail(SO101_ECUDA, std::string("batch init: ") + cudaGetErrorString(e));
  }
  *out = b;
  return SO101_OK;
}
void so101_batch_destroy(So101Batch* b) {
  if (!b) return;
  DeviceGuard g(b->device);
  if (b->owns_state) cudaFree(b->state);
  cudaFree(b->stats);
  for (int k = 0; k < 4; k++) cudaFree(b->hull_dev[k]);
  cudaFree(b->frz_at); cudaFree(b->frz_list); cudaFree(b->frz_count);
  cudaFree(b->ctrl_stage);
  cudaFree(b->obs_stage);
  cudaFree(b->init_stage);
  cudaFree(b->u_stage);
  cudaFree(b->rows_stage);
  if (b->pipe_ready) {
    cudaStreamDestroy(b->s_up);
    cudaStreamDestroy(b->s_down);
    for (int c = 0; c < So101Batch::MAXCHUNK; c++) { cudaEventDestroy(b->ev_up[c]); cudaEventDestroy(b->ev_k[c]); }
    cudaEventDestroy(b->ev_start);
  }
  delete b;
}

int so101_batch_set_option(So101Batch* b, int option, int value) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  switch (option) {
    case SO101_OPT_KERNEL_FAMILY:
      if (value < SO101_FAMILY_AUTO || value > SO101_FAMILY_TEAM) return fail(SO101_EINVAL, "kernel family must be 0 (auto), 1 (one-warp) or 2 (team)");
      b->opt_family = value;
      return SO101_OK;
    case SO101_OPT_BLOCK: {
      const int lim = b->dtype == SO101_F64 ? SO101_F64_THREADS : SO101_F32_THREADS;   // the kernels' __launch_bounds__
      if (value != 0 && (value < 32 || value > lim || value % 32)) return fail(SO101_EINVAL, "block size must be 0 (auto) or a multiple of 32 up to the launch bound (256 f64 / 512 f32)");
      b->opt_block = value;
      return SO101_OK;
    }
    case SO101_OPT_HOST_CHUNKS:
      if (value < 0 || value > So101Batch::MAXCHUNK) return fail(SO101_EINVAL, "host pipeline chunks must be 0 (auto) .. 12");
      b->opt_host_chunks = value;
      return SO101_OK;
    case SO101_OPT_HOST_EVEN:
      b->opt_host_even = value != 0;
      return SO101_OK;
    case SO101_OPT_CONTACT_MODE:
      if (value < 0 || value > 2) return fail(SO101_EINVAL, "contact mode must be 0 (auto), 1 (in-kernel) or 2 (freeze and resume)");
      b->opt_contact = value;
      return SO101_OK;
    default:
      return fail(SO101_EINVAL, "unknown option");
  }
}

#define DISPATCH(b, CALL_D, CALL_F)              \
  do {                                           \
    if ((b)->dtype == SO101_F64) { CALL_D; }     \
    else { CALL_F; }                             \
  } while (0)

int so101_batch_reset(So101Batch* b, const void* qpos0, const void* qvel0, void* obs, void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int blk = pick_block(b->n);
  DISPATCH(b,
    (k_reset<double><<<grid_for(b->n, blk), blk, 0, st>>>(b->dm_d, view<double>(b), (const double*)qpos0,
        (const double*)qvel0, 0, 0, 0, 0.0, 0.0, (float*)obs)),
    (k_reset<float><<<grid_for(b->n, blk), blk, 0, st>>>(b->dm_f, view<float>(b), (const float*)qpos0,
        (const float*)qvel0, 0, 0, 0, 0.0, 0.0, (float*)obs)));
  CUDA_TRY(cudaGetLastError());
  return SO101_OK;
}

int so101_batch_reset_random(So101Batch* b, uint64_t seed, int64_t env_offset, double lo, double hi, void* obs,
                             void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int blk = pick_block(b->n);
  DISPATCH(b,
    (k_reset<double><<<grid_for(b->n, blk), blk, 0, st>>>(b->dm_d, view<double>(b), nullptr, nullptr, 1, seed,
        env_offset, lo, hi, (float*)obs)),
    (k_reset<float><<<grid_for(b->n, blk), blk, 0, st>>>(b->dm_f, view<float>(b), nullptr, nullptr, 1, seed,
        env_offset, lo, hi, (float*)obs)));
  CUDA_TRY(cudaGetLastError());
  return SO101_OK;
}

int so101_batch_forward(So101Batch* b, void* obs, void* qfrc_bias, void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int blk = pick_block(b->n);
  DISPATCH(b,
    (k_forward<double><<<grid_for(b->n, blk), blk, 0, st>>>(b->dm_d, view<double>(b), (float*)obs, (double*)qfrc_bias)),
    (k_forward<float><<<grid_for(b->n, blk), blk, 0, st>>>(b->dm_f, view<float>(b), (float*)obs, (float*)qfrc_bias)));
  CUDA_TRY(cudaGetLastError());
  return SO101_OK;
}

int so101_batch_step(So101Batch* b, const void* ctrl, int n_ctrl, int n_substeps, void* obs, void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  if (n_ctrl < 0 || n_ctrl > NV || n_substeps < 0) return fail(SO101_EINVAL, "n_ctrl must be 0..6, n_substeps >= 0");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int blk; unsigned grid; bool split;
  CUDA_TRY(frz_begin(b, 0, st));
  for (int pass = 0; pass < 3; pass++) {
    // one-warp kernels: pass 0 freezes the envs that reach the table, passes 1 / 2 (contact kernel, team / one-warp
    // version: the one the list length calls for does the work) finish them; team kernels: one pass, contact path inside
    if (pass == 2 && b->n <= TEAM_LIST_MAX) break;
    const Frz fz = pass ? frz_resume(b, 0, pass == 1) : frz_fast(b, 0);
#define SO101_STEP(TT, SS, MD, mdl) \
  k_step<TT, SS, MD><<<grid, blk, 0, st>>>(b->dm_##mdl, v, (const TT*)ctrl, n_ctrl, n_substeps, (float*)obs, b->stats, fz)
    if (b->dtype == SO101_F64) {
      StateView<double> v = step_view<double>(b, blk, grid, split);
      if (split) { SO101_STEP(double, true, MODE_INLINE, d); }
      else if (contact_inline(b)) { SO101_STEP(double, false, MODE_INLINE, d); }
      else if (pass == 1) { resume_shape(b, true, blk, grid); SO101_STEP(double, true, MODE_RESUME, d); }
      else if (pass == 2) { resume_shape(b, false, blk, grid); SO101_STEP(double, false, MODE_RESUME, d); }
      else SO101_STEP(double, false, MODE_FREEZE, d);
    } else {
      StateView<float> v = step_view<float>(b, blk, grid, split);
      if (split) { SO101_STEP(float, true, MODE_INLINE, f); }
      else if (contact_inline(b)) { SO101_STEP(float, false, MODE_INLINE, f); }
      else if (pass == 1) { resume_shape(b, true, blk, grid); SO101_STEP(float, true, MODE_RESUME, f); }
      else if (pass == 2) { resume_shape(b, false, blk, grid); SO101_STEP(float, false, MODE_RESUME, f); }
      else SO101_STEP(float, false, MODE_FREEZE, f);
    }
#undef SO101_STEP
    CUDA_TRY(cudaGetLastError());
    if (split || contact_inline(b) || !has_contact(b)) break;
    if (!pass) { k_snapshot<<<1, 1, 0, st>>>(b->frz_count, b->frz_snap); CUDA_TRY(cudaGetLastError()); }
  }
  return SO101_OK;
}

static int ensure_stage(So101Batch* b) {
  if (!b->ctrl_stage) CUDA_TRY(cudaMalloc(&b->ctrl_stage, (size_t)NV * b->n * elem_size(b->dtype)));
  if (!b->obs_stage) CUDA_TRY(cudaMalloc(&b->obs_stage, (size_t)SO101_NOBS * b->n * sizeof(float)));
  return SO101_OK;
}

int so101_batch_step_host(So101Batch* b, const void* ctrl_host, int n_ctrl, int n_substeps, void* obs_host,
                          uint32_t* flags_host, void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  if (n_ctrl < 0 || n_ctrl > NV) return fail(SO101_EINVAL, "n_ctrl must be 0..6");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_stage(b);
  if (rc) return rc;
  if (ctrl_host && n_ctrl > 0)
    CUDA_TRY(cudaMemcpyAsync(b->ctrl_stage, ctrl_host, (size_t)n_ctrl * b->n * elem_size(b->dtype),
                             cudaMemcpyHostToDevice, st));
  rc = so101_batch_step(b, ctrl_host ? b->ctrl_stage : nullptr, n_ctrl, n_substeps, b->obs_stage, stream);
  if (rc) return rc;
  if (obs_host)
    CUDA_TRY(cudaMemcpyAsync(obs_host, b->obs_stage, (size_t)SO101_NOBS * b->n * sizeof(float),
                             cudaMemcpyDeviceToHost, st));
  if (flags_host)
    CUDA_TRY(cudaMemcpyAsync(flags_host, static_cast<char*>(b->state) + (size_t)NROWS * b->n * elem_size(b->dtype),
                             (size_t)b->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (flags_host)
    for (int64_t i = 0; i < b->n; i++) flags_host[i] &= ~SO101_FLAG_FROZEN;   // internal book-keeping bit
  return SO101_OK;
}

int so101_batch_reset_host(So101Batch* b, const void* qpos0_host, const void* qvel0_host, void* obs_host,
                           void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = ensure_stage(b);
  if (rc) return rc;
  size_t blockb = (size_t)NV * b->n * elem_size(b->dtype);
  if (!b->init_stage) CUDA_TRY(cudaMalloc(&b->init_stage, 2 * blockb));
  char* qd = static_cast<char*>(b->init_stage);
  if (qpos0_host) CUDA_TRY(cudaMemcpyAsync(qd, qpos0_host, blockb, cudaMemcpyHostToDevice, st));
  if (qvel0_host) CUDA_TRY(cudaMemcpyAsync(qd + blockb, qvel0_host, blockb, cudaMemcpyHostToDevice, st));
  rc = so101_batch_reset(b, qpos0_host ? qd : nullptr, qvel0_host ? qd + blockb : nullptr, b->obs_stage, stream);
  if (rc) return rc;
  if (obs_host)
    CUDA_TRY(cudaMemcpyAsync(obs_host, b->obs_stage, (size_t)SO101_NOBS * b->n * sizeof(float),
                             cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return SO101_OK;
}

// One time chunk (t0, t1] of a rollout; `chunk` = index of the chunk within the API call (chunk 0 empties the list of
// frozen envs).  One-warp kernels: the fast kernel, then the contact kernel for the envs frozen so far.
static int rollout_range(So101Batch* b, const So101CtrlSpec* spec, int t0, int t1, int T, int frame_skip, void* rows,
                         uint32_t flags, void* stream, int chunk = 0) {
  if (!b || !spec) return fail(SO101_EINVAL, "null argument");
  if (T < 0 || frame_skip < 1) return fail(SO101_EINVAL, "T must be >= 0 and frame_skip >= 1");
  if (spec->kind < SO101_CTRL_RANDOM || spec->kind > SO101_CTRL_TENSOR) return fail(SO101_EINVAL, "bad control kind");
  if (spec->kind == SO101_CTRL_TENSOR && !spec->u) return fail(SO101_EINVAL, "SO101_CTRL_TENSOR needs spec->u");
  if (spec->kind == SO101_CTRL_CHIRP && spec->t_total <= 0) return fail(SO101_EINVAL, "chirp needs t_total > 0");
  if (chunk < 0 || chunk > So101Batch::MAXCHUNK) return fail(SO101_EINVAL, "too many time chunks");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DevSpec ds;
  ds.kind = spec->kind; ds.t_total = spec->t_total; ds.seed = spec->seed; ds.env_offset = spec->env_offset;
  ds.amp = spec->amp; ds.freq_lo = spec->freq_lo; ds.freq_hi = spec->freq_hi;
  ds.reset_lo = spec->reset_lo; ds.reset_hi = spec->reset_hi; ds.u = spec->u;
  int blk; unsigned grid; bool split;
  const bool r32 = flags & SO101_ROLL_ROWS_F32;
  CUDA_TRY(frz_begin(b, chunk, st));
  for (int pass = 0; pass < 3; pass++) {
    if (pass == 2 && b->n <= TEAM_LIST_MAX) break;
    const Frz fz = pass ? frz_resume(b, chunk, pass == 1) : frz_fast(b, chunk);
#define SO101_ROLL(TT, RR, SS, MD, mdl) \
  k_rollout<TT, RR, SS, MD><<<grid, blk, 0, st>>>(b->dm_##mdl, v, ds, t0, t1, T, frame_skip, (RR*)rows, flags, b->stats, fz)
#define SO101_ROLL_ROWS(TT, SS, MD, mdl) do { if (r32) SO101_ROLL(TT, float, SS, MD, mdl); else SO101_ROLL(TT, double, SS, MD, mdl); } while (0)
    if (b->dtype == SO101_F64) {
      StateView<double> v = step_view<double>(b, blk, grid, split);
      if (split) SO101_ROLL_ROWS(double, true, MODE_INLINE, d);
      else if (contact_inline(b)) SO101_ROLL_ROWS(double, false, MODE_INLINE, d);
      else if (pass == 1) { resume_shape(b, true, blk, grid); SO101_ROLL_ROWS(double, true, MODE_RESUME, d); }
      else if (pass == 2) { resume_shape(b, false, blk, grid); SO101_ROLL_ROWS(double, false, MODE_RESUME, d); }
      else SO101_ROLL_ROWS(double, false, MODE_FREEZE, d);
    } else {
      StateView<float> v = step_view<float>(b, blk, grid, split);
      if (split) SO101_ROLL_ROWS(float, true, MODE_INLINE, f);
      else if (contact_inline(b)) SO101_ROLL_ROWS(float, false, MODE_INLINE, f);
      else if (pass == 1) { resume_shape(b, true, blk, grid); SO101_ROLL_ROWS(float, true, MODE_RESUME, f); }
      else if (pass == 2) { resume_shape(b, false, blk, grid); SO101_ROLL_ROWS(float, false, MODE_RESUME, f); }
      else SO101_ROLL_ROWS(float, false, MODE_FREEZE, f);
    }
#undef SO101_ROLL_ROWS
#undef SO101_ROLL
    CUDA_TRY(cudaGetLastError());
    if (split || contact_inline(b) || !has_contact(b)) break;
    if (!pass) { k_snapshot<<<1, 1, 0, st>>>(b->frz_count, b->frz_snap + chunk); CUDA_TRY(cudaGetLastError()); }
  }
  return SO101_OK;
}

int so101_batch_rollout(So101Batch* b, const So101CtrlSpec* spec, int T, int frame_skip, void* rows, uint32_t flags,
                        void* stream) {
  return rollout_range(b, spec, 0, T, T, frame_skip, rows, flags, stream);
}

static int grow(void** buf, size_t* have, size_t need) {
  if (*have >= need) return SO101_OK;
  if (*buf) cudaFree(*buf);
  *buf = nullptr; *have = 0;
  CUDA_TRY(cudaMalloc(buf, need));
  *have = need;
  return SO101_OK;
}

// Host-buffer rollout, pipelined over time: the rollout is cut into up to 12 chunks of control steps; the control
// tensor of chunk c+1 is uploaded (stream up) and the rows of chunk c-1 are downloaded (stream down, strided copy into
// the caller's [N][T+1][13] layout) while chunk c computes on the caller's stream.  Only the first upload and the last
// download are exposed.
static int ensure_pipe(So101Batch* b) {
  if (b->pipe_ready) return SO101_OK;
  CUDA_TRY(cudaStreamCreateWithFlags(&b->s_up, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&b->s_down, cudaStreamNonBlocking));
  for (int c = 0; c < So101Batch::MAXCHUNK; c++) {
    CUDA_TRY(cudaEventCreateWithFlags(&b->ev_up[c], cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&b->ev_k[c], cudaEventDisableTiming));
  }
  CUDA_TRY(cudaEventCreateWithFlags(&b->ev_start, cudaEventDisableTiming));
  b->pipe_ready = true;
  return SO101_OK;
}

int so101_batch_rollout_host(So101Batch* b, const So101CtrlSpec* spec, const void* qpos0_host, int T, int frame_skip,
                             void* rows_host, uint32_t flags, void* stream) {
  if (!b || !spec || !rows_host) return fail(SO101_EINVAL, "null argument");
  if (T < 0) return fail(SO101_EINVAL, "T must be >= 0");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t es = elem_size(b->dtype);
  const size_t rs = (flags & SO101_ROLL_ROWS_F32) ? sizeof(float) : sizeof(double);
  So101CtrlSpec dspec = *spec;
  int rc;
  if ((rc = ensure_pipe(b))) return rc;
  const bool tensor = spec->kind == SO101_CTRL_TENSOR;
  const size_t ustep = (size_t)SO101_NU_ENV * b->n * es;         // one control step of the tensor
  if (tensor) {
    if (!spec->u) return fail(SO101_EINVAL, "SO101_CTRL_TENSOR needs spec->u");
    if ((rc = grow(&b->u_stage, &b->u_stage_bytes, (size_t)(T + 1) * ustep))) return rc;
    dspec.u = b->u_stage;
  }
  const size_t rb = (size_t)b->n * (T + 1) * SO101_ROW * rs;
  if ((rc = grow(&b->rows_stage, &b->rows_stage_bytes, rb))) return rc;
  // chunking: worth it only when there is something to overlap
  int nchunk = 1;
  if (T >= 8 && rb >= ((size_t)4 << 20)) {   // >= 2 MB of rows per chunk
    nchunk = (int)(rb >> 21);
    if (nchunk > So101Batch::MAXCHUNK) nchunk = So101Batch::MAXCHUNK;
    if (nchunk < 2) nchunk = 2;
  }
  if (b->opt_host_chunks) nchunk = b->opt_host_chunks;
  if (nchunk > T) nchunk = T > 0 ? T : 1;
  // An error return inside the pipeline must not leave copies in flight that touch the caller's host buffers.
  auto drain = [&]() { cudaStreamSynchronize(b->s_up); cudaStreamSynchronize(b->s_down); cudaStreamSynchronize(st); };
#define PIPE_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t err__ = (expr);                                                                     \
    if (err__ != cudaSuccess) {                                                                     \
      drain();                                                                                      \
      return fail(SO101_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));              \
    }                                                                                               \
  } while (0)
  // the side streams start after whatever the caller queued on `st`
  PIPE_TRY(cudaEventRecord(b->ev_start, st));
  PIPE_TRY(cudaStreamWaitEvent(b->s_up, b->ev_start, 0));
  PIPE_TRY(cudaStreamWaitEvent(b->s_down, b->ev_start, 0));
  if (qpos0_host) {   // explicit initial joint angles: reset on the device, then continue from that state
    const size_t qb = (size_t)NV * b->n * es;
    if (!b->init_stage) PIPE_TRY(cudaMalloc(&b->init_stage, 2 * qb));
    PIPE_TRY(cudaMemcpyAsync(b->init_stage, qpos0_host, qb, cudaMemcpyHostToDevice, st));
    if ((rc = so101_batch_reset(b, b->init_stage, nullptr, nullptr, stream))) { drain(); return rc; }
    flags |= SO101_ROLL_NO_RESET;
  }
  // chunk c covers control steps (tb[c], tb[c+1]]; rows tb[c]+1 .. tb[c+1] (+ row 0 for the first chunk)
  // Boundaries: chunks that double from 2 control steps (every upload finishes while the previous, half as long
  // chunk computes, even at a fraction of the PCIe rate - right after the caller flushed L2 uploads are ~3x slower),
  // then chunks that halve towards the end (the download of the last chunk is the only exposed download).
  // SO101_OPT_HOST_EVEN: equal chunks.
  int tb[So101Batch::MAXCHUNK + 1];
  if (nchunk >= 4 && !b->opt_host_even) {
    int k = 0, pos = 0;
    tb[k++] = 0;
    for (int sz = 2; pos + sz < T / 2 && k < nchunk - 3; sz *= 2) { pos += sz; tb[k++] = pos; }
    if (k == 1) { pos = T < 2 ? T : 2; tb[k++] = pos; }
    const int rest = T - pos;
    for (int sh = 1; k < nchunk && tb[k - 1] < T; sh++) {
      const int t = T - (rest >> sh);
      if (t > tb[k - 1]) tb[k++] = t;
      if ((rest >> sh) == 0) break;
    }
    if (tb[k - 1] < T) tb[k++] = T;
    nchunk = k - 1;
  } else {
    for (int c = 0; c <= nchunk; c++) tb[c] = (int)((int64_t)T * c / nchunk);
  }
  if (tensor) {
    for (int c = 0; c < nchunk; c++) {   // u_t for t in [first, tb[c+1]]: chunk c reads u at tb[c] .. tb[c+1]
      const int first = c == 0 ? 0 : tb[c] + 1, last = tb[c + 1];
      PIPE_TRY(cudaMemcpyAsync(static_cast<char*>(b->u_stage) + first * ustep,
                               static_cast<const char*>(spec->u) + first * ustep, (size_t)(last - first + 1) * ustep,
                               cudaMemcpyHostToDevice, b->s_up));
      PIPE_TRY(cudaEventRecord(b->ev_up[c], b->s_up));
    }
  }
  const size_t pitch = (size_t)(T + 1) * SO101_ROW * rs;
  for (int c = 0; c < nchunk; c++) {
    if (tensor) PIPE_TRY(cudaStreamWaitEvent(st, b->ev_up[c], 0));
    const uint32_t f = c == 0 ? flags : (flags | SO101_ROLL_NO_RESET);
    if ((rc = rollout_range(b, &dspec, tb[c], tb[c + 1], T, frame_skip, b->rows_stage, f, stream, c))) { drain(); return rc; }
    PIPE_TRY(cudaEventRecord(b->ev_k[c], st));
    PIPE_TRY(cudaStreamWaitEvent(b->s_down, b->ev_k[c], 0));
    const int first = c == 0 ? 0 : tb[c] + 1, last = tb[c + 1];
    const size_t off = (size_t)first * SO101_ROW * rs, width = (size_t)(last - first + 1) * SO101_ROW * rs;
    if (nchunk == 1)
      PIPE_TRY(cudaMemcpyAsync(rows_host, b->rows_stage, rb, cudaMemcpyDeviceToHost, b->s_down));
    else
      PIPE_TRY(cudaMemcpy2DAsync(static_cast<char*>(rows_host) + off, pitch, static_cast<char*>(b->rows_stage) + off,
                                 pitch, width, (size_t)b->n, cudaMemcpyDeviceToHost, b->s_down));
  }
#undef PIPE_TRY
  CUDA_TRY(cudaStreamSynchronize(b->s_down));
  CUDA_TRY(cudaStreamSynchronize(st));
  return SO101_OK;
}

int so101_batch_shoot(So101Batch* b, const double* state0, const void* U, int H, int frame_skip, void* X,
                      uint32_t flags, void* stream) {
  if (H < 0 || frame_skip < 1) return fail(SO101_EINVAL, "H must be >= 0 and frame_skip >= 1");
  if (!b || !state0 || (!U && H > 0) || !X) return fail(SO101_EINVAL, "null argument");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  State0 s0;
  std::memcpy(s0.v, state0, sizeof s0.v);
  int blk; unsigned grid; bool split;
  CUDA_TRY(frz_begin(b, 0, st));
  for (int pass = 0; pass < 3; pass++) {
    if (pass == 2 && b->n <= TEAM_LIST_MAX) break;
    const Frz fz = pass ? frz_resume(b, 0, pass == 1) : frz_fast(b, 0);
#define SO101_SHOOT(TT, SS, MD, mdl) \
  k_shoot<TT, SS, MD><<<grid, blk, 0, st>>>(b->dm_##mdl, v, s0, (const TT*)U, H, frame_skip, (float*)X, flags, b->stats, fz)
    if (b->dtype == SO101_F64) {
      StateView<double> v = step_view<double>(b, blk, grid, split);
      if (split) { SO101_SHOOT(double, true, MODE_INLINE, d); }
      else if (contact_inline(b)) { SO101_SHOOT(double, false, MODE_INLINE, d); }
      else if (pass == 1) { resume_shape(b, true, blk, grid); SO101_SHOOT(double, true, MODE_RESUME, d); }
      else if (pass == 2) { resume_shape(b, false, blk, grid); SO101_SHOOT(double, false, MODE_RESUME, d); }
      else SO101_SHOOT(double, false, MODE_FREEZE, d);
    } else {
      StateView<float> v = step_view<float>(b, blk, grid, split);
      if (split) { SO101_SHOOT(float, true, MODE_INLINE, f); }
      else if (contact_inline(b)) { SO101_SHOOT(float, false, MODE_INLINE, f); }
      else if (pass == 1) { resume_shape(b, true, blk, grid); SO101_SHOOT(float, true, MODE_RESUME, f); }
      else if (pass == 2) { resume_shape(b, false, blk, grid); SO101_SHOOT(float, false, MODE_RESUME, f); }
      else SO101_SHOOT(float, false, MODE_FREEZE, f);
    }
#undef SO101_SHOOT
    CUDA_TRY(cudaGetLastError());
    if (split || contact_inline(b) || !has_contact(b)) break;
    if (!pass) { k_snapshot<<<1, 1, 0, st>>>(b->frz_count, b->frz_snap); CUDA_TRY(cudaGetLastError()); }
  }
  return SO101_OK;
}

static int copy_rows(So101Batch* b, int row, void* dst, const void* src, void* stream) {
  size_t es = elem_size(b->dtype), bytes = (size_t)NV * b->n * es;
  char* base = static_cast<char*>(b->state) + (size_t)row * b->n * es;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dst) CUDA_TRY(cudaMemcpyAsync(dst, base, bytes, cudaMemcpyDeviceToDevice, st));
  if (src) CUDA_TRY(cudaMemcpyAsync(base, src, bytes, cudaMemcpyDeviceToDevice, st));
  return SO101_OK;
}
int so101_batch_get_state(So101Batch* b, void* qpos, void* qvel, void* warm, void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  DeviceGuard g(b->device);
  int rc;
  if ((rc = copy_rows(b, ROW_Q, qpos, nullptr, stream))) return rc;
  if ((rc = copy_rows(b, ROW_QD, qvel, nullptr, stream))) return rc;
  return copy_rows(b, ROW_WARM, warm, nullptr, stream);
}
int so101_batch_set_state(So101Batch* b, const void* qpos, const void* qvel, const void* warm, void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  DeviceGuard g(b->device);
  int rc;
  if ((rc = copy_rows(b, ROW_Q, nullptr, qpos, stream))) return rc;
  if ((rc = copy_rows(b, ROW_QD, nullptr, qvel, stream))) return rc;
  return copy_rows(b, ROW_WARM, nullptr, warm, stream);
}
int so101_batch_set_qfrc_applied(So101Batch* b, const void* fa, void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  DeviceGuard g(b->device);
  return copy_rows(b, ROW_FA, nullptr, fa, stream);
}
int so101_batch_get_flags(So101Batch* b, uint32_t* flags_dev, void* stream) {
  if (!b || !flags_dev) return fail(SO101_EINVAL, "null argument");
  DeviceGuard g(b->device);
  const char* src = static_cast<char*>(b->state) + (size_t)NROWS * b->n * elem_size(b->dtype);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaMemcpyAsync(flags_dev, src, (size_t)b->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  if (has_contact(b)) {   // the FROZEN bit is internal book-keeping
    k_mask_flags<<<(unsigned)((b->n + 255) / 256), 256, 0, st>>>(flags_dev, b->n, ~SO101_FLAG_FROZEN);
    CUDA_TRY(cudaGetLastError());
  }
  return SO101_OK;
}
int so101_batch_clear_flags(So101Batch* b, void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  DeviceGuard g(b->device);
  char* dst = static_cast<char*>(b->state) + (size_t)NROWS * b->n * elem_size(b->dtype);
  CUDA_TRY(cudaMemsetAsync(dst, 0, (size_t)b->n * sizeof(uint32_t), static_cast<cudaStream_t>(stream)));
  return SO101_OK;
}
int so101_batch_stats(So101Batch* b, uint64_t* stats_host, void* stream) {
  if (!b || !stats_host) return fail(SO101_EINVAL, "null argument");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaMemcpyAsync(stats_host, b->stats, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemsetAsync(b->stats, 0, 4 * sizeof(uint64_t), st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return SO101_OK;
}

// ---- dataset buffer shared between the ranks of one node (CUDA IPC) -----------------------------------------------
int so101_shared_alloc(int device, size_t bytes, void** ptr, unsigned char* handle) {
  if (!ptr || !handle || bytes == 0) return fail(SO101_EINVAL, "bad argument");
  if (so101_device_count() <= 0) return fail(SO101_ENODEVICE, "no CUDA device visible");
  DeviceGuard g(device);
  if (!g.ok) return fail(SO101_ECUDA, "cudaSetDevice failed");
  static_assert(sizeof(cudaIpcMemHandle_t) == SO101_IPC_HANDLE_BYTES, "handle size");
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return fail(SO101_ECUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
  std::memcpy(handle, &h, sizeof h);
  *ptr = p;
  return SO101_OK;
}
int so101_shared_open(int device, const unsigned char* handle, void** ptr) {
  if (!ptr || !handle) return fail(SO101_EINVAL, "bad argument");
  if (so101_device_count() <= 0) return fail(SO101_ENODEVICE, "no CUDA device visible");
  DeviceGuard g(device);
  if (!g.ok) return fail(SO101_ECUDA, "cudaSetDevice failed");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof h);
  CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SO101_OK;
}
int so101_shared_close(int device, void* ptr) {
  if (!ptr) return SO101_OK;
  DeviceGuard g(device);
  CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return SO101_OK;
}
int so101_shared_free(int device, void* ptr) {
  if (!ptr) return SO101_OK;
  DeviceGuard g(device);
  CUDA_TRY(cudaFree(ptr));
  return SO101_OK;
}

int so101_fma_peak(int dtype, int device, double* tflops_out) {
  if (!tflops_out) return fail(SO101_EINVAL, "null argument");
  if (so101_device_count() <= 0) return fail(SO101_ENODEVICE, "no CUDA device visible");
  DeviceGuard g(device);
  if (!g.ok) return fail(SO101_ECUDA, "cudaSetDevice failed");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
  void* out = nullptr;
  CUDA_TRY(cudaMalloc(&out, 64));
  cudaEvent_t e0, e1;
  CUDA_TRY(cudaEventCreate(&e0));
  CUDA_TRY(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    CUDA_TRY(cudaEventRecord(e0));
    if (dtype == SO101_F64) k_fma_peak<double><<<blocks, threads>>>((double*)out, iters, 0.999999, 1e-7);
    else k_fma_peak<float><<<blocks, threads>>>((float*)out, iters, 0.999999f, 1e-7f);
    CUDA_TRY(cudaEventRecord(e1));
    CUDA_TRY(cudaEventSynchronize(e1));
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  double flops = 2.0 * 8 * (double)iters * threads * blocks;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  return SO101_OK;
}

}  // extern "C"

#include "so101_koopman.cuh"
#include "so101_ik.cuh"
