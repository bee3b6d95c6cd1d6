// so101_kernels.cuh - the stepping kernels (k_step, k_rollout, k_shoot) and their device helpers.  Compiled once per
// (dtype, kernel family) by so101_kernels.cu so that the four translation units build in parallel; the host side
// (so101_capi.cu) sees only the launchers declared in so101_launch.h.
//
// One thread owns one environment for a whole launch: its state (qpos, qvel, qacc_warmstart,
// qfrc_applied: 25 scalars) is loaded once from the structure-of-arrays buffer, stepped
// n_substeps (step) or T*frame_skip (rollout, shoot) times entirely in registers, and stored
// once.  Loads/stores are coalesced (lane i <-> env i, consecutive addresses per SoA row).
// The path is FP64/FP32-pipe bound (about 160 FLOP per byte of state traffic); tensor cores
// and TMA are deliberately unused (no dense contraction, 100 bytes of state per thread).
#pragma once
#include "so101_physics.cuh"
#include "so101_launch.h"

using namespace so101;

// ==========================================================================================
// device helpers
// ==========================================================================================

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
template <typename T>
SO101_DEV void helper_role(const DevModel<T>& m, SplitXch<T>& x, int64_t nsteps, int frame_skip) {
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const bool trip = m.ntrip > 0;
  T q[NV], qd[NV];
  int32_t vcache[SO101_MAXTRIP];   // lookout: support vertex of each hull when it was last tested (see contact_geometry)
#pragma unroll
  for (int g = 0; g < SO101_MAXTRIP; g++) vcache[g] = -1;
  vcache[SELF_BUDGET_SLOT] = vcache[SELF_REST_SLOT] = 0;   // float bits: separation budgets of the self-collision test (so101_physics.cuh)
  __syncthreads();   // (0) initial state published
#pragma unroll
  for (int k = 0; k < NV; k++) { q[k] = x.q[k][lane]; qd[k] = x.qd[k][lane]; }
  if (role == 1) {
#pragma unroll 1
    for (int64_t n = 0; n < nsteps; n++) split_geometry_step(m, x, lane, q, qd, n);
  } else {
    int ss = 0;
    x.trip[lane] = 0;   // the lookout's own word (the dynamics warp reads it after barrier (A)); see split_lookout_step
#pragma unroll 1
    for (int64_t n = 0; n < nsteps; n++) {
      split_lookout_step<T>(m, x, lane, q, qd, ss == frame_skip - 1, trip, vcache);
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
template <typename T, bool SPLIT>
SO101_DEV void step_env(const DevModel<T>& m, SplitXch<T>& x, Env<T>& e, const T (&ctrl)[NV], bool gravcomp_capture,
                        bool want_site, T (&site)[3], bool trip, Counters& cnt, int64_t nstep, int32_t* vcache) {
  if (SPLIT) split_dynamics_step<T>(m, x, threadIdx.x & 31, e, ctrl, gravcomp_capture, want_site, site, trip, cnt, nstep);
#ifdef SO101_EXP_NOSYNC   // experiment: no block barriers in the one-warp kernels (see profiles/README.md)
  else physics_step<T, false>(m, e, ctrl, gravcomp_capture, want_site, site, trip, cnt, vcache);
#else
  else physics_step<T, true>(m, e, ctrl, gravcomp_capture, want_site, site, trip, cnt, vcache);
#endif
}

// thread -> env of a stepping kernel; `active` = this thread owns the env (loads, stores, writes rows); exit_block = the
// whole block lies beyond the batch (uniform over the block)
template <typename T, bool SPLIT>
SO101_DEV int64_t env_slot(const StateView<T>& s, bool& active, bool& exit_block) {
  const int64_t j = SPLIT ? (int64_t)blockIdx.x * 32 + (threadIdx.x & 31) : (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t first = SPLIT ? (int64_t)blockIdx.x * 32 : (int64_t)blockIdx.x * blockDim.x;
  exit_block = first >= s.n;
  if (exit_block) { active = false; return 0; }
  active = (!SPLIT || threadIdx.x < 32) && j < s.n;
  return j < s.n ? j : s.n - 1;      // tail threads shadow a valid env and never store
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
static __device__ __noinline__ void uniform5(uint64_t seed, int64_t env, uint32_t step, uint32_t stream, double* out) {
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

// SOARM101Env.step: ctrl rows [n_ctrl][N] (missing rows = 0), nsub x mj_step, observation.  Table contact is handled
// inside the step (so101_contact.cuh).  (A freeze-and-resume variant - fast kernels without contact code that froze an env
// at its first table contact, contact kernels that finished the listed envs - was built and measured in round 2: never
// faster than the in-kernel path, see profiles/README.md; removed.)
#define SO101_STEP_KERNEL2(T) \
  template <typename T, bool SPLIT> __global__ void __launch_bounds__(SPLIT ? 32 * TEAM_WARPS : SO101_LB_THREADS, SPLIT ? SO101_TEAM_MINBLOCKS : SO101_LB_BLOCKS)
SO101_STEP_KERNEL2(T)
k_step(const __grid_constant__ DevModel<T> m, StateView<T> s, const T* ctrl, int n_ctrl, int nsub, float* obs,
       unsigned long long* stats, uint32_t sflags) {
  __shared__ XchStorage<T, SPLIT> xst;
  SplitXch<T>& xch = xch_of(xst);
  bool active, exit_block;
  const int64_t i = env_slot<T, SPLIT>(s, active, exit_block);
  if (exit_block) return;
  if (SPLIT && threadIdx.x >= 32) { helper_role<T>(m, xch, nsub, nsub); return; }
  Env<T> e;
  load_env(s, i, e);
  if (SPLIT) publish_state(xch, e);
  T u[NV], site[3] = {T(0), T(0), T(0)};
#pragma unroll
  for (int k = 0; k < NV; k++) u[k] = (ctrl && k < n_ctrl) ? ctrl[k * s.n + i] : T(0);
  clamp_ctrl(m, u);
  Counters cnt = {0, 0, 0, 0};
  const bool trip = m.ntrip > 0;
  int32_t vcache[SO101_MAXTRIP];   // support vertex of each hull when it was last tested (see contact_geometry)
#pragma unroll
  for (int g = 0; g < SO101_MAXTRIP; g++) vcache[g] = -1;
  vcache[SELF_BUDGET_SLOT] = vcache[SELF_REST_SLOT] = 0;   // float bits: separation budgets of the self-collision test (so101_physics.cuh)
  const bool hold = sflags & SO101_ROLL_GRAVCOMP_HOLD;   // qfrc_applied = qfrc_bias of the state the env step starts from
#pragma unroll 1
  for (int ss = 0; ss < nsub; ss++) step_env<T, SPLIT>(m, xch, e, u, hold && ss == 0, ss == nsub - 1, site, trip, cnt, ss, vcache);
  if (nsub == 0) site_fk(m, e.q, site);
  if (active) {
    store_env(s, i, e);
    if (obs) {
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

// Row t of the 32 envs of a warp -> rows[env][t][13].  A thread writing its own row issues 13 stores of 8 bytes at a
// stride of (T+1) * 104 bytes between lanes: every store instruction touches 32 sectors and fills a quarter of each.
// Harmless in local HBM (L2 merges them), but in the multi-GPU job the rows go straight into rank 0's memory over NVLink,
// where the partial sectors capped the delivery at ~90 GB/s (8 GPUs: 21.8 ms exposed behind a 10.7 ms simulation).  Here
// the warp transposes its rows through shared memory (two passes of 7 and 6 columns) and consecutive lanes store
// consecutive words of the 104-byte rows: ~4x fewer sectors on the wire.  env0 = env of lane 0 (lanes hold consecutive
// envs; lanes beyond the batch shadow a valid env and are skipped).
template <typename ROW> struct RowBuf { ROW w[7 * 32]; };
template <typename ROW>
SO101_DEV void write_rows_warp(RowBuf<ROW>& buf, const ROW (&v)[SO101_ROW], ROW* rows, int64_t env0, int64_t n, int Tn, int t) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int c0 = 0; c0 < SO101_ROW; c0 += 7) {
    const int nc = SO101_ROW - c0 < 7 ? SO101_ROW - c0 : 7;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 7; c++)
      if (c < nc) buf.w[lane * nc + c] = v[c0 + c];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 7; j++) {
      const int el = j * 32 + lane;               // element of the [32 envs][nc] block
      if (j < nc) {
        const int envl = el / nc, c = el - envl * nc;
        if (env0 + envl < n) rows[((env0 + envl) * (int64_t)(Tn + 1) + t) * SO101_ROW + c0 + c] = buf.w[el];
      }
    }
  }
}

// the same for lanes that hold arbitrary envs (regrouped batches, see k_rollout_sliced): env ids through shared memory
template <typename ROW> struct RowBufP { ROW w[7 * 32]; int64_t env[32]; };
template <typename ROW>
SO101_DEV void write_rows_warp_perm(RowBufP<ROW>& buf, const ROW (&v)[SO101_ROW], ROW* rows, int64_t env, bool active, int Tn, int t) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  buf.env[lane] = active ? env : -1;
#pragma unroll
  for (int c0 = 0; c0 < SO101_ROW; c0 += 7) {
    const int nc = SO101_ROW - c0 < 7 ? SO101_ROW - c0 : 7;
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 7; c++)
      if (c < nc) buf.w[lane * nc + c] = v[c0 + c];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 7; j++) {
      const int el = j * 32 + lane;
      if (j < nc) {
        const int envl = el / nc, c = el - envl * nc;
        const int64_t ei = buf.env[envl];
        if (ei >= 0) rows[(ei * (int64_t)(Tn + 1) + t) * SO101_ROW + c0 + c] = buf.w[el];
      }
    }
  }
}

// SOARM101DataGenerator.generate_physics_based_data, one env per thread:
// rows[N][T+1][13] = [u_t(5) | float32(ee_pos)(3) | float32(qpos[0:5])(5)]
template <typename T, typename ROW, bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? 32 * TEAM_WARPS : SO101_LB_THREADS, SPLIT ? SO101_TEAM_MINBLOCKS : SO101_LB_BLOCKS)
k_rollout(const __grid_constant__ DevModel<T> m, StateView<T> s, DevSpec spec, int t0, int t1, int Tn, int frame_skip,
          ROW* rows, uint32_t rflags, unsigned long long* stats) {
  // control steps (t0, t1] of a rollout of Tn steps; t0 > 0 continues a previous launch (row t0 is already written,
  // u_t0 is regenerated: the control stream is a pure function of (seed, env, t)).
  __shared__ XchStorage<T, SPLIT> xst;
  __shared__ RowBuf<ROW> rowbuf[SPLIT ? 1 : SO101_LB_THREADS / 32];
  SplitXch<T>& xch = xch_of(xst);
  bool active, exit_block;
  const int64_t i = env_slot<T, SPLIT>(s, active, exit_block);
  if (exit_block) return;
  if (SPLIT && threadIdx.x >= 32) { helper_role<T>(m, xch, (int64_t)(t1 - t0) * frame_skip, frame_skip); return; }
  const int64_t env = spec.env_offset + i;
  Env<T> e;
  if (rflags & SO101_ROLL_NO_RESET) {
    load_env(s, i, e);
  } else {
    reset_env(m, e);
    double r[5];
    uniform5(spec.seed, env, 0, STREAM_RESET, r);
#pragma unroll
    for (int k = 0; k < 5; k++) e.q[k] = (T)muladd_rn(spec.reset_lo, spec.reset_hi - spec.reset_lo, r[k]);
  }
  CtrlGen g;
  ctrl_init(spec, env, g);
  Counters cnt = {0, 0, 0, 0};
  const bool trip = m.ntrip > 0;
  int32_t vcache[SO101_MAXTRIP];   // support vertex of each hull when it was last tested (see contact_geometry)
#pragma unroll
  for (int g = 0; g < SO101_MAXTRIP; g++) vcache[g] = -1;
  vcache[SELF_BUDGET_SLOT] = vcache[SELF_REST_SLOT] = 0;   // float bits: separation budgets of the self-collision test (so101_physics.cuh)
  const bool hold = rflags & SO101_ROLL_GRAVCOMP_HOLD;
  T site[3];
  site_fk(m, e.q, site);
  if (SPLIT) publish_state(xch, e);
  double u[5];
  T uc[NV] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  int64_t nstep = 0;
#pragma unroll 1
  for (int t = t0; t <= t1; t++) {
    if (t > t0) {
#pragma unroll 1
      for (int ss = 0; ss < frame_skip; ss++, nstep++)
        step_env<T, SPLIT>(m, xch, e, uc, hold && ss == 0, ss == frame_skip - 1, site, trip, cnt, nstep, vcache);
    }
    ctrl_gen<T>(spec, g, env, i, s.n, t, u);
#pragma unroll
    for (int k = 0; k < 5; k++) uc[k] = (T)u[k];
    clamp_ctrl(m, uc);   // rows keep the unclamped u, as the reference's dataset does
    if (rows && (t > t0 || t0 == 0)) {     // warp-uniform: the whole warp writes its 32 rows together
      ROW v[SO101_ROW];
#pragma unroll
      for (int k = 0; k < 5; k++) v[k] = (ROW)u[k];
#pragma unroll
      for (int k = 0; k < 3; k++) v[5 + k] = (ROW)(float)site[k];
#pragma unroll
      for (int k = 0; k < 5; k++) v[8 + k] = (ROW)(float)e.q[k];
      write_rows_warp<ROW>(rowbuf[SPLIT ? 0 : threadIdx.x >> 5], v, rows,
                           SPLIT ? (int64_t)blockIdx.x * 32 : (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u), s.n, Tn, t);
    }
  }
  if (active) store_env(s, i, e);
  else cnt = {0, 0, 0, 0};
  add_stats(stats, cnt);
}

// ---- the same rollout, time-sliced over a persistent grid (one-warp kernels, large batches) ---------------------------
// A resident block is latency bound, so a launch of `k_rollout` lasts ceil(groups / resident blocks) block times: 131072
// envs = 512 groups of 256 on 148 SMs are 3.46 waves and take 4 (a 13 % tail).  Here the unit of work is
// (group, time chunk): `gridDim` = the resident blocks, each walks units u = blockIdx, blockIdx + gridDim, ... in
// chunk-major order (u = chunk * groups + group), so every SM stays busy until the last chunk and the launch takes
// groups * T / resident block-steps.  Chunk c of a group continues from the state chunk c-1 stored (the continuation of
// `k_rollout`: row t0 already written, u_t0 regenerated), possibly on another SM: `progress[group]` counts the chunks
// done (release store after a fence, acquire load before the state is read through L2).  A unit's predecessor lies a
// whole pass over the groups back, i.e. it finished long ago; the wait exists for correctness, and is bounded.
__device__ __forceinline__ int ld_acquire_(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_(int32_t* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <typename T> SO101_DEV void load_env_cg(const StateView<T>& s, int64_t i, Env<T>& e) {   // through L2: another SM wrote it
#pragma unroll
  for (int k = 0; k < NV; k++) {
    e.q[k] = __ldcg(&s.base[(ROW_Q + k) * s.n + i]);
    e.qd[k] = __ldcg(&s.base[(ROW_QD + k) * s.n + i]);
    e.warm[k] = __ldcg(&s.base[(ROW_WARM + k) * s.n + i]);
    e.fa[k] = __ldcg(&s.base[(ROW_FA + k) * s.n + i]);
  }
  e.time = __ldcg(&s.base[ROW_TIME * s.n + i]);
  e.flags = __ldcg(&s.flags[i]);
}
template <typename T, typename ROW>
__global__ void __launch_bounds__(SO101_LB_THREADS, SO101_LB_BLOCKS)
k_rollout_sliced(const __grid_constant__ DevModel<T> m, StateView<T> s, DevSpec spec, int t0, int t1, int Tn, int frame_skip,
                 ROW* rows, uint32_t rflags, unsigned long long* stats, int tchunk, int32_t* progress, int32_t* fault,
                 const int32_t* perm, uint8_t* recent) {
  // perm (nullable): slot -> env.  A block pays for a table contact of any of its lanes (its warps meet at two barriers per
  // step), so between the time chunks of a long rollout the host regroups the envs that touched the table into the same
  // blocks (so101_capi.cu: regroup); recent[env] reports which envs did during this launch.
  __shared__ RowBufP<ROW> rowbuf[SO101_LB_THREADS / 32];
  SplitXch<T>& xch = *reinterpret_cast<SplitXch<T>*>(rowbuf);      // never touched by the one-warp path
  const int64_t ngroups = (s.n + blockDim.x - 1) / blockDim.x;
  const int nchunks = (t1 - t0 + tchunk - 1) / tchunk;
  const bool trip = m.ntrip > 0;
  const bool hold = rflags & SO101_ROLL_GRAVCOMP_HOLD;
  Counters cnt = {0, 0, 0, 0};
#pragma unroll 1
  for (int64_t unit = blockIdx.x; unit < ngroups * nchunks; unit += gridDim.x) {
    const int c = (int)(unit / ngroups);
    const int64_t grp = unit - (int64_t)c * ngroups;
    const int ut0 = t0 + c * tchunk, ut1 = ut0 + tchunk < t1 ? ut0 + tchunk : t1;
    const int64_t j = grp * blockDim.x + threadIdx.x;
    const bool active = j < s.n;
    const int64_t slot = active ? j : s.n - 1;                      // tail threads shadow a valid env and never store
    const int64_t i = perm ? (int64_t)perm[slot] : slot;
    const int64_t env = spec.env_offset + i;
    if (c > 0) {
      if (threadIdx.x == 0) {
        long long spins = 0;
        while (ld_acquire_(progress + grp) < c) {
          __nanosleep(200);
          if (++spins > (1ll << 24)) { atomicExch(fault, 1); break; }   // ~seconds: never in a correct schedule
        }
      }
      __syncthreads();
    }
    Env<T> e;
    if (c > 0 || (rflags & SO101_ROLL_NO_RESET)) {
      load_env_cg(s, i, e);
    } else {
      reset_env(m, e);
      double r[5];
      uniform5(spec.seed, env, 0, STREAM_RESET, r);
#pragma unroll
      for (int k = 0; k < 5; k++) e.q[k] = (T)muladd_rn(spec.reset_lo, spec.reset_hi - spec.reset_lo, r[k]);
    }
    const uint32_t flags_in = e.flags;
    e.flags &= ~(uint32_t)SO101_FLAG_CONTACT;                       // so that the flag tells what happened in THIS unit
    CtrlGen g;
    ctrl_init(spec, env, g);
    Counters ucnt = {0, 0, 0, 0};
    int32_t vcache[SO101_MAXTRIP];
#pragma unroll
    for (int k = 0; k < SO101_MAXTRIP; k++) vcache[k] = -1;
    vcache[SELF_BUDGET_SLOT] = vcache[SELF_REST_SLOT] = 0;
    T site[3];
    site_fk(m, e.q, site);
    double u[5];
    T uc[NV] = {T(0), T(0), T(0), T(0), T(0), T(0)};
#pragma unroll 1
    for (int t = ut0; t <= ut1; t++) {
      if (t > ut0) {
#pragma unroll 1
        for (int ss = 0; ss < frame_skip; ss++)
          step_env<T, false>(m, xch, e, uc, hold && ss == 0, ss == frame_skip - 1, site, trip, ucnt, 0, vcache);
      }
      ctrl_gen<T>(spec, g, env, i, s.n, t, u);
#pragma unroll
      for (int k = 0; k < 5; k++) uc[k] = (T)u[k];
      clamp_ctrl(m, uc);
      if (rows && (t > ut0 || ut0 == 0)) {
        ROW v[SO101_ROW];
#pragma unroll
        for (int k = 0; k < 5; k++) v[k] = (ROW)u[k];
#pragma unroll
        for (int k = 0; k < 3; k++) v[5 + k] = (ROW)(float)site[k];
#pragma unroll
        for (int k = 0; k < 5; k++) v[8 + k] = (ROW)(float)e.q[k];
        write_rows_warp_perm<ROW>(rowbuf[threadIdx.x >> 5], v, rows, i, active, Tn, t);
      }
    }
    if (active && recent) {   // a hull of this env was tested during the unit: one of its boxes was below the table top
      bool tested = (e.flags & SO101_FLAG_CONTACT) != 0;
#pragma unroll
      for (int k = 0; k <= SELF_PAIR_SLOT; k++) tested |= vcache[k] >= 0;   // ... or it took the self-collision test (outside the joint box)
      if (tested) recent[i] = 1;
    }
    e.flags |= flags_in;
    if (active) {
      store_env(s, i, e);
      cnt.steps += ucnt.steps; cnt.newton += ucnt.newton; cnt.lsevals += ucnt.lsevals; cnt.limsteps += ucnt.limsteps;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_(progress + grp, c + 1);
  }
  add_stats(stats, cnt);
}
template <typename T>
cudaError_t launch_rollout_sliced(const DevModel<T>& m, StateView<T> v, unsigned grid, int blk, cudaStream_t st, const DevSpec& ds,
                                  int t0, int t1, int Tn, int frame_skip, void* rows, bool rows_f32, uint32_t rflags,
                                  unsigned long long* stats, int tchunk, int32_t* progress, int32_t* fault,
                                  const int32_t* perm, uint8_t* recent) {
  if (rows_f32) k_rollout_sliced<T, float><<<grid, blk, 0, st>>>(m, v, ds, t0, t1, Tn, frame_skip, (float*)rows, rflags, stats, tchunk, progress, fault, perm, recent);
  else k_rollout_sliced<T, double><<<grid, blk, 0, st>>>(m, v, ds, t0, t1, Tn, frame_skip, (double*)rows, rflags, stats, tchunk, progress, fault, perm, recent);
  return cudaGetLastError();
}
template <typename T> int rollout_sliced_blocks_per_sm(int blk) {
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_rollout_sliced<T, double>, blk, 0) != cudaSuccess) return 0;
  int nf = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nf, k_rollout_sliced<T, float>, blk, 0) != cudaSuccess) return 0;
  return nb < nf ? nb : nf;
}

// B control sequences U[H][5][B] from one shared state; X[B][H+1][8] float32 observations
SO101_STEP_KERNEL2(T)
k_shoot(const __grid_constant__ DevModel<T> m, StateView<T> s, const __grid_constant__ State0 s0, const T* U, int H,
        int frame_skip, float* X, uint32_t rflags, unsigned long long* stats) {
  __shared__ XchStorage<T, SPLIT> xst;
  SplitXch<T>& xch = xch_of(xst);
  bool active, exit_block;
  const int64_t i = env_slot<T, SPLIT>(s, active, exit_block);
  if (exit_block) return;
  if (SPLIT && threadIdx.x >= 32) { helper_role<T>(m, xch, (int64_t)H * frame_skip, frame_skip); return; }
  Env<T> e;
  reset_env(m, e);
#pragma unroll
  for (int k = 0; k < NV; k++) { e.q[k] = (T)s0.v[k]; e.qd[k] = (T)s0.v[6 + k]; e.warm[k] = (T)s0.v[12 + k]; }
  Counters cnt = {0, 0, 0, 0};
  const bool trip = m.ntrip > 0;
  int32_t vcache[SO101_MAXTRIP];   // support vertex of each hull when it was last tested (see contact_geometry)
#pragma unroll
  for (int g = 0; g < SO101_MAXTRIP; g++) vcache[g] = -1;
  vcache[SELF_BUDGET_SLOT] = vcache[SELF_REST_SLOT] = 0;   // float bits: separation budgets of the self-collision test (so101_physics.cuh)
  const bool hold = rflags & SO101_ROLL_GRAVCOMP_HOLD;
  T site[3];
  site_fk(m, e.q, site);
  if (SPLIT) publish_state(xch, e);
  T uc[NV] = {T(0), T(0), T(0), T(0), T(0), T(0)};
  int64_t nstep = 0;
#pragma unroll 1
  for (int t = 0; t <= H; t++) {
    if (t > 0) {
#pragma unroll
      for (int k = 0; k < 5; k++) uc[k] = U[((int64_t)(t - 1) * 5 + k) * s.n + i];
      clamp_ctrl(m, uc);
#pragma unroll 1
      for (int ss = 0; ss < frame_skip; ss++, nstep++)
        step_env<T, SPLIT>(m, xch, e, uc, hold && ss == 0, ss == frame_skip - 1, site, trip, cnt, nstep, vcache);
    }
    if (active) {
      float* x = X + ((int64_t)i * (H + 1) + t) * SO101_NOBS;
#pragma unroll
      for (int k = 0; k < 3; k++) x[k] = (float)site[k];
#pragma unroll
      for (int k = 0; k < 5; k++) x[3 + k] = (float)e.q[k];
    }
  }
  if (active) store_env(s, i, e);
  else cnt = {0, 0, 0, 0};
  add_stats(stats, cnt);
}

// ---- launchers (declared in so101_launch.h; instantiated per translation unit by so101_kernels.cu) -----------------
template <typename T, bool SPLIT>
cudaError_t launch_step(const DevModel<T>& m, StateView<T> v, unsigned grid, int blk, cudaStream_t st, const T* ctrl,
                        int n_ctrl, int nsub, float* obs, unsigned long long* stats, uint32_t sflags) {
  k_step<T, SPLIT><<<grid, blk, 0, st>>>(m, v, ctrl, n_ctrl, nsub, obs, stats, sflags);
  return cudaGetLastError();
}
template <typename T, bool SPLIT>
cudaError_t launch_rollout(const DevModel<T>& m, StateView<T> v, unsigned grid, int blk, cudaStream_t st, const DevSpec& ds,
                           int t0, int t1, int Tn, int frame_skip, void* rows, bool rows_f32, uint32_t rflags,
                           unsigned long long* stats) {
  if (rows_f32) k_rollout<T, float, SPLIT><<<grid, blk, 0, st>>>(m, v, ds, t0, t1, Tn, frame_skip, (float*)rows, rflags, stats);
  else k_rollout<T, double, SPLIT><<<grid, blk, 0, st>>>(m, v, ds, t0, t1, Tn, frame_skip, (double*)rows, rflags, stats);
  return cudaGetLastError();
}
template <typename T, bool SPLIT>
cudaError_t launch_shoot(const DevModel<T>& m, StateView<T> v, unsigned grid, int blk, cudaStream_t st, const State0& s0,
                         const T* U, int H, int frame_skip, float* X, uint32_t rflags, unsigned long long* stats) {
  k_shoot<T, SPLIT><<<grid, blk, 0, st>>>(m, v, s0, U, H, frame_skip, X, rflags, stats);
  return cudaGetLastError();
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

