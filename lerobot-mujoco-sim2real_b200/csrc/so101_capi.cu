// so101_capi.cu — host side of the C ABI of include/so101_b200.h (sm_100a): handles, launch shapes, the host-buffer
// pipeline.  The stepping kernels live in so101_kernels.cuh and are compiled by so101_kernels.cu.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "so101_kernels.cuh"

// the stepping kernels are compiled in so101_kernels.cu (four translation units); none is instantiated here
SO101_LAUNCHERS(extern, double, false)
SO101_LAUNCHERS(extern, double, true)
SO101_LAUNCHERS(extern, float, false)
SO101_LAUNCHERS(extern, float, true)
SO101_SLICED_LAUNCHERS(extern, double)
SO101_SLICED_LAUNCHERS(extern, float)

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
  unsigned long long* stats;  // device, 4 counters
  // time-sliced rollout (k_rollout_sliced): chunks done per env group, fault flag, resident blocks of the kernel
  int32_t* sched_progress; size_t sched_groups; int32_t* sched_fault; int sched_resident;
  // regrouping of long rollouts (see rollout_range): slot -> env permutation, envs that touched the table in the last
  // chunk, the two fill counters of the partition
  int32_t* rg_perm; uint8_t* rg_recent; int32_t* rg_count;
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
  int opt_family, opt_block, opt_host_chunks, opt_host_even, opt_sliced, opt_regroup, opt_host_direct;
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

// partition of the envs for the next time chunk of a regrouped rollout: the envs that touched the table first.  The order
// inside the two classes depends on the scheduling of the atomics; no result does (an env's arithmetic does not depend on
// the lane or block it sits in, and its rows go to its own place).
__global__ void k_regroup(const uint8_t* recent, int64_t n, int32_t* perm, int32_t* count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (recent[i]) perm[atomicAdd(&count[0], 1)] = (int32_t)i;
  else perm[n - 1 - atomicAdd(&count[1], 1)] = (int32_t)i;
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
  }
  hostbuild::convert<float>(b->dm_d, b->dm_f);
  if (e != cudaSuccess) {
    if (b->owns_state) cudaFree(b->state);
    cudaFree(b->stats);
    for (int k = 0; k < 4; k++) cudaFree(b->hull_dev[k]);
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
  cudaFree(b->sched_progress); cudaFree(b->sched_fault);
  cudaFree(b->rg_perm); cudaFree(b->rg_recent); cudaFree(b->rg_count);
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
      b->sched_resident = 0;                              // resident blocks of the time-sliced kernel depend on the block size
      return SO101_OK;
    }
    case SO101_OPT_HOST_CHUNKS:
      if (value < 0 || value > So101Batch::MAXCHUNK) return fail(SO101_EINVAL, "host pipeline chunks must be 0 (auto) .. 12");
      b->opt_host_chunks = value;
      return SO101_OK;
    case SO101_OPT_HOST_EVEN:
      b->opt_host_even = value != 0;
      return SO101_OK;
    case SO101_OPT_HOST_DIRECT:
      if (value < 0 || value > 2) return fail(SO101_EINVAL, "direct host rows must be 0 (auto), 1 (whenever possible) or 2 (never)");
      b->opt_host_direct = value;
      return SO101_OK;
    case SO101_OPT_SELF_TEST:
      if (value < 0 || value > 1) return fail(SO101_EINVAL, "self-collision test must be 0 (box-box test) or 1 (joint box only)");
      b->dm_d.self_boxes = b->dm_f.self_boxes = value == 0;
      return SO101_OK;
    case SO101_OPT_REGROUP:
      if (value < 0 || value > 2) return fail(SO101_EINVAL, "regrouping must be 0 (auto), 1 (always, one-warp kernels) or 2 (never)");
      b->opt_regroup = value;
      return SO101_OK;
    case SO101_OPT_SLICED:
      if (value < 0 || value > 2) return fail(SO101_EINVAL, "sliced rollout must be 0 (auto), 1 (always, one-warp kernels) or 2 (never)");
      b->opt_sliced = value;
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

int so101_batch_step_flags(So101Batch* b, const void* ctrl, int n_ctrl, int n_substeps, void* obs, uint32_t flags,
                           void* stream) {
  if (!b) return fail(SO101_EINVAL, "null batch");
  if (n_ctrl < 0 || n_ctrl > NV || n_substeps < 0) return fail(SO101_EINVAL, "n_ctrl must be 0..6, n_substeps >= 0");
  if (flags & ~(uint32_t)SO101_ROLL_GRAVCOMP_HOLD) return fail(SO101_EINVAL, "step flags: only SO101_ROLL_GRAVCOMP_HOLD");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int blk; unsigned grid; bool split;
  if (b->dtype == SO101_F64) {
    StateView<double> v = step_view<double>(b, blk, grid, split);
    CUDA_TRY((split ? launch_step<double, true>(b->dm_d, v, grid, blk, st, (const double*)ctrl, n_ctrl, n_substeps, (float*)obs, b->stats, flags)
                   : launch_step<double, false>(b->dm_d, v, grid, blk, st, (const double*)ctrl, n_ctrl, n_substeps, (float*)obs, b->stats, flags)));
  } else {
    StateView<float> v = step_view<float>(b, blk, grid, split);
    CUDA_TRY((split ? launch_step<float, true>(b->dm_f, v, grid, blk, st, (const float*)ctrl, n_ctrl, n_substeps, (float*)obs, b->stats, flags)
                   : launch_step<float, false>(b->dm_f, v, grid, blk, st, (const float*)ctrl, n_ctrl, n_substeps, (float*)obs, b->stats, flags)));
  }
  return SO101_OK;
}
int so101_batch_step(So101Batch* b, const void* ctrl, int n_ctrl, int n_substeps, void* obs, void* stream) {
  return so101_batch_step_flags(b, ctrl, n_ctrl, n_substeps, obs, 0, stream);
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

// One time chunk (t0, t1] of a rollout
static int rollout_range(So101Batch* b, const So101CtrlSpec* spec, int t0, int t1, int T, int frame_skip, void* rows,
                         uint32_t flags, void* stream) {
  if (!b || !spec) return fail(SO101_EINVAL, "null argument");
  if (T < 0 || frame_skip < 1) return fail(SO101_EINVAL, "T must be >= 0 and frame_skip >= 1");
  if (spec->kind < SO101_CTRL_RANDOM || spec->kind > SO101_CTRL_TENSOR) return fail(SO101_EINVAL, "bad control kind");
  if (spec->kind == SO101_CTRL_TENSOR && !spec->u) return fail(SO101_EINVAL, "SO101_CTRL_TENSOR needs spec->u");
  if (spec->kind == SO101_CTRL_CHIRP && spec->t_total <= 0) return fail(SO101_EINVAL, "chirp needs t_total > 0");
  DeviceGuard g(b->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DevSpec ds;
  ds.kind = spec->kind; ds.t_total = spec->t_total; ds.seed = spec->seed; ds.env_offset = spec->env_offset;
  ds.amp = spec->amp; ds.freq_lo = spec->freq_lo; ds.freq_hi = spec->freq_hi;
  ds.reset_lo = spec->reset_lo; ds.reset_hi = spec->reset_hi; ds.u = spec->u;
  int blk; unsigned grid; bool split;
  const bool r32 = flags & SO101_ROLL_ROWS_F32;
  {
    // Large batches whose groups do not fill whole waves of resident blocks: time-sliced persistent launch (k_rollout_sliced).
    // Long rollouts of large batches with table contact: the same kernel, launched per chunk of control steps, with the envs
    // regrouped between the chunks so that the ones touching the table share blocks (a block pays for a contact of any of
    // its lanes: its warps meet at two barriers per step; measured x1.5 on a chirp T=200 set with 4 % of the envs in contact).
    if (b->dtype == SO101_F64) step_view<double>(b, blk, grid, split); else step_view<float>(b, blk, grid, split);
    if (!split && !b->opt_block && b->opt_sliced != 2 && t1 - t0 >= 2) {
      if (!b->sched_resident) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, b->device));
        const int bps = b->dtype == SO101_F64 ? rollout_sliced_blocks_per_sm<double>(blk) : rollout_sliced_blocks_per_sm<float>(blk);
        b->sched_resident = bps > 0 ? bps * prop.multiProcessorCount : -1;
      }
      const int64_t ngroups = (b->n + blk - 1) / blk;
      const int64_t res = b->sched_resident;
      const int64_t waves_up = res > 0 ? (ngroups + res - 1) / res : 0;
      // gain of even waves over whole ones; below 3 % (or a single wave) the plain launch is as good
      const bool worth = res > 0 && ngroups > res && (double)ngroups / (double)(waves_up * res) < 0.97;
      const bool regroup = res > 0 && b->dm_d.con_enabled && b->opt_regroup != 2 && ngroups > 1 &&
                           (b->opt_regroup == 1 || (t1 - t0 >= 40 && ngroups >= 2 * res));
      if (worth || regroup || (b->opt_sliced == 1 && res > 0 && ngroups > 1)) {
        if ((size_t)ngroups > b->sched_groups) {
          cudaFree(b->sched_progress);
          b->sched_progress = nullptr; b->sched_groups = 0;
          CUDA_TRY(cudaMalloc(&b->sched_progress, (size_t)ngroups * sizeof(int32_t)));
          b->sched_groups = (size_t)ngroups;
        }
        if (!b->sched_fault) { CUDA_TRY(cudaMalloc(&b->sched_fault, sizeof(int32_t))); CUDA_TRY(cudaMemset(b->sched_fault, 0, sizeof(int32_t))); }
        if (regroup && !b->rg_perm) {
          CUDA_TRY(cudaMalloc(&b->rg_perm, (size_t)b->n * sizeof(int32_t)));
          CUDA_TRY(cudaMalloc(&b->rg_recent, (size_t)b->n));
          CUDA_TRY(cudaMalloc(&b->rg_count, 2 * sizeof(int32_t)));
        }
        const int rc = regroup ? 20 : t1 - t0;            // control steps between two regroupings
        for (int c0 = t0; c0 < t1; c0 += rc) {
          const int c1 = c0 + rc < t1 ? c0 + rc : t1;
          // units per resident block: enough to even out the waves; per-unit overhead is a state round trip
          int nchunks = (int)(((regroup ? 6 : 24) * res + ngroups - 1) / ngroups);
          if (nchunks > c1 - c0) nchunks = c1 - c0;
          if (nchunks < 1) nchunks = 1;
          if (!regroup && nchunks < 2) nchunks = 2 < c1 - c0 ? 2 : c1 - c0;
          const int tchunk = (c1 - c0 + nchunks - 1) / nchunks;
          const uint32_t f = c0 > t0 ? (flags | SO101_ROLL_NO_RESET) : flags;
          const int32_t* perm = (regroup && c0 > t0) ? b->rg_perm : nullptr;
          uint8_t* recent = (regroup && c1 < t1) ? b->rg_recent : nullptr;
          CUDA_TRY(cudaMemsetAsync(b->sched_progress, 0, (size_t)ngroups * sizeof(int32_t), st));
          if (recent) CUDA_TRY(cudaMemsetAsync(recent, 0, (size_t)b->n, st));
          const int64_t units = ngroups * ((c1 - c0 + tchunk - 1) / tchunk);
          const unsigned pgrid = (unsigned)(res < units ? res : units);
          if (b->dtype == SO101_F64)
            CUDA_TRY(launch_rollout_sliced<double>(b->dm_d, view<double>(b), pgrid, blk, st, ds, c0, c1, T, frame_skip, rows, r32, f, b->stats, tchunk, b->sched_progress, b->sched_fault, perm, recent));
          else
            CUDA_TRY(launch_rollout_sliced<float>(b->dm_f, view<float>(b), pgrid, blk, st, ds, c0, c1, T, frame_skip, rows, r32, f, b->stats, tchunk, b->sched_progress, b->sched_fault, perm, recent));
          if (recent) {
            CUDA_TRY(cudaMemsetAsync(b->rg_count, 0, 2 * sizeof(int32_t), st));
            k_regroup<<<(unsigned)((b->n + 255) / 256), 256, 0, st>>>(recent, b->n, b->rg_perm, b->rg_count);
            CUDA_TRY(cudaGetLastError());
          }
        }
        return SO101_OK;
      }
    }
  }
  if (b->dtype == SO101_F64) {
    StateView<double> v = step_view<double>(b, blk, grid, split);
    CUDA_TRY((split ? launch_rollout<double, true>(b->dm_d, v, grid, blk, st, ds, t0, t1, T, frame_skip, rows, r32, flags, b->stats)
                   : launch_rollout<double, false>(b->dm_d, v, grid, blk, st, ds, t0, t1, T, frame_skip, rows, r32, flags, b->stats)));
  } else {
    StateView<float> v = step_view<float>(b, blk, grid, split);
    CUDA_TRY((split ? launch_rollout<float, true>(b->dm_f, v, grid, blk, st, ds, t0, t1, T, frame_skip, rows, r32, flags, b->stats)
                   : launch_rollout<float, false>(b->dm_f, v, grid, blk, st, ds, t0, t1, T, frame_skip, rows, r32, flags, b->stats)));
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
  // Rows straight into the caller's buffer: pinned host memory is device-accessible under unified addressing, and the row
  // writer's stores (a warp's 32 rows, consecutive words) go over PCIe while the launch computes - 43 MB in 7 ms is a
  // fraction of the link.  No staging copy, no download phase, and no chunk boundaries (a boundary costs what the slowest
  // block of each chunk costs).  Large datasets keep the copy-engine pipeline below: there the link is the bottleneck and
  // stores that wait for it would stall the SMs (measured, tools/e2e_probe.py: the stores sustain ~16 GB/s; 43 / 71 / 89 MB
  // of rows over 1000 / 200 / 500 physics steps: 8.34 -> 7.92, 4.55 -> 4.36, 9.67 -> 8.98 ms; 143 / 286 MB over 200 steps:
  // 6.6 -> 9.2, 12.5 -> 19.2 ms).
  void* rows_direct = nullptr;
  if (b->opt_host_direct != 2 && (b->opt_host_direct == 1 || (rb < ((size_t)128 << 20) && (int64_t)T * frame_skip >= 200))) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, rows_host) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
      rows_direct = pa.devicePointer;
    else
      cudaGetLastError();   // pageable memory: not an error, the staged path takes it
  }
  if (!rows_direct && (rc = grow(&b->rows_stage, &b->rows_stage_bytes, rb))) return rc;
  // chunking: worth it only when there is something to overlap
  // A chunk boundary is not free: every launch lasts as long as its slowest block, and with table contact the slow block
  // is a different one in every chunk, so the sum of the chunks exceeds the undivided launch (measured, 4096 envs x 100
  // steps, 43 MB of rows: 1 / 2 / 3 / 4 / 8 / 12 chunks 8.67 / 8.33 / 8.53 / 8.61 / 8.65 / 8.75 ms, tools/e2e_probe.py).
  // Two chunks while the copies are small beside the compute; more only where the download dominates (>= 128 MB of
  // rows) and has to start early.
  int nchunk = 1;
  if (T >= 8 && rb >= ((size_t)4 << 20)) {
    nchunk = 2;
    if (rb >= ((size_t)128 << 20)) {
      nchunk = (int)(rb >> 25);                 // >= 32 MB of rows per chunk
      if (nchunk > So101Batch::MAXCHUNK) nchunk = So101Batch::MAXCHUNK;
      if (nchunk < 3) nchunk = 3;
    }
  }
  if (rows_direct) nchunk = 1;
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
  if (rows_direct && tensor && !b->opt_host_chunks && T >= 20) {
    // direct rows leave one copy exposed, the upload of the control tensor: start on the first tenth of it and let the rest
    // arrive while that computes.  The one boundary comes early, before the blocks have drifted apart.
    nchunk = 2;
    tb[0] = 0; tb[1] = T / 10 < 2 ? 2 : T / 10; tb[2] = T;
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
    if ((rc = rollout_range(b, &dspec, tb[c], tb[c + 1], T, frame_skip, rows_direct ? rows_direct : b->rows_stage, f, stream))) { drain(); return rc; }
    if (rows_direct) continue;
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
  if (b->dtype == SO101_F64) {
    StateView<double> v = step_view<double>(b, blk, grid, split);
    CUDA_TRY((split ? launch_shoot<double, true>(b->dm_d, v, grid, blk, st, s0, (const double*)U, H, frame_skip, (float*)X, flags, b->stats)
                   : launch_shoot<double, false>(b->dm_d, v, grid, blk, st, s0, (const double*)U, H, frame_skip, (float*)X, flags, b->stats)));
  } else {
    StateView<float> v = step_view<float>(b, blk, grid, split);
    CUDA_TRY((split ? launch_shoot<float, true>(b->dm_f, v, grid, blk, st, s0, (const float*)U, H, frame_skip, (float*)X, flags, b->stats)
                   : launch_shoot<float, false>(b->dm_f, v, grid, blk, st, s0, (const float*)U, H, frame_skip, (float*)X, flags, b->stats)));
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
  int32_t fault = 0;
  if (b->sched_fault) CUDA_TRY(cudaMemcpyAsync(&fault, b->sched_fault, sizeof fault, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (fault) return fail(SO101_ECUDA, "time-sliced rollout: a unit waited for its predecessor beyond the bound (scheduling fault)");
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
