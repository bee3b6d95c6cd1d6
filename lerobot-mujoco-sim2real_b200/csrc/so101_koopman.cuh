// so101_koopman.cuh - scoring of control sequences under the reference's lifted linear (Koopman) model.
//
// the reference's MPC [REF control/MPC_Controler.py:65-98] rolls a lifted state
//   z_{t+1} = A z_t + B u_t        (A = net.lA.weight [nz x nz], B = net.lB.weight [nz x nu]; KoopmanBase.py:49-57)
// over a horizon and minimises  sum_t (z_{t+1} - zref_t)' Q (z_{t+1} - zref_t) + u_t' R u_t  with Q = 50 I, R = 0.5 I
// (`state_full`, MPC_Controler.py:38-40; the `args.model == 'IBKN' or 'IKN'` test there is always true).
// This kernel evaluates that model and that cost for n control sequences at once - the model-side counterpart of
// `k_shoot`, which rolls the same sequences through the physics: together they are the "Koopman_MPC evaluation
// workload" of BASELINE config 5 (model prediction vs simulated outcome per candidate sequence).
// The lift z0 = [x, encoder(x)] is a 5-layer MLP evaluated once per state on the host side (koopman.py).
//
// One thread per sequence, z in registers/local memory, A, B, z0 and the reference staged in shared memory (all
// threads read the same element: broadcast).  Arithmetic in double whatever the dtype of U.
// Included at the end of so101_capi.cu (same translation unit: shares fail()/CUDA_TRY).
#pragma once

constexpr int KOOP_MAXZ = 64, KOOP_MAXU = 8;

template <typename T>
__global__ void __launch_bounds__(128)
k_koopman_score(const double* __restrict__ pack, int nz, int nu, int H, int has_ref, double qw, double rw, const T* U,
                int64_t n, int nobs, float* Xhat, double* cost) {
  extern __shared__ double sh[];
  double* A = sh;                       // [nz][nz]
  double* B = A + nz * nz;              // [nz][nu]
  double* z0 = B + nz * nu;             // [nz]
  double* zref = z0 + nz;               // [H][nz]
  const int total = nz * nz + nz * nu + nz + (has_ref ? H * nz : 0);
  for (int i = threadIdx.x; i < total; i += blockDim.x) sh[i] = pack[i];
  __syncthreads();
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  double z[KOOP_MAXZ], zn[KOOP_MAXZ], u[KOOP_MAXU];
  for (int i = 0; i < nz; i++) z[i] = z0[i];
  if (Xhat) for (int i = 0; i < nobs; i++) Xhat[(b * (H + 1)) * nobs + i] = (float)z[i];
  double c = 0.0;
  for (int t = 0; t < H; t++) {
    double uu = 0.0;
    for (int k = 0; k < nu; k++) { u[k] = (double)U[((int64_t)t * nu + k) * n + b]; uu = fma(u[k], u[k], uu); }
    double ee = 0.0;
    for (int i = 0; i < nz; i++) {
      double s = 0.0;
      const double* Ai = A + i * nz;
#pragma unroll 4
      for (int j = 0; j < nz; j++) s = fma(Ai[j], z[j], s);
      for (int k = 0; k < nu; k++) s = fma(B[i * nu + k], u[k], s);
      zn[i] = s;
      if (has_ref) { const double d = s - zref[t * nz + i]; ee = fma(d, d, ee); }
    }
    c += qw * ee + rw * uu;
    for (int i = 0; i < nz; i++) z[i] = zn[i];
    if (Xhat) for (int i = 0; i < nobs; i++) Xhat[(b * (H + 1) + t + 1) * nobs + i] = (float)z[i];
  }
  if (cost) cost[b] = c;
}

extern "C" int so101_koopman_score(const double* A, const double* B, int nz, int nu, const double* z0,
                                   const double* zref, double q_weight, double r_weight, const void* U, int H,
                                   int64_t n, int dtype, int device, int nobs, void* Xhat, void* cost, void* stream) {
  if (!A || !B || !z0 || (!U && H > 0) || n <= 0) return fail(SO101_EINVAL, "null argument");
  if (nz < 1 || nz > KOOP_MAXZ || nu < 1 || nu > KOOP_MAXU || H < 0 || nobs < 0 || nobs > nz)
    return fail(SO101_EINVAL, "koopman_score: need 1 <= nz <= 64, 1 <= nu <= 8, H >= 0, nobs <= nz");
  if (dtype != SO101_F64 && dtype != SO101_F32) return fail(SO101_EINVAL, "dtype must be SO101_F64 or SO101_F32");
  if (so101_device_count() <= 0) return fail(SO101_ENODEVICE, "no CUDA device visible: this library has no CPU fallback");
  DeviceGuard g(device);
  if (!g.ok) return fail(SO101_ECUDA, "cudaSetDevice failed");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t cnt = (size_t)nz * nz + (size_t)nz * nu + nz + (zref ? (size_t)H * nz : 0);
  const size_t bytes = cnt * sizeof(double);
  if (bytes > 200 * 1024) return fail(SO101_EINVAL, "koopman_score: model + reference exceed 200 KB of shared memory");
  std::vector<double> host(cnt);
  size_t o = 0;
  std::memcpy(&host[o], A, sizeof(double) * nz * nz); o += (size_t)nz * nz;
  std::memcpy(&host[o], B, sizeof(double) * nz * nu); o += (size_t)nz * nu;
  std::memcpy(&host[o], z0, sizeof(double) * nz); o += nz;
  if (zref) std::memcpy(&host[o], zref, sizeof(double) * H * nz);
  double* pack = nullptr;
  CUDA_TRY(cudaMallocAsync(&pack, bytes, st));
  // every exit below releases `pack` (stream ordered)
  cudaError_t err = cudaMemcpyAsync(pack, host.data(), bytes, cudaMemcpyHostToDevice, st);   // pageable source: staged before return
  const int blk = 128;
  const unsigned grid = (unsigned)((n + blk - 1) / blk);
  if (err == cudaSuccess && bytes > 48 * 1024)
    err = dtype == SO101_F64
              ? cudaFuncSetAttribute(k_koopman_score<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)
              : cudaFuncSetAttribute(k_koopman_score<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (err == cudaSuccess) {
    if (dtype == SO101_F64)
      k_koopman_score<double><<<grid, blk, bytes, st>>>(pack, nz, nu, H, zref != nullptr, q_weight, r_weight,
                                                        (const double*)U, n, nobs, (float*)Xhat, (double*)cost);
    else
      k_koopman_score<float><<<grid, blk, bytes, st>>>(pack, nz, nu, H, zref != nullptr, q_weight, r_weight,
                                                       (const float*)U, n, nobs, (float*)Xhat, (double*)cost);
    err = cudaGetLastError();
  }
  if (err != cudaSuccess) {
    cudaFreeAsync(pack, st);
    return fail(SO101_ECUDA, std::string("koopman_score: ") + cudaGetErrorString(err));
  }
  CUDA_TRY(cudaFreeAsync(pack, st));
  return SO101_OK;
}

// ======================================================================================================================
// The reference's MPC loop body for a batch of environments: lift, gain product, clip.
//
// Reference, per frame [REF Koopman_MPC.py:197-222, control/MPC_Controler.py:143-152, 154-166]:
//   z0 = Psi(state) = [x, encoder(x)]      encoder = MLP 8 -> 64 -> 64 -> 64 -> 64 -> 24, ReLU between layers
//                                          [REF models/KoopmanBase.py:12-47, args.py:103]
//   zref_t = Psi(state_all_ref[k+1+t]), t = 0..H-1 (rows past the end of the trajectory stay ZERO in lifted space)
//   u_opt = argmin of the MPC problem  ('mpc': over u; 'delta_mpc', the default [REF args.py:75]: over delta_u with
//           u_t = u_prev + sum_{s<=t} delta_u_s and the cost on delta_u)      [REF MPC_Controler.py:65-141]
//   u0 = u_opt[0] + u_eso(=0) + u_prev ;  a = clip(u0, -0.5, 0.5) ;  u_prev <- u0 (runMPC overwrites get_control's
//           `u_prev = a` with the unclipped u, in both modes [REF Koopman_MPC.py:217])
// The problem is an unconstrained quadratic (linear model), so u_opt[0] = Kz z0 + sum_t Kr_t zref_t + Ku u_prev with
// gain matrices that depend on the model only (host: koopman.py mpc_gains).  The reference part sum_t Kr_t zref_t does
// not depend on the state: so101_koopman_feedforward lifts every reference row ONCE and folds the windows into
// uff[n][P][nu]; the per-frame kernel then lifts the state, adds the three terms, clips and writes the control rows the
// stepper takes.
//
// MLP kernel: a 256-thread block works on tiles of 64 rows (lane = 2 rows); warp w computes output neurons
// [w*JT, (w+1)*JT) of the current layer for all 64 rows; activations ping-pong between two shared buffers [k][row]
// (conflict-free loads), the transposed weights of all layers sit in shared memory (117 KB for the reference's encoder)
// and are read as broadcasts.  FP64 throughout, as the reference's DoubleTensor network.  DFMA-bound in principle
// (14.3 k FMA per lift); per k-step a warp issues 2 activation loads + JT/2 broadcast double2 loads for 2*JT DFMA.
// ======================================================================================================================
constexpr int KM_MAXL = 8, KM_ROWS = 64, KM_THREADS = 256, KM_MAXW = 64, KM_MAXX = 16;
constexpr int KM_PAD = 4, KM_AS = KM_MAXW + KM_PAD;   // activations [row][KM_AS], weights [k][dout + KM_PAD]: the strides (= 4 mod 16
                                                      // doubles) make the 8-byte fragment loads of mma.m8n8k4 conflict-free

struct So101Koopman {
  int device, sms, n_layers, dims[KM_MAXL + 1];
  bool attr_lift[2], attr_mpc[4];                // dynamic shared memory limit raised (per kernel instantiation)
  int x_dim, nz;
  size_t woff[KM_MAXL], boff[KM_MAXL], wcount;   // offsets (doubles) of Wt[l] ([din][dout], transposed) and b[l] in `weights`
  double* weights;                               // device
  // gains (so101_koopman_set_gains)
  int H, nu;
  double* gains;                                 // device: Kz [nu][nz] | Ku [nu][nu] | Kr [nu][H][nz]
};

struct KmLayers {
  int n_layers, dims[KM_MAXL + 1];
  int woff[KM_MAXL], boff[KM_MAXL], wcount;
};

// encoder of the 64 rows whose inputs sit in act0[row][k] (row stride KM_AS, k < dims[0] rounded up to 4 with zeros);
// returns the buffer holding the last layer's output in the same layout.  The layers are GEMMs [64 rows x din] x
// [din x dout] in FP64: they run on the tensor cores as mma.sync.m8n8k4.f64 (DMMA; tcgen05 has no FP64 kind).  A warp
// owns 32 rows x 16 neurons = 4 x 2 accumulator tiles of 8 x 8 (16 doubles per lane); per k-step of 4 it loads 4 A
// fragments (activations) and 2 B fragments (weights), one double per lane each, for 8 DMMA = 2048 FMA.  (The first
// version of this kernel did the same tile with DFMA: 64 DFMA + 16 shared loads for the same 2048 FMA; 35 % of the FP64
// pipe with the shared-memory path 67 % busy, profiles/r2_koopman_mpc_ncu_summary.json.)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double* km_encode(const KmLayers& L, const double* __restrict__ w, double* act0, double* act1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;                   // fragment coordinates: groupID, threadID_in_group
  const int row0 = (warp >> 2) * 32;                       // this warp's 32 rows
  const int col0 = (warp & 3) * 16;                        // this warp's 16 neurons
  double* in = act0;
  double* out = act1;
  for (int l = 0; l < L.n_layers; l++) {
    const int din = (L.dims[l] + 3) & ~3, dout = L.dims[l + 1], ws = dout + KM_PAD;
    const double* Wt = w + L.woff[l];
    const double* bias = w + L.boff[l];
    const bool relu = l != L.n_layers - 1;
    const bool jb_on[2] = {col0 < dout, col0 + 8 < dout};  // widths are multiples of 8 here (checked at creation)
    if (jb_on[0]) {
      double acc[4][2][2];
#pragma unroll
      for (int jb = 0; jb < 2; jb++) {
        const double b0 = jb_on[jb] ? bias[col0 + jb * 8 + 2 * t] : 0.0, b1 = jb_on[jb] ? bias[col0 + jb * 8 + 2 * t + 1] : 0.0;
#pragma unroll
        for (int rb = 0; rb < 4; rb++) { acc[rb][jb][0] = b0; acc[rb][jb][1] = b1; }
      }
#pragma unroll 2
      for (int k0 = 0; k0 < din; k0 += 4) {
        double af[4], bf[2];
#pragma unroll
        for (int rb = 0; rb < 4; rb++) af[rb] = in[(row0 + rb * 8 + g) * KM_AS + k0 + t];       // A[row = g][k = t]
#pragma unroll
        for (int jb = 0; jb < 2; jb++) bf[jb] = jb_on[jb] ? Wt[(k0 + t) * ws + col0 + jb * 8 + g] : 0.0;   // B[k = t][col = g]
#pragma unroll
        for (int rb = 0; rb < 4; rb++)
#pragma unroll
          for (int jb = 0; jb < 2; jb++) dmma884(acc[rb][jb][0], acc[rb][jb][1], af[rb], bf[jb]);
      }
#pragma unroll
      for (int rb = 0; rb < 4; rb++)
#pragma unroll
        for (int jb = 0; jb < 2; jb++) {
          if (!jb_on[jb]) continue;
          double2 y;                                                                            // C[row = g][col = 2t, 2t+1]
          y.x = relu ? fmax(acc[rb][jb][0], 0.0) : acc[rb][jb][0];
          y.y = relu ? fmax(acc[rb][jb][1], 0.0) : acc[rb][jb][1];
          *reinterpret_cast<double2*>(out + (row0 + rb * 8 + g) * KM_AS + col0 + jb * 8 + 2 * t) = y;
        }
    }
    __syncthreads();
    double* tmp = in; in = out; out = tmp;
  }
  return in;
}

// X: layout 0 = rows [n][ldx] (first x_dim columns), 1 = structure of arrays [x_dim][n]
template <typename TX>
__device__ __forceinline__ void km_load_tile(const TX* X, int layout, int64_t ldx, int64_t n, int64_t row0, int x_dim,
                                             double* xs /* [x_dim][KM_ROWS] */, double* act /* [KM_ROWS][KM_AS] */) {
  const int xpad = (x_dim + 3) & ~3;
  for (int i = threadIdx.x; i < xpad * KM_ROWS; i += KM_THREADS) {
    int k, r;
    if (layout == 0) { r = i / xpad; k = i - r * xpad; } else { k = i / KM_ROWS; r = i - k * KM_ROWS; }
    const int64_t row = row0 + r;
    double v = 0.0;
    if (row < n && k < x_dim) v = (double)(layout == 0 ? X[row * ldx + k] : X[(int64_t)k * n + row]);
    if (k < x_dim) xs[k * KM_ROWS + r] = v;
    act[r * KM_AS + k] = v;
  }
}

// z = [x | encoder(x)] for n rows -> Z [n][nz]
template <typename TX>
__global__ void __launch_bounds__(KM_THREADS, 1)
k_koopman_lift(const __grid_constant__ KmLayers L, const double* __restrict__ wg, const TX* X, int layout, int64_t ldx,
               int64_t n, double* Z) {
  extern __shared__ double sh[];
  double* w = sh;
  double* xs = w + L.wcount;                       // [x_dim][64]: the inputs are kept (they are the first nz coordinates)
  double* act0 = xs + KM_MAXX * KM_ROWS;
  double* act1 = act0 + KM_ROWS * KM_AS;
  for (int i = threadIdx.x; i < L.wcount; i += KM_THREADS) w[i] = wg[i];
  const int x_dim = L.dims[0], enc = L.dims[L.n_layers], nz = x_dim + enc;
  const int64_t ntile = (n + KM_ROWS - 1) / KM_ROWS;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t row0 = tile * KM_ROWS;
    __syncthreads();
    km_load_tile(X, layout, ldx, n, row0, x_dim, xs, act0);
    __syncthreads();
    const double* h = km_encode(L, w, act0, act1);
    for (int i = threadIdx.x; i < nz * KM_ROWS; i += KM_THREADS) {
      const int r = i / nz, c = i - r * nz;
      if (row0 + r < n) Z[(row0 + r) * nz + c] = c < x_dim ? xs[c * KM_ROWS + r] : h[r * KM_AS + c - x_dim];
    }
  }
}

// uff[e][k][i] = sum_{t < H, k+1+t < P} sum_c Kr[i][t][c] Z[e][k+1+t][c] for the envs of one chunk: per env a banded
// product [P frames x (H nz)] x [(H nz) x nu], on the FP64 tensor cores as well.  A warp takes 8 consecutive frames of one
// env: A fragment (frame g, column t of the k-step) = Z[e][k0 + g + 1 + tw][c0 + t] straight from global memory (rows past
// the end of the trajectory are zero in lifted space), B fragment = the reference gain in shared memory as [tw][c][8]
// (the nu outputs padded to 8: the 32 lanes of a load hit 32 consecutive doubles), 8 x 8 accumulator tile, nu columns kept.
__global__ void __launch_bounds__(128)
k_koopman_window(const double* __restrict__ Z, const double* __restrict__ gains, int nz, int nu, int H, int64_t ne, int P,
                 double* uff) {
  extern __shared__ double sh[];                     // KrS[H][nz][8]
  const double* Kr_g = gains + (size_t)nu * nz + (size_t)nu * nu;      // [nu][H][nz]
  for (int i = threadIdx.x; i < H * nz * 8; i += blockDim.x) {
    const int o = i & 7, c = (i >> 3) % nz, tw = (i >> 3) / nz;
    sh[i] = o < nu ? Kr_g[((size_t)o * H + tw) * nz + c] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t nblk = (P + 7) / 8;
  const int64_t ntask = ne * nblk, nwarp = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); task < ntask; task += nwarp) {
    const int64_t e = task / nblk;
    const int k0 = (int)(task - e * nblk) * 8;
    const double* Ze = Z + (size_t)e * P * nz;
    double c0 = 0.0, c1 = 0.0;
    for (int tw = 0; tw < H; tw++) {
      const int row = k0 + g + 1 + tw;
      const bool live = row < P;
      const double* zr = Ze + (size_t)(live ? row : 0) * nz;
      const double* kb = sh + (size_t)tw * nz * 8;
      for (int cc = 0; cc < nz; cc += 4) {           // nz is a multiple of 4 (checked by the caller)
        const double af = live ? zr[cc + t] : 0.0;
        const double bf = kb[(cc + t) * 8 + g];
        dmma884(c0, c1, af, bf);
      }
    }
    const int k = k0 + g;
    if (k < P) {
      double* o = uff + ((size_t)e * P + k) * nu;
      if (2 * t < nu) o[2 * t] = c0;
      if (2 * t + 1 < nu) o[2 * t + 1] = c1;
    }
  }
}

// One MPC frame for n envs: lift the observation, u_opt0 = Kz z0 + uff + Ku u_prev, u0 = u_opt0 + u_prev, a = clip(u0),
// u_prev <- u0; a goes to the control rows ctrl[nu][n] (the stepper's layout and dtype) and, if asked, to a_out[n][nu].
template <typename TX, typename TC>
__global__ void __launch_bounds__(KM_THREADS, 1)
k_koopman_mpc(const __grid_constant__ KmLayers L, const double* __restrict__ wg, const double* __restrict__ gains, int nu,
              const TX* X, int layout, int64_t ldx, int64_t n, const double* uff, int64_t uff_stride, double* u_prev,
              TC* ctrl, double* a_out, double clip) {
  extern __shared__ double sh[];
  double* w = sh;
  double* xs = w + L.wcount;
  double* act0 = xs + KM_MAXX * KM_ROWS;
  double* act1 = act0 + KM_ROWS * KM_AS;
  double* g = act1 + KM_ROWS * KM_AS;              // Kz [nu][nz] | Ku [nu][nu]
  const int x_dim = L.dims[0], enc = L.dims[L.n_layers], nz = x_dim + enc;
  double* park = g + nu * nz + nu * nu;            // u0 of the tile [nu][KM_ROWS], until every component has read u_prev
  for (int i = threadIdx.x; i < L.wcount; i += KM_THREADS) w[i] = wg[i];
  for (int i = threadIdx.x; i < nu * nz + nu * nu; i += KM_THREADS) g[i] = gains[i];
  const int64_t ntile = (n + KM_ROWS - 1) / KM_ROWS;
  for (int64_t tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const int64_t row0 = tile * KM_ROWS;
    __syncthreads();
    km_load_tile(X, layout, ldx, n, row0, x_dim, xs, act0);
    __syncthreads();
    const double* h = km_encode(L, w, act0, act1);
    for (int i = threadIdx.x; i < nu * KM_ROWS; i += KM_THREADS) {
      const int c = i / KM_ROWS, r = i - c * KM_ROWS;       // control component c of row r (consecutive threads: rows)
      const int64_t row = row0 + r;
      if (row >= n) continue;
      double u = uff ? uff[row * uff_stride + c] : 0.0;
      const double* kz = g + c * nz;
      for (int k = 0; k < x_dim; k++) u = fma(kz[k], xs[k * KM_ROWS + r], u);
      for (int k = 0; k < enc; k++) u = fma(kz[x_dim + k], h[r * KM_AS + k], u);
      const double* ku = g + nu * nz + c * nu;
      for (int j = 0; j < nu; j++) u = fma(ku[j], u_prev[(int64_t)j * n + row], u);
      const double u0 = u + u_prev[(int64_t)c * n + row];
      const double a = fmin(fmax(u0, -clip), clip);
      // every component reads the whole u_prev of its row before any is overwritten: the write happens after the barrier
      park[c * KM_ROWS + r] = u0;
      ctrl[(int64_t)c * n + row] = (TC)a;
      if (a_out) a_out[row * nu + c] = a;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nu * KM_ROWS; i += KM_THREADS) {
      const int c = i / KM_ROWS, r = i - c * KM_ROWS;
      if (row0 + r < n) u_prev[(int64_t)c * n + row0 + r] = park[c * KM_ROWS + r];
    }
  }
}

static size_t km_smem_bytes(const So101Koopman* k, bool mpc) {
  size_t d = k->wcount + (size_t)KM_MAXX * KM_ROWS + 2 * (size_t)KM_ROWS * KM_AS;
  if (mpc) d += (size_t)k->nu * k->nz + (size_t)k->nu * k->nu + (size_t)k->nu * KM_ROWS;
  return d * sizeof(double);
}
static KmLayers km_layers(const So101Koopman* k) {
  KmLayers L;
  L.n_layers = k->n_layers;
  for (int i = 0; i <= k->n_layers; i++) L.dims[i] = k->dims[i];
  for (int i = 0; i < k->n_layers; i++) { L.woff[i] = (int)k->woff[i]; L.boff[i] = (int)k->boff[i]; }
  L.wcount = (int)k->wcount;
  return L;
}
static int km_grid(const So101Koopman* k, int64_t n) {
  const int64_t tiles = (n + KM_ROWS - 1) / KM_ROWS;
  return (int)(tiles < k->sms ? tiles : k->sms);  // persistent: one block per SM (the weights fill its shared memory)
}

extern "C" int so101_koopman_create(int n_layers, const int32_t* dims, const double* const* W, const double* const* b,
                                    int device, So101Koopman** out) {
  if (!dims || !W || !b || !out) return fail(SO101_EINVAL, "null argument");
  if (n_layers < 1 || n_layers > KM_MAXL) return fail(SO101_EINVAL, "koopman: 1..8 layers");
  if (dims[0] < 1 || dims[0] > KM_MAXX) return fail(SO101_EINVAL, "koopman: input width 1..16");
  for (int l = 1; l <= n_layers; l++)
    if (dims[l] < 8 || dims[l] > KM_MAXW || (dims[l] % 8)) return fail(SO101_EINVAL, "koopman: layer widths must be multiples of 8, 8..64");
  if (dims[0] + dims[n_layers] > KOOP_MAXZ) return fail(SO101_EINVAL, "koopman: lifted dimension exceeds 64");
  if (so101_device_count() <= 0) return fail(SO101_ENODEVICE, "no CUDA device visible: this library has no CPU fallback");
  DeviceGuard g(device);
  if (!g.ok) return fail(SO101_ECUDA, "cudaSetDevice failed");
  So101Koopman* k = new So101Koopman();
  std::memset(k, 0, sizeof *k);
  k->device = device; k->n_layers = n_layers;
  k->sms = 148;
  cudaDeviceGetAttribute(&k->sms, cudaDevAttrMultiProcessorCount, device);
  for (int l = 0; l <= n_layers; l++) k->dims[l] = dims[l];
  k->x_dim = dims[0]; k->nz = dims[0] + dims[n_layers];
  size_t off = 0;
  for (int l = 0; l < n_layers; l++) {               // Wt[l]: [din rounded up to 4][dout + KM_PAD], zero padded
    k->woff[l] = off; off += (size_t)((dims[l] + 3) & ~3) * (dims[l + 1] + KM_PAD);
    k->boff[l] = off; off += dims[l + 1];
    off += off & 1;
  }
  k->wcount = off;
  if (km_smem_bytes(k, false) + 4096 > 227 * 1024) { delete k; return fail(SO101_EINVAL, "koopman: the encoder does not fit shared memory"); }
  std::vector<double> host(off, 0.0);
  for (int l = 0; l < n_layers; l++) {
    if (!W[l] || !b[l]) { delete k; return fail(SO101_EINVAL, "null layer"); }
    const int din = dims[l], dout = dims[l + 1];
    for (int j = 0; j < dout; j++) {
      for (int c = 0; c < din; c++) host[k->woff[l] + (size_t)c * (dout + KM_PAD) + j] = W[l][(size_t)j * din + c];   // transposed
      host[k->boff[l] + j] = b[l][j];
    }
  }
  cudaError_t e = cudaMalloc(&k->weights, off * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(k->weights, host.data(), off * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cudaFree(k->weights); delete k; return fail(SO101_ECUDA, std::string("koopman_create: ") + cudaGetErrorString(e)); }
  *out = k;
  return SO101_OK;
}
extern "C" void so101_koopman_destroy(So101Koopman* k) {
  if (!k) return;
  DeviceGuard g(k->device);
  cudaFree(k->weights);
  cudaFree(k->gains);
  delete k;
}
extern "C" int so101_koopman_set_gains(So101Koopman* k, int H, int nu, const double* Kz, const double* Kr, const double* Ku) {
  if (!k || !Kz || !Kr || !Ku) return fail(SO101_EINVAL, "null argument");
  if (H < 1 || nu < 1 || nu > KOOP_MAXU) return fail(SO101_EINVAL, "koopman gains: H >= 1, 1 <= nu <= 8");
  if ((size_t)nu * H * k->nz * sizeof(double) > 96 * 1024) return fail(SO101_EINVAL, "koopman gains: reference gain exceeds 96 KB");
  DeviceGuard g(k->device);
  const size_t nzv = (size_t)nu * k->nz, nuv = (size_t)nu * nu, nrv = (size_t)nu * H * k->nz;
  std::vector<double> host(nzv + nuv + nrv);
  std::memcpy(&host[0], Kz, nzv * sizeof(double));
  std::memcpy(&host[nzv], Ku, nuv * sizeof(double));
  std::memcpy(&host[nzv + nuv], Kr, nrv * sizeof(double));
  cudaFree(k->gains);
  k->gains = nullptr;
  CUDA_TRY(cudaMalloc(&k->gains, host.size() * sizeof(double)));
  CUDA_TRY(cudaMemcpy(k->gains, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice));
  k->H = H; k->nu = nu;
  return SO101_OK;
}

template <typename TX>
static int km_launch_lift(So101Koopman* k, const TX* X, int layout, int64_t ldx, int64_t n, double* Z, cudaStream_t st) {
  const size_t smem = km_smem_bytes(k, false);
  bool& done = k->attr_lift[sizeof(TX) == 4];
  if (!done) { CUDA_TRY(cudaFuncSetAttribute(k_koopman_lift<TX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); done = true; }
  k_koopman_lift<TX><<<km_grid(k, n), KM_THREADS, smem, st>>>(km_layers(k), k->weights, X, layout, ldx, n, Z);
  CUDA_TRY(cudaGetLastError());
  return SO101_OK;
}
extern "C" int so101_koopman_lift(So101Koopman* k, const void* X, int dtype, int layout, int64_t ldx, int64_t n, double* Z,
                                  void* stream) {
  if (!k || !X || !Z || n <= 0) return fail(SO101_EINVAL, "null argument");
  if (layout != 0 && layout != 1) return fail(SO101_EINVAL, "layout: 0 = rows [n][ldx], 1 = structure of arrays [x_dim][n]");
  if (layout == 0 && ldx < k->x_dim) return fail(SO101_EINVAL, "ldx < x_dim");
  DeviceGuard g(k->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == SO101_F64) return km_launch_lift<double>(k, (const double*)X, layout, ldx, n, Z, st);
  if (dtype == SO101_F32) return km_launch_lift<float>(k, (const float*)X, layout, ldx, n, Z, st);
  return fail(SO101_EINVAL, "dtype must be SO101_F64 or SO101_F32");
}

extern "C" int so101_koopman_feedforward(So101Koopman* k, const double* Xref, int64_t n, int P, double* uff, void* stream) {
  if (!k || !Xref || !uff || n <= 0 || P <= 0) return fail(SO101_EINVAL, "null argument");
  if (!k->gains) return fail(SO101_EINVAL, "koopman_feedforward: set the gains first");
  DeviceGuard g(k->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the lifted reference rows of a chunk of envs go through a scratch buffer of at most 256 MB
  int64_t chunk = ((int64_t)256 << 20) / ((int64_t)P * k->nz * (int64_t)sizeof(double));
  if (chunk < 1) chunk = 1;
  if (chunk > n) chunk = n;
  double* Z = nullptr;
  CUDA_TRY(cudaMallocAsync(&Z, (size_t)chunk * P * k->nz * sizeof(double), st));
  if (k->nz % 4) return fail(SO101_EINVAL, "koopman_feedforward: the lifted dimension must be a multiple of 4");
  const size_t wsmem = (size_t)8 * k->H * k->nz * sizeof(double);
  cudaError_t err = cudaSuccess;
  if (wsmem > 48 * 1024) err = cudaFuncSetAttribute(k_koopman_window, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem);
  int rc = SO101_OK;
  for (int64_t e0 = 0; e0 < n && rc == SO101_OK && err == cudaSuccess; e0 += chunk) {
    const int64_t ne = n - e0 < chunk ? n - e0 : chunk;
    rc = km_launch_lift<double>(k, Xref + (size_t)e0 * P * k->x_dim, 0, k->x_dim, ne * P, Z, st);
    if (rc != SO101_OK) break;
    const int64_t tasks = ne * ((P + 7) / 8);       // one warp per 8 frames of an env, grid-stride
    const int64_t wblocks = (tasks + 3) / 4;
    k_koopman_window<<<(unsigned)(wblocks < (int64_t)k->sms * 8 ? wblocks : (int64_t)k->sms * 8), 128, wsmem, st>>>(
        Z, k->gains, k->nz, k->nu, k->H, ne, P, uff + (size_t)e0 * P * k->nu);
    err = cudaGetLastError();
  }
  cudaFreeAsync(Z, st);
  if (rc != SO101_OK) return rc;
  if (err != cudaSuccess) return fail(SO101_ECUDA, std::string("koopman_feedforward: ") + cudaGetErrorString(err));
  return SO101_OK;
}

template <typename TX, typename TC>
static int km_launch_mpc(So101Koopman* k, const TX* X, int layout, int64_t ldx, int64_t n, const double* uff, int64_t uff_stride,
                         double* u_prev, TC* ctrl, double* a_out, double clip, cudaStream_t st) {
  const size_t smem = km_smem_bytes(k, true);
  bool& done = k->attr_mpc[(sizeof(TX) == 4) * 2 + (sizeof(TC) == 4)];
  if (!done) { CUDA_TRY(cudaFuncSetAttribute(k_koopman_mpc<TX, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); done = true; }
  k_koopman_mpc<TX, TC><<<km_grid(k, n), KM_THREADS, smem, st>>>(km_layers(k), k->weights, k->gains, k->nu, X, layout, ldx, n,
                                                               uff, uff_stride, u_prev, ctrl, a_out, clip);
  CUDA_TRY(cudaGetLastError());
  return SO101_OK;
}
extern "C" int so101_koopman_mpc_step(So101Koopman* k, const void* obs, int obs_dtype, int obs_layout, int64_t ldx,
                                      const double* uff, int64_t uff_stride, double* u_prev, void* ctrl, int ctrl_dtype,
                                      double* a_out, double clip, int64_t n, void* stream) {
  if (!k || !obs || !u_prev || !ctrl || n <= 0) return fail(SO101_EINVAL, "null argument");
  if (!k->gains) return fail(SO101_EINVAL, "koopman_mpc_step: set the gains first");
  if (obs_layout != 0 && obs_layout != 1) return fail(SO101_EINVAL, "layout: 0 = rows [n][ldx], 1 = structure of arrays [x_dim][n]");
  if (obs_layout == 0 && ldx < k->x_dim) return fail(SO101_EINVAL, "ldx < x_dim");
  if ((obs_dtype != SO101_F64 && obs_dtype != SO101_F32) || (ctrl_dtype != SO101_F64 && ctrl_dtype != SO101_F32))
    return fail(SO101_EINVAL, "dtype must be SO101_F64 or SO101_F32");
  DeviceGuard g(k->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (obs_dtype == SO101_F64 && ctrl_dtype == SO101_F64)
    return km_launch_mpc<double, double>(k, (const double*)obs, obs_layout, ldx, n, uff, uff_stride, u_prev, (double*)ctrl, a_out, clip, st);
  if (obs_dtype == SO101_F32 && ctrl_dtype == SO101_F64)
    return km_launch_mpc<float, double>(k, (const float*)obs, obs_layout, ldx, n, uff, uff_stride, u_prev, (double*)ctrl, a_out, clip, st);
  if (obs_dtype == SO101_F64 && ctrl_dtype == SO101_F32)
    return km_launch_mpc<double, float>(k, (const double*)obs, obs_layout, ldx, n, uff, uff_stride, u_prev, (float*)ctrl, a_out, clip, st);
  return km_launch_mpc<float, float>(k, (const float*)obs, obs_layout, ldx, n, uff, uff_stride, u_prev, (float*)ctrl, a_out, clip, st);
}
