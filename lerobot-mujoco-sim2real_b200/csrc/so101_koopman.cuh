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
