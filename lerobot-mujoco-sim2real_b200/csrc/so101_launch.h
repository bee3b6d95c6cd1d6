// so101_launch.h - what the host side (so101_capi.cu) and the kernel translation units (so101_kernels.cu) share: the view
// of the state buffer, the control specification as the kernels take it, and the launchers of the stepping kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "so101_model.h"

// rows of the structure-of-arrays state buffer [NROWS][N] (+ uint32 flags[N]): see DESIGN.md section 3
constexpr int ROW_Q = 0, ROW_QD = 6, ROW_WARM = 12, ROW_FA = 18, ROW_TIME = 24, NROWS = 25;

template <typename T>
struct StateView {
  T* base;
  uint32_t* flags;
  int64_t n;
};
struct DevSpec {
  int32_t kind, t_total;
  uint64_t seed;
  int64_t env_offset;
  double amp, freq_lo, freq_hi, reset_lo, reset_hi;
  const void* u;
};

struct State0 { double v[18]; };

// One launcher per stepping kernel, instantiated for T in {double, float} and SPLIT in {false: one warp per 32 envs,
// true: a team of three warps per 32 envs} in four translation units that compile in parallel.
template <typename T, bool SPLIT>
cudaError_t launch_step(const so101::DevModel<T>& m, StateView<T> v, unsigned grid, int blk, cudaStream_t st, const T* ctrl,
                        int n_ctrl, int nsub, float* obs, unsigned long long* stats, uint32_t sflags);
template <typename T, bool SPLIT>
cudaError_t launch_rollout(const so101::DevModel<T>& m, StateView<T> v, unsigned grid, int blk, cudaStream_t st,
                           const DevSpec& ds, int t0, int t1, int Tn, int frame_skip, void* rows, bool rows_f32,
                           uint32_t rflags, unsigned long long* stats);
template <typename T, bool SPLIT>
cudaError_t launch_shoot(const so101::DevModel<T>& m, StateView<T> v, unsigned grid, int blk, cudaStream_t st,
                         const State0& s0, const T* U, int H, int frame_skip, float* X, uint32_t rflags,
                         unsigned long long* stats);
// the time-sliced rollout of the one-warp kernels (see k_rollout_sliced) and the number of its blocks an SM holds
template <typename T>
cudaError_t launch_rollout_sliced(const so101::DevModel<T>& m, StateView<T> v, unsigned grid, int blk, cudaStream_t st,
                                  const DevSpec& ds, int t0, int t1, int Tn, int frame_skip, void* rows, bool rows_f32,
                                  uint32_t rflags, unsigned long long* stats, int tchunk, int32_t* progress, int32_t* fault,
                                  const int32_t* perm, uint8_t* recent);
template <typename T> int rollout_sliced_blocks_per_sm(int blk);
#define SO101_SLICED_LAUNCHERS(KW, T)                                                                                      \
  KW template cudaError_t launch_rollout_sliced<T>(const so101::DevModel<T>&, StateView<T>, unsigned, int, cudaStream_t,    \
                                                   const DevSpec&, int, int, int, int, void*, bool, uint32_t,              \
                                                   unsigned long long*, int, int32_t*, int32_t*, const int32_t*, uint8_t*); \
  KW template int rollout_sliced_blocks_per_sm<T>(int);
#define SO101_LAUNCHERS(KW, T, SPLIT)                                                                                      \
  KW template cudaError_t launch_step<T, SPLIT>(const so101::DevModel<T>&, StateView<T>, unsigned, int, cudaStream_t,       \
                                                const T*, int, int, float*, unsigned long long*, uint32_t);                 \
  KW template cudaError_t launch_rollout<T, SPLIT>(const so101::DevModel<T>&, StateView<T>, unsigned, int, cudaStream_t,    \
                                                   const DevSpec&, int, int, int, int, void*, bool, uint32_t,              \
                                                   unsigned long long*);                                                    \
  KW template cudaError_t launch_shoot<T, SPLIT>(const so101::DevModel<T>&, StateView<T>, unsigned, int, cudaStream_t,      \
                                                 const State0&, const T*, int, int, float*, uint32_t, unsigned long long*);
