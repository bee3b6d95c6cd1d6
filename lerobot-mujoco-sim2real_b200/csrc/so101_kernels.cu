// so101_kernels.cu - one translation unit per (dtype, kernel family): compiled four times by build.py with
//   -DSO101_TU_T=double|float -DSO101_TU_SPLIT=false|true
// so that the stepping kernels build in parallel.  Everything is in so101_kernels.cuh.
#include "so101_kernels.cuh"
#define SO101_CAT2(a, b) a##b
#define SO101_CAT(a, b) SO101_CAT2(a, b)
#define SO101_TIMING_NAME SO101_CAT(so101_debug_timing_, SO101_TU_T)

SO101_LAUNCHERS(, SO101_TU_T, SO101_TU_SPLIT)
#if !SO101_TU_SPLIT
SO101_SLICED_LAUNCHERS(, SO101_TU_T)
#endif

#ifdef SO101_TIMING
#if SO101_TU_SPLIT
extern "C" int SO101_TIMING_NAME(unsigned long long* out, int reset) {
  cudaMemcpyFromSymbol(out, so101::g_timing, sizeof(unsigned long long) * 16);
  cudaMemcpyFromSymbol(out + 16, so101::g_timing_helpers, sizeof(unsigned long long) * 8);
  cudaMemcpyFromSymbol(out + 24, so101::g_timing_con, sizeof(unsigned long long) * 4);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(so101::g_timing, z, sizeof z);
    cudaMemcpyToSymbol(so101::g_timing_helpers, z, sizeof(unsigned long long) * 8);
    cudaMemcpyToSymbol(so101::g_timing_con, z, sizeof(unsigned long long) * 4);
  }
  return 0;
}
#endif
#endif
