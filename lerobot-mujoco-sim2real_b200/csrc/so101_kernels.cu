// so101_kernels.cu - one translation unit per (dtype, kernel family): compiled four times by build.py with
//   -DSO101_TU_T=double|float -DSO101_TU_SPLIT=false|true
// so that the stepping kernels build in parallel.  Everything is in so101_kernels.cuh.
#include "so101_kernels.cuh"

SO101_LAUNCHERS(, SO101_TU_T, SO101_TU_SPLIT)
