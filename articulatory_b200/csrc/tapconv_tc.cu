// tcgen05 implicit-GEMM path for the dense bf16 contractions (placeholder until the kernel lands).
#include "common.cuh"

int artic_tapconv_tc_try(const artic_tapconv_t* p, cudaStream_t st) {
  (void)p; (void)st;
  return 0;  // not eligible -> generic kernel
}
