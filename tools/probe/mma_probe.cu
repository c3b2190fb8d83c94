// Micro-benchmark: rate of tcgen05.mma (M=128, N=bn, K=16, bf16, SS operands, K-major SW128) with
//   data   : 0 = constant operands, 1 = random bf16 operands
//   bulk   : 0 = no other smem traffic, 1 = a second warp streams 16 KB cp.async.bulk copies into smem
//   walk   : 0 = same operand addresses every MMA, 1 = operands walk over a 64 KB window (as a real main loop)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../articulatory_b200/csrc/tc_common.cuh"
using namespace artic::tc;

__global__ void __launch_bounds__(128, 1) probe(int bn, int n_mma, int data, int bulk, int walk, const uint8_t* gsrc, long long* out_all) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t done, cp_bar;
  __shared__ volatile int stop;
  const uint32_t s0 = (smem_u32(smem) + 1023u) & ~1023u;
  uint32_t seed = threadIdx.x * 2654435761u + 12345u;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) {
    seed = seed * 1664525u + 1013904223u;
    uint32_t v = 0x3c003c00u;
    if (data) v = (seed & 0x807f807fu) | 0x3f003f00u;      // random sign / mantissa, exponent ~1
    reinterpret_cast<uint32_t*>(smem)[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(&done, 1); mbar_init(&cp_bar, 1); stop = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (threadIdx.x < 32) tmem_alloc(&tmem_base_s, 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    long long* out = out_all + 2 * blockIdx.x;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t hi = ((1024u >> 4) & 0x3fffu) | (1u << 14) | (2u << 29);
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t k = (i & 3) * 2;
      const uint32_t tile = walk ? ((i >> 2) & 3) : 0;        // 4 A tiles (16 KB apart) and 4 B tiles (32 KB apart... capped)
      const uint32_t a0 = (1u << 16) | (((s0 + tile * 16384) >> 4) & 0x3fffu);
      const uint32_t b0 = (1u << 16) | (((s0 + 65536 + (tile & 1) * 32768) >> 4) & 0x3fffu);
      umma_bf16(tmem_base_s, ((uint64_t)hi << 32) | (a0 + k), ((uint64_t)hi << 32) | (b0 + k), idesc, i > 0);
    }
    const long long t1 = clock64();
    umma_commit(&done);
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    stop = 1;
  } else if (threadIdx.x == 32 && bulk) {
    // background: stream 16 KB bulk copies global -> smem (region above 128 KB), back to back
    uint32_t phase = 0;
    const uint32_t dst = s0 + 131072;
    while (!stop) {
      mbar_expect_tx(&cp_bar, 16384);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(dst), "l"(gsrc + (size_t)blockIdx.x * 16384), "r"(16384), "r"(smem_u32(&cp_bar)) : "memory");
      mbar_wait(&cp_bar, phase);
      phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base_s, 512); }
}

int main() {
  long long* d;
  uint8_t* g;
  cudaMalloc(&d, 16 * 148);
  cudaMalloc(&g, 16384 * 148);
  cudaMemset(g, 0x3c, 16384 * 148);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int n = 4096;
  for (int bn : {64, 128, 256})
    for (int ctas : {1, 148})
      for (int data : {0, 1})
        for (int bulk : {0, 1})
          for (int walk : {0, 1}) {
            long long h[2];
            for (int rep = 0; rep < 2; ++rep) {
              probe<<<ctas, 128, 200 * 1024>>>(bn, n, data, bulk, walk, g, d);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("bn %3d ctas %3d data %d bulk %d walk %d : issue %.1f clk/mma, complete %.1f clk/mma\n", bn, ctas, data, bulk,
                   walk, (double)h[0] / n, (double)h[1] / n);
          }
  return 0;
}
