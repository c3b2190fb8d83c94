// Bidirectional GRU layer, inference forward (speech-to-EMA inversion encoder: reference
// models/pytorch_models.py:22-77, torch.nn.GRU(batch_first, bidirectional) at :27,30 / :63-66).
//
// The input projections W_ih x + b_ih of ALL time steps are one plain GEMM (done by the caller through the
// tap-gather contraction, k = 1); what remains is the strictly sequential part
//
//     gh = W_hh h + b_hh;  r = s(gi_r + gh_r);  z = s(gi_z + gh_z);  n = tanh(gi_n + r * gh_n);  h' = (1 - z) n + z h
//
// T dependent steps of a (items x H) x (H x 3H) product — latency bound, so the design is a PERSISTENT kernel that
// keeps everything on chip for the whole sequence:
//
//   * one thread-block CLUSTER per (direction, group of 8 batch items); the cluster's CTAs split the hidden
//     units (64 per CTA, cluster size ceil(H / 64) <= 4: 33 such clusters are co-resident on a B200, so batch 128 x
//     2 directions is ONE wave; clusters of 8 are limited to ~1 per GPC and needed two);
//   * each CTA holds ITS rows of W_hh (3 gates x 64 units x H, fp32) in shared memory for the whole sequence
//     (192 KB at H = 256) — the weights are read from HBM exactly once;
//   * h lives in shared memory, replicated in every CTA of the cluster ([k][item]: two broadcast 128-bit reads feed
//     all eight items); after each step a CTA writes its 64 new units straight into the OTHER CTAs' copies through
//     distributed shared memory and the cluster synchronises once (double-buffered h);
//   * fp32 FFMA throughout (lanes along the hidden units: conflict-free weight reads): the recurrence feeds its
//     own rounding error back T times, and the parity gate is 1e-3 against the fp32 reference;
//   * gi of step t+1 is fetched into registers before the FMA loop of step t (hides the L2 latency).
#include <cooperative_groups.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace artic {

constexpr int GRU_ITEMS = 8;       // batch items per cluster
constexpr int GRU_UNITS = 64;      // hidden units per CTA
constexpr int GRU_THREADS = 256;   // 8 warps = (unit half of 32 lanes) x (quarter of the k range)

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// Remote shared-memory store that signals the destination CTA's mbarrier when it lands (no fence, no cluster barrier:
// `cluster.sync` has release semantics over ALL earlier memory operations, so every step waited for its global output
// stores to be acknowledged — 18 % of the kernel's stall samples, profiles/r2_ncu_gru_*).
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f2(uint32_t remote_addr, float a, float b, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
               ::"r"(remote_addr), "f"(a), "f"(b), "r"(remote_bar) : "memory");
}

// Thread (unit u, k quarter kq) accumulates ALL 8 items x 3 gates over its quarter of k — every weight is read from
// shared memory once per step per CTA (3 conflict-free 128-byte wavefronts + 2 broadcast reads of h per 24 FMAs) —
// then the four partial sums meet through shared memory and the same thread finishes items 2kq, 2kq+1 of its unit.
__global__ void __launch_bounds__(GRU_THREADS, 1)
bigru_layer_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                   float* __restrict__ out, int N, int T, int H) {
  cg::cluster_group cluster = cg::this_cluster();
  const int csize = (int)cluster.num_blocks();          // = ceil(H / 64)
  const int rank = (int)cluster.block_rank();
  const int cid = (int)blockIdx.x / csize;               // cluster index
  const int n_groups = (N + GRU_ITEMS - 1) / GRU_ITEMS;
  const int dir = cid / n_groups;                        // 0 forward, 1 reverse
  const int n0 = (cid % n_groups) * GRU_ITEMS;

  __shared__ __align__(8) uint64_t hbar[2];             // "h buffer b is complete" (transaction barriers)
  extern __shared__ __align__(16) float smem[];
  float* Wt = smem;                                      // [H k][3 gates][64 units]
  float* hbuf = Wt + 3 * H * GRU_UNITS;                  // [2][H k][8 items]
  float* red = hbuf + 2 * H * GRU_ITEMS;                 // [4 dest kq][3 sources][6 = gate x item][64 units]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ul = (warp & 1) * 32 + lane;                 // unit inside the CTA
  const int kq = warp >> 1;                              // k quarter; this thread FINISHES items 2kq, 2kq + 1
  const int unit = rank * GRU_UNITS + ul;                // hidden unit
  const bool live = unit < H;

  // ---- one-time: this CTA's rows of W_hh (transposed to [gate][k][unit]), zero initial state
  const float* Wd = w_hh + (size_t)dir * 3 * H * H;
  for (int i = tid; i < 3 * GRU_UNITS * H; i += GRU_THREADS) {
    const int k = i % H, u = (i / H) % GRU_UNITS, g = i / (H * GRU_UNITS);
    Wt[(k * 3 + g) * GRU_UNITS + u] = (rank * GRU_UNITS + u < H) ? __ldg(Wd + ((size_t)g * H + rank * GRU_UNITS + u) * H + k) : 0.f;
  }
  for (int i = tid; i < 2 * H * GRU_ITEMS; i += GRU_THREADS) hbuf[i] = 0.f;
  if (tid == 0) {
    tc::mbar_init(&hbar[0], 1);
    tc::mbar_init(&hbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t h_bytes = (uint32_t)H * GRU_ITEMS * sizeof(float);     // one complete h: every live unit x 8 items
  float bh[3] = {0.f, 0.f, 0.f};
  if (live) {
#pragma unroll
    for (int g = 0; g < 3; ++g) bh[g] = __ldg(b_hh + (size_t)dir * 3 * H + g * H + unit);
  }
  cluster.sync();

  const int GS = 2 * 3 * H;                              // gi row stride (both directions)
  const float* gi_d = gi + (size_t)dir * 3 * H + unit;
  const int na = n0 + 2 * kq, nb = na + 1;               // the two batch items this thread finishes
  float gin[3][2];                                       // their gi of the NEXT step (gate, item)
  auto fetch = [&](int t) {
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      gin[g][0] = (na < N && live) ? __ldg(gi_d + ((size_t)na * T + t) * GS + g * H) : 0.f;
      gin[g][1] = (nb < N && live) ? __ldg(gi_d + ((size_t)nb * T + t) * GS + g * H) : 0.f;
    }
  };
  if (T > 0) fetch(dir ? T - 1 : 0);

  const int kb = kq * (H / 4), ke = kb + H / 4;          // H % 4 == 0 (checked by the caller)
  for (int s = 0; s < T; ++s) {
    const int t = dir ? T - 1 - s : s;
    const float* hc = hbuf + (size_t)(s & 1) * H * GRU_ITEMS;          // h_t (complete, all units)
    float* hn_local = hbuf + (size_t)((s + 1) & 1) * H * GRU_ITEMS;    // h_{t+1} (being assembled)
    // arm the barrier of the buffer that fills during this step, then wait for h_t (filled during step s - 1: the
    // ((s - 1) / 2)-th fill of buffer s & 1).  A peer can run at most one step ahead (it needs OUR h_{t+1} to go further),
    // so its writes into hn_local never overtake our reads of that buffer from step s - 1.
    if (tid == 0) tc::mbar_expect_tx(&hbar[(s + 1) & 1], h_bytes);
    if (s > 0) tc::mbar_wait(&hbar[s & 1], (uint32_t)(((s - 1) >> 1) & 1));
    float gcur[3][2];
#pragma unroll
    for (int g = 0; g < 3; ++g) { gcur[g][0] = gin[g][0]; gcur[g][1] = gin[g][1]; }
    if (s + 1 < T) fetch(dir ? t - 1 : t + 1);
    // packed fp32 FMAs (fma.rn.f32x2: two items per instruction) — the loop is FMA-issue bound
    float2 acc2[3][4];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc2[g][i] = make_float2(0.f, 0.f);
    const float* wp = Wt + (size_t)kb * 3 * GRU_UNITS + ul;
    const float* hp = hc + (size_t)kb * GRU_ITEMS;
#pragma unroll 4
    for (int k = kb; k < ke; ++k, wp += 3 * GRU_UNITS, hp += GRU_ITEMS) {
      const float4 ha = *reinterpret_cast<const float4*>(hp);          // warp-wide broadcasts
      const float4 hb = *reinterpret_cast<const float4*>(hp + 4);
      const float2 h2[4] = {make_float2(ha.x, ha.y), make_float2(ha.z, ha.w), make_float2(hb.x, hb.y), make_float2(hb.z, hb.w)};
      const float w0 = wp[0], w1 = wp[GRU_UNITS], w2 = wp[2 * GRU_UNITS];
      const float2 w0p = make_float2(w0, w0), w1p = make_float2(w1, w1), w2p = make_float2(w2, w2);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc2[0][i] = __ffma2_rn(w0p, h2[i], acc2[0][i]);
        acc2[1][i] = __ffma2_rn(w1p, h2[i], acc2[1][i]);
        acc2[2][i] = __ffma2_rn(w2p, h2[i], acc2[2][i]);
      }
    }
    float acc[3][8];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int i = 0; i < 4; ++i) { acc[g][2 * i] = acc2[g][i].x; acc[g][2 * i + 1] = acc2[g][i].y; }
    // hand the partial sums of the items finished by the other three k quarters to them
    float own[3][2];
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      if (d == kq) {
#pragma unroll
        for (int g = 0; g < 3; ++g) { own[g][0] = acc[g][2 * d]; own[g][1] = acc[g][2 * d + 1]; }
      } else {
        const int slot = kq < d ? kq : kq - 1;
        float* r = red + (size_t)((d * 3 + slot) * 6) * GRU_UNITS + ul;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          r[(g * 2 + 0) * GRU_UNITS] = acc[g][2 * d];
          r[(g * 2 + 1) * GRU_UNITS] = acc[g][2 * d + 1];
        }
      }
    }
    __syncthreads();
    if (live) {
      float gh[3][2];
#pragma unroll
      for (int g = 0; g < 3; ++g) { gh[g][0] = own[g][0] + bh[g]; gh[g][1] = own[g][1] + bh[g]; }
#pragma unroll
      for (int slot = 0; slot < 3; ++slot) {
        const float* r = red + (size_t)((kq * 3 + slot) * 6) * GRU_UNITS + ul;
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          gh[g][0] += r[(g * 2 + 0) * GRU_UNITS];
          gh[g][1] += r[(g * 2 + 1) * GRU_UNITS];
        }
      }
      const float2 hold = *reinterpret_cast<const float2*>(hc + unit * GRU_ITEMS + 2 * kq);
      const float ho[2] = {hold.x, hold.y};
      float hv2[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float rg = sigmoidf_(gcur[0][j] + gh[0][j]);
        const float zg = sigmoidf_(gcur[1][j] + gh[1][j]);
        const float ng = tanhf(gcur[2][j] + rg * gh[2][j]);
        hv2[j] = (1.f - zg) * ng + zg * ho[j];
        const int n = na + j;
        if (n < N) out[((size_t)n * T + t) * (2 * H) + dir * H + unit] = hv2[j];
      }
      // publish this unit's two items of h_{t+1} to every CTA of the cluster (distributed shared memory); each store
      // counts its 8 bytes on the destination CTA's barrier
      const uint32_t dst_local = tc::smem_u32(hn_local + unit * GRU_ITEMS + 2 * kq);
      const uint32_t bar_local = tc::smem_u32(&hbar[(s + 1) & 1]);
      for (int pr = 0; pr < csize; ++pr)
        st_async_f2(map_to_rank(dst_local, (uint32_t)pr), hv2[0], hv2[1], map_to_rank(bar_local, (uint32_t)pr));
    }
    __syncthreads();    // `red` and this CTA's reads of h_t are done before the next step overwrites them
  }
  cluster.sync();       // no CTA exits while a peer may still write into its shared memory
}

}  // namespace artic

using namespace artic;

/* see include/artic.h */
extern "C" int artic_bigru_layer(const float* gi, const float* w_hh, const float* b_hh, float* out, int32_t N, int32_t T,
                                 int32_t H, void* stream) {
  ARTIC_CHECK_ARG(gi && w_hh && b_hh && out, "null pointer");
  ARTIC_CHECK_ARG(N >= 0 && T >= 0, "bad dims");
  ARTIC_CHECK_ARG(H >= 4 && H <= 256 && H % 4 == 0, "hidden size must be a multiple of 4, <= 256 (64 units per CTA, cluster of <= 4)");
  if (N == 0 || T == 0) return ARTIC_OK;
  const int csize = (H + 63) / 64;
  const int n_groups = (N + GRU_ITEMS - 1) / GRU_ITEMS;
  const size_t smem = sizeof(float) * ((size_t)3 * H * GRU_UNITS + 2 * (size_t)H * GRU_ITEMS + 4 * 3 * 6 * GRU_UNITS);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(bigru_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      set_error("artic_bigru_layer: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
      return ARTIC_ECUDA;
    }
    smem_set = smem;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * n_groups * csize));
  cfg.blockDim = dim3(GRU_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, bigru_layer_kernel, gi, w_hh, b_hh, out, (int)N, (int)T, (int)H);
  if (e != cudaSuccess) {
    set_error("artic_bigru_layer: launch failed: %s", cudaGetErrorString(e));
    return ARTIC_ECUDA;
  }
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}
