// Fused residual unit of the HiFi-GAN MRF blocks on tcgen05 (reference layers/residual_block.py:207-222):
//
//     at = lrelu(conv1_dilated(ax) + b1);   xn = conv2(at) + b2 + x;   axn = lrelu(xn)          (ax = lrelu(x))
//
// in ONE kernel for the narrow stages (C = 32 / 64), where a conv is a tiny GEMM (K = C * k) and each launch costs more
// in CTA start-up, pipeline fill and drain than in math: the intermediate `at` goes TMEM -> registers -> shared memory
// (written by the epilogue warps in the canonical K-major swizzled layout) and is the A operand of the second GEMM
// without ever round-tripping HBM (it is still written out once when the backward needs it).
//
// Tile = 128 INTERMEDIATE rows of one sequence: rows [q0 - p2, q0 - p2 + 128) of `at`, which give R = 128 - (k - 1) output
// rows [q0, q0 + R) (p2 = (k-1)/2; the 2-8 % overlap between neighbouring tiles is recomputed).  Both weights stay
// resident in shared memory for the CTA's lifetime; the input tile (128 + 2 p1 rows, p1 = p2 * dilation) is staged once
// by TMA (zero fill outside the sequence = conv1's padding) and every tap reads it through a row-shifted descriptor.
// Intermediate rows outside [0, L) are forced to zero (conv2's padding).
// mode 1 is the matching DATA GRADIENT of the unit, the same two-GEMM structure with the convolutions transposed
// (taps reversed, conv2's first):  dt = conv2^T(gx) * lrelu'(at);  gn = conv1^T(dt) * lrelu'(ax) + gx  — `dt` is
// written out for conv1's weight gradient.  Here the SECOND GEMM is the dilated one, so R = 128 - (k - 1) * dilation.
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..9 = epilogue (stage 1: TMEM -> at; stage 2: TMEM -> xn / axn; C = 64:
// every warp runs both stages for its 32-channel chunk; C = 32: warps 2..5 run stage 1, warps 6..9 stage 2, concurrently).
// The MMA warp runs GEMM1 of tile i+1 before GEMM2 of tile i, so the tensor pipe has work while the epilogue warps
// convert tile i's intermediate; accumulators and the intermediate buffer are double-buffered.
#include "tc_common.cuh"

namespace artic {
namespace tc {

constexpr int RU_THREADS = 320;
constexpr int RU_MAX_AS = 4;

struct RUPlan {
  int32_t C, k, dil1, dil2, p1, p2, R;   // dil1 / p1: first GEMM, dil2 / p2: second GEMM
  int32_t mode, rev;                     // mode 1 = data gradient (mask epilogues); rev = taps of both weights reversed
  int32_t row_bytes, layout_type;
  int32_t tiles_per_seq, total_tiles;
  int32_t split;                         // epilogue stages split between the warp quartets (C = 32 always; C = 64: debug key 30)
  int32_t nbox, a1_stage_bytes, n_as, n_a2;   // n_a2: intermediate buffers (2; 1 when the weights leave no room)
  int32_t w_tile_bytes, a2_bytes, tmem_cols;
  int32_t N, L;
  int64_t s_outer;
  float slope;
};

struct RUArgs {
  const __nv_bfloat16* m1;      // mode 1: activation whose sign masks the intermediate (at)
  const __nv_bfloat16* m2;      // mode 1: activation whose sign masks the output (ax)
  const __nv_bfloat16* xres;
  const float* b1;
  const float* b2;
  __nv_bfloat16* at;
  __nv_bfloat16* y;
  __nv_bfloat16* y2;
};

__global__ void __launch_bounds__(RU_THREADS, 1)
resunit_tc_kernel(const __grid_constant__ RUPlan pl, const __grid_constant__ RUArgs ar, const __grid_constant__ CUtensorMap map_x,
                  const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_w2) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[RU_MAX_AS], a_empty[RU_MAX_AS];
  __shared__ __align__(8) uint64_t w_full, acc1_full[2], acc1_empty[2], a2_full[2], a2_empty[2], acc2_full[2], acc2_empty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[2][64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = pl.C, k = pl.k;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a1_base = smem0;
  const uint32_t w_base = a1_base + (uint32_t)pl.n_as * pl.a1_stage_bytes;           // [W1 taps | W2 taps]
  const uint32_t a2_base = w_base + (uint32_t)(2 * k) * pl.w_tile_bytes;              // two intermediate buffers
  const int n_epi = (int)blockDim.x - 64;

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_w1);
    prefetch_tmap(&map_w2);
    for (int i = 0; i < pl.n_as; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    mbar_init(&w_full, 1);
    // C = 32: the two epilogue stages are SPLIT between the warp quartets (one channel chunk: half of the warps would
    // idle otherwise) and run concurrently; C = 64: every warp does both stages for its 32-channel chunk
    const uint32_t n_arr = pl.split ? (uint32_t)n_epi / 2 : (uint32_t)n_epi;
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], n_arr);
      mbar_init(&a2_full[i], n_arr); mbar_init(&a2_empty[i], 1);
      mbar_init(&acc2_full[i], 1); mbar_init(&acc2_empty[i], n_arr);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 2 * 64) {
    const int w = threadIdx.x >> 6, c = threadIdx.x & 63;
    const float* b = w ? ar.b2 : ar.b1;
    bias_s[w][c] = (b != nullptr && c < C) ? __ldg(b + c) : 0.f;
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)pl.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int cta = (int)blockIdx.x, ncta = (int)gridDim.x;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      pdl_wait();
      mbar_expect_tx(&w_full, (uint32_t)(2 * k) * (uint32_t)(C * pl.row_bytes));
      for (int t = 0; t < k; ++t) tma_load_2d(w_base + (uint32_t)t * pl.w_tile_bytes, &map_w1, &w_full, 0, (pl.rev ? k - 1 - t : t) * C);
      for (int t = 0; t < k; ++t) tma_load_2d(w_base + (uint32_t)(k + t) * pl.w_tile_bytes, &map_w2, &w_full, 0, (pl.rev ? k - 1 - t : t) * C);
      PipeState as(pl.n_as);
      for (int tile = cta; tile < pl.total_tiles; tile += ncta) {
        const int n = tile / pl.tiles_per_seq;
        const int q0 = (tile % pl.tiles_per_seq) * pl.R;
        const int row0 = q0 - pl.p2 - pl.p1;
        mbar_wait(&a_empty[as.stage], as.phase ^ 1);
        mbar_expect_tx(&a_full[as.stage], (uint32_t)pl.nbox * 64u * (uint32_t)pl.row_bytes);
        const uint32_t dst = a1_base + (uint32_t)as.stage * pl.a1_stage_bytes;
        for (int b = 0; b < pl.nbox; ++b)
          tma_load_4d(dst + (uint32_t)b * 64u * pl.row_bytes, &map_x, &a_full[as.stage], 0, 0, row0 + b * 64, n);
        as.next();
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t desc_hi = (((8u * (uint32_t)pl.row_bytes) >> 4) & 0x3fffu) | (1u << 14) | ((uint32_t)(pl.layout_type & 7) << 29);
    const uint32_t desc_lo = 1u << 16;
    const int ksteps = C / 16;
    const uint32_t rb16 = (uint32_t)pl.row_bytes >> 4, wt16 = (uint32_t)pl.w_tile_bytes >> 4;
    const uint32_t w16 = w_base >> 4;
    PipeState as(pl.n_as);
    mbar_wait(&w_full, 0);
    tc_fence_after();
    auto gemm2 = [&](int i) {                       // conv2 of the CTA's i-th tile: A = the intermediate in shared memory
      const int s = i & 1, ph = (i >> 1) & 1;
      const int s2 = i % pl.n_a2, ph2 = (i / pl.n_a2) & 1;
      mbar_wait(&a2_full[s2], (uint32_t)ph2);
      mbar_wait(&acc2_empty[s], (uint32_t)(ph ^ 1));
      tc_fence_after();
      const uint32_t a16 = desc_lo | (((a2_base + (uint32_t)s2 * pl.a2_bytes) >> 4) & 0x3fffu);
      const uint32_t d = tmem_base + (uint32_t)(2 * C) + (uint32_t)s * C;
      uint32_t accum = 0;
      for (int t = 0; t < k; ++t) {
        const uint32_t at16 = a16 + (uint32_t)(t * pl.dil2) * rb16;
        const uint32_t b16 = desc_lo | ((w16 + (uint32_t)(k + t) * wt16) & 0x3fffu);
        for (int ks = 0; ks < ksteps; ++ks) {
          if (leader) umma_bf16(d, ((uint64_t)desc_hi << 32) | (at16 + 2 * ks), ((uint64_t)desc_hi << 32) | (b16 + 2 * ks), idesc, accum);
          accum = 1;
        }
      }
      if (leader) { umma_commit(&a2_empty[s2]); umma_commit(&acc2_full[s]); }
    };
    int i = 0;
    for (int tile = cta; tile < pl.total_tiles; tile += ncta, ++i) {
      const int s = i & 1, ph = (i >> 1) & 1;
      // ---- conv1 (dilated): A = staged input tile, tap t shifted by t * dil rows
      mbar_wait(&acc1_empty[s], (uint32_t)(ph ^ 1));
      mbar_wait(&a_full[as.stage], as.phase);
      tc_fence_after();
      const uint32_t a16 = desc_lo | (((a1_base + (uint32_t)as.stage * pl.a1_stage_bytes) >> 4) & 0x3fffu);
      const uint32_t d = tmem_base + (uint32_t)s * C;
      uint32_t accum = 0;
      for (int t = 0; t < k; ++t) {
        const uint32_t at16 = a16 + (uint32_t)(t * pl.dil1) * rb16;
        const uint32_t b16 = desc_lo | ((w16 + (uint32_t)t * wt16) & 0x3fffu);
        for (int ks = 0; ks < ksteps; ++ks) {
          if (leader) umma_bf16(d, ((uint64_t)desc_hi << 32) | (at16 + 2 * ks), ((uint64_t)desc_hi << 32) | (b16 + 2 * ks), idesc, accum);
          accum = 1;
        }
      }
      if (leader) { umma_commit(&a_empty[as.stage]); umma_commit(&acc1_full[s]); }
      as.next();
      if (i > 0) gemm2(i - 1);          // conv2 of the previous tile, whose intermediate the epilogue warps wrote meanwhile
    }
    if (i > 0) gemm2(i - 1);
  } else {
    // =============================== epilogue ===================================
    pdl_wait();
    const int ew = warp & 3;                   // TMEM lane quarter
    const int eh = (warp - 2) >> 2;            // C = 64: channel chunk; C = 32: which epilogue stage this quartet runs
    const int r = ew * 32 + lane;              // tile row = TMEM lane
    const bool split = pl.split != 0;
    const bool has_ch = true;
    const uint32_t swz = pl.row_bytes == 128 ? (uint32_t)(r & 7) : (uint32_t)((r >> 1) & 3);
    const uint32_t t_lane = tmem_base + ((uint32_t)(ew * 32) << 16);
    uint8_t* a2_ptr = smem_raw + (a2_base - smem_u32(smem_raw));
    using TO = __nv_bfloat16;

    auto stage1 = [&](int i, int tile, int c0, bool first, bool last) {   // TMEM acc1 -> at (shared memory + optional HBM copy)
      const int s = i & 1, ph = (i >> 1) & 1;
      const int s2 = i % pl.n_a2, ph2 = (i / pl.n_a2) & 1;
      const int n = tile / pl.tiles_per_seq;
      const int q0 = (tile % pl.tiles_per_seq) * pl.R;
      const int gpos = q0 - pl.p2 + r;
      const bool inside = gpos >= 0 && gpos < pl.L;
      if (first) {
        mbar_wait(&acc1_full[s], (uint32_t)ph);
        mbar_wait(&a2_empty[s2], (uint32_t)(ph2 ^ 1));   // conv2 of tile i - n_a2 has finished reading this buffer
        tc_fence_after();
      }
      if (has_ch) {
        uint32_t acc_r[32];
        tmem_ld32(t_lane + (uint32_t)s * C + c0, acc_r);
        tmem_ld_wait();
        uint8_t* row = a2_ptr + (size_t)s2 * pl.a2_bytes + (size_t)r * pl.row_bytes;
        const bool keep = ar.at != nullptr && inside && r >= pl.p2 && r < pl.p2 + pl.R;
        TO* g = keep ? ar.at + (int64_t)n * pl.s_outer + (int64_t)gpos * C + c0 : nullptr;
        const TO* mk = (pl.mode == 1 && inside) ? ar.m1 + (int64_t)n * pl.s_outer + (int64_t)gpos * C + c0 : nullptr;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float v[8];
          if (pl.mode == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float x = __uint_as_float(acc_r[8 * u + j]) + bias_s[0][c0 + 8 * u + j];
              x = x > 0.f ? x : pl.slope * x;
              v[j] = inside ? x : 0.f;                   // conv2 sees zero padding outside the sequence
            }
          } else {
            float m[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) m[j] = 1.f;
            if (mk != nullptr) unpack8<TO>(__ldg(reinterpret_cast<const uint4*>(mk + 8 * u)), m);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = inside ? __uint_as_float(acc_r[8 * u + j]) * (m[j] > 0.f ? 1.f : pl.slope) : 0.f;
          }
          const uint4 pk = pack8(v);
          *reinterpret_cast<uint4*>(row + ((((uint32_t)(c0 >> 3) + u) ^ swz) << 4)) = pk;
          if (g != nullptr) *reinterpret_cast<uint4*>(g + 8 * u) = pk;
        }
      }
      if (last) {
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
        mbar_arrive(&acc1_empty[s]);
        mbar_arrive(&a2_full[s2]);
      }
    };
    auto stage2 = [&](int i, int tile, int c0, bool first, bool last) {   // TMEM acc2 -> xn = conv2 + b2 + x, axn = lrelu(xn)
      const int s = i & 1, ph = (i >> 1) & 1;
      const int n = tile / pl.tiles_per_seq;
      const int q0 = (tile % pl.tiles_per_seq) * pl.R;
      const int gpos = q0 + r;
      const bool valid = r < pl.R && gpos < pl.L && has_ch;
      const int64_t o = (int64_t)n * pl.s_outer + (int64_t)gpos * C + c0;
      uint4 q_rs[4], q_mk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        q_rs[u] = valid ? __ldg(reinterpret_cast<const uint4*>(ar.xres + o + 8 * u)) : make_uint4(0, 0, 0, 0);
        q_mk[u] = (valid && pl.mode == 1) ? __ldg(reinterpret_cast<const uint4*>(ar.m2 + o + 8 * u)) : make_uint4(0, 0, 0, 0);
      }
      if (first) {
        mbar_wait(&acc2_full[s], (uint32_t)ph);
        tc_fence_after();
      }
      if (has_ch) {
        uint32_t acc_r[32];
        tmem_ld32(t_lane + (uint32_t)(2 * C) + (uint32_t)s * C + c0, acc_r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float v[8], res[8];
            unpack8<TO>(q_rs[u], res);
            if (pl.mode == 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc_r[8 * u + j]) + bias_s[1][c0 + 8 * u + j] + res[j];
            } else {
              float m[8];
              unpack8<TO>(q_mk[u], m);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc_r[8 * u + j]) * (m[j] > 0.f ? 1.f : pl.slope) + res[j];
            }
            if (ar.y != nullptr) *reinterpret_cast<uint4*>(ar.y + o + 8 * u) = pack8(v);
            if (ar.y2 != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = v[j] > 0.f ? v[j] : pl.slope * v[j];
              *reinterpret_cast<uint4*>(ar.y2 + o + 8 * u) = pack8(v);
            }
          }
        }
      }
      if (last) {
        tc_fence_before();
        mbar_arrive(&acc2_empty[s]);
      }
    };
    if (split) {
      int i = 0;
      if (eh == 0) {
        for (int tile = cta; tile < pl.total_tiles; tile += ncta, ++i)
          for (int c0 = 0; c0 < C; c0 += 32) stage1(i, tile, c0, c0 == 0, c0 + 32 >= C);
      } else {
        for (int tile = cta; tile < pl.total_tiles; tile += ncta, ++i)
          for (int c0 = 0; c0 < C; c0 += 32) stage2(i, tile, c0, c0 == 0, c0 + 32 >= C);
      }
    } else {
      int i = 0, prev_tile = -1;
      const int c0 = eh * 32;
      for (int tile = cta; tile < pl.total_tiles; tile += ncta, ++i) {
        stage1(i, tile, c0, true, true);
        if (i > 0) stage2(i - 1, prev_tile, c0, true, true);
        prev_tile = tile;
      }
      if (i > 0) stage2(i - 1, prev_tile, c0, true, true);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)pl.tmem_cols);
  }
}

static int ru_smem_limit() {
  static int dyn = 0;
  if (dyn == 0) {
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (optin <= 0) optin = 227 * 1024;
    cudaFuncAttributes fa;
    int stat = 2048;
    if (cudaFuncGetAttributes(&fa, resunit_tc_kernel) == cudaSuccess) stat = (int)fa.sharedSizeBytes;
    dyn = optin - stat;
    if (cudaFuncSetAttribute(resunit_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess) {
      cudaGetLastError();
      dyn = 48 * 1024;
    }
  }
  return dyn;
}

}  // namespace tc
}  // namespace artic

using namespace artic;

/* see include/artic.h */
extern "C" int artic_resunit_fwd(const artic_resunit_t* pp, void* stream) {
  ARTIC_CHECK_ARG(pp != nullptr, "null params");
  const artic_resunit_t& p = *pp;
  ARTIC_CHECK_ARG(p.AX && p.XRES && p.W1t && p.W2t && (p.Y || p.Y2), "AX, XRES, W1t, W2t and one of Y / Y2 are required");
  ARTIC_CHECK_ARG(p.mode == 0 || (p.mode == 1 && p.M1 && p.M2), "mode 1 (data gradient) needs the two mask tensors");
  ARTIC_CHECK_ARG(p.N >= 0 && p.L >= 1, "bad dims");
  if (p.C != 32 && p.C != 64) { set_error("artic_resunit_fwd: C must be 32 or 64"); return ARTIC_ENOSUP; }
  if (p.k < 1 || p.k > 11 || !(p.k & 1) || p.dil < 1 || (p.k / 2) * p.dil > 32) { set_error("artic_resunit_fwd: unsupported kernel size / dilation"); return ARTIC_ENOSUP; }
  const void* ptrs[] = {p.AX, p.XRES, p.W1t, p.W2t, p.AT, p.Y, p.Y2, p.M1, p.M2};
  for (const void* q : ptrs)
    if (q != nullptr && (reinterpret_cast<uintptr_t>(q) & 15)) { set_error("artic_resunit_fwd: 16-byte alignment required"); return ARTIC_ENOSUP; }
  if (p.N == 0) return ARTIC_OK;
  tc::EncodeTiledFn enc = tc::encode_fn();
  if (enc == nullptr) { set_error("artic_resunit_fwd: cuTensorMapEncodeTiled unavailable"); return ARTIC_ECUDA; }

  tc::RUPlan pl;
  memset(&pl, 0, sizeof(pl));
  pl.C = p.C; pl.k = p.k;
  pl.mode = p.mode;
  pl.split = (p.C == 32 || tc::g_debug[30] == 1) ? 1 : 0;
  pl.rev = p.mode == 1 ? 1 : 0;                            // transposed convolutions: taps in reverse order
  pl.dil1 = p.mode == 0 ? p.dil : 1;                       // forward: conv1 (dilated) first; data gradient: conv2^T first
  pl.dil2 = p.mode == 0 ? 1 : p.dil;
  pl.p1 = (p.k / 2) * pl.dil1;
  pl.p2 = (p.k / 2) * pl.dil2;
  pl.R = 128 - 2 * pl.p2;
  pl.row_bytes = p.C * 2;
  pl.layout_type = pl.row_bytes == 128 ? 2 : 4;
  pl.tiles_per_seq = (p.L + pl.R - 1) / pl.R;
  const int64_t total = (int64_t)p.N * pl.tiles_per_seq;
  if (total > (1 << 30)) { set_error("artic_resunit_fwd: too many tiles"); return ARTIC_ENOSUP; }
  pl.total_tiles = (int)total;
  pl.nbox = (128 + 2 * pl.p1 + 63) / 64;
  pl.a1_stage_bytes = pl.nbox * 64 * pl.row_bytes;
  pl.w_tile_bytes = ((p.C * pl.row_bytes + 1023) / 1024) * 1024;
  pl.a2_bytes = (((128 + 2 * pl.p2 + 8) * pl.row_bytes + 1023) / 1024) * 1024;   // rows >= 128 only feed discarded outputs
  pl.tmem_cols = 4 * p.C;                                  // two accumulators x two stages (128 or 256 columns)
  pl.N = p.N; pl.L = p.L; pl.s_outer = (int64_t)p.L * p.C;
  pl.slope = p.slope;
  // C = 64 with k = 11: the two resident weights take 176 KB; one input stage and one intermediate buffer still fit (the
  // accumulators stay double-buffered, so conv1 of tile i + 1 still overlaps the conversion of tile i)
  pl.n_a2 = 2;
  int fixed = 2 * p.k * pl.w_tile_bytes + pl.n_a2 * pl.a2_bytes + 1024;
  if (tc::ru_smem_limit() - fixed < 2 * pl.a1_stage_bytes) {
    pl.n_a2 = 1;
    fixed = 2 * p.k * pl.w_tile_bytes + pl.n_a2 * pl.a2_bytes + 1024;
  }
  const int budget = tc::ru_smem_limit() - fixed;
  pl.n_as = budget / pl.a1_stage_bytes;
  if (pl.n_as > tc::RU_MAX_AS) pl.n_as = tc::RU_MAX_AS;
  if (pl.n_as < 1) { set_error("artic_resunit_fwd: weights do not fit in shared memory (C %d, k %d)", p.C, p.k); return ARTIC_ENOSUP; }

  CUtensorMap map_x, map_w1, map_w2;
  {
    cuuint64_t dims[4] = {(cuuint64_t)p.C, 1, (cuuint64_t)p.L, (cuuint64_t)p.N};
    cuuint64_t strides[3] = {(cuuint64_t)p.C * 2, (cuuint64_t)p.C * 2, (cuuint64_t)p.L * p.C * 2};
    cuuint32_t box[4] = {(cuuint32_t)p.C, 1, 64, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult rc = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.AX), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, tc::swizzle_of(pl.row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("artic_resunit_fwd: cuTensorMapEncodeTiled(X) failed (%d)", (int)rc); return ARTIC_ECUDA; }
  }
  for (int w = 0; w < 2; ++w) {
    cuuint64_t dims[2] = {(cuuint64_t)p.C, (cuuint64_t)p.k * p.C};
    cuuint64_t strides[1] = {(cuuint64_t)p.C * 2};
    cuuint32_t box[2] = {(cuuint32_t)p.C, (cuuint32_t)p.C};
    cuuint32_t es[2] = {1, 1};
    CUresult rc = enc(w ? &map_w2 : &map_w1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w ? p.W2t : p.W1t), dims, strides,
                      box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, tc::swizzle_of(pl.row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("artic_resunit_fwd: cuTensorMapEncodeTiled(W) failed (%d)", (int)rc); return ARTIC_ECUDA; }
  }
  tc::RUArgs ar;
  ar.m1 = reinterpret_cast<const __nv_bfloat16*>(p.M1);
  ar.m2 = reinterpret_cast<const __nv_bfloat16*>(p.M2);
  ar.xres = reinterpret_cast<const __nv_bfloat16*>(p.XRES);
  ar.b1 = p.b1;
  ar.b2 = p.b2;
  ar.at = reinterpret_cast<__nv_bfloat16*>(p.AT);
  ar.y = reinterpret_cast<__nv_bfloat16*>(p.Y);
  ar.y2 = reinterpret_cast<__nv_bfloat16*>(p.Y2);
  const int smem_bytes = pl.n_as * pl.a1_stage_bytes + fixed;
  // as few CTAs as finish in the same number of tile rounds: 525 tiles take 4 rounds on 148 or on 132 SMs, and the 16
  // SMs left free go to the launches of the other MRF branches that run beside this one
  const int rounds = (pl.total_tiles + num_sms() - 1) / num_sms();
  const int grid = (pl.total_tiles + rounds - 1) / rounds;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(tc::RU_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tc::g_debug[11] != 1 ? 1 : 0;
  cudaError_t le = cudaLaunchKernelEx(&cfg, tc::resunit_tc_kernel, pl, ar, map_x, map_w1, map_w2);
  if (le == cudaSuccess) le = cudaGetLastError();
  if (le != cudaSuccess) {
    set_error("artic_resunit_fwd: launch failed: %s (grid %d, smem %d, C %d k %d mode %d)", cudaGetErrorString(le), grid, smem_bytes, p.C, p.k, p.mode);
    return ARTIC_ECUDA;
  }
  ++g_path_counts[10];     // fused residual units (two convs each)
  return ARTIC_OK;
}
