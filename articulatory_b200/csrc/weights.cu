// Weight preparation (weight-norm + relayout), its backward, and the fused Adam step.
//
// Three generations of the relayout live here; the host (artic_weights_prep / artic_weights_unprep) picks per layer:
//   * row-run kernels (wprep_rows / wunprep_rows + warp-per-row wn_scale_rows / wn_bwd_rows): the default for every
//     conv / linear / transposed-conv weight (taps innermost in the torch layout) — see the block comment above RT_O;
//   * 32 x 32 x <= 8-tap tile kernels (wprep / wunprep + block-per-row wn_scale4 / wn_bwd4): any strides
//     (ARTIC_WEIGHTS_GENERIC=1 forces them; debug key 27 selects the block-per-row weight-norm kernels);
//   * the three-pass wperm<0,1,2> + wn_scale / wn_bwd originals (debug key 12), kept as the plainest statement.
#include "common.cuh"
#include "tc_common.cuh"

namespace artic {

// ---- batched ("multi-tensor") weight preparation ----------------------------------------
// One launch handles every layer of a network: blockIdx.y = layer, blockIdx.x strides over the
// layer's rows / tiles.  The descriptor table lives in device memory (uploaded once).

// scale[row] = g[row] / ||v[row]||, scale[rows + row] = ||v[row]||
__global__ void __launch_bounds__(256) wn_scale_kernel(const artic_wdesc_t* __restrict__ descs) {
  __shared__ float red[32];
  const artic_wdesc_t& d = descs[blockIdx.y];
  if (d.g == nullptr) return;
  for (int row = blockIdx.x; row < d.rows; row += gridDim.x) {
    const float* vr = d.v + (int64_t)row * d.row_len;
    float s = 0.f;
    for (int64_t e = threadIdx.x; e < d.row_len; e += blockDim.x) {
      const float x = vr[e];
      s = fmaf(x, x, s);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
      const float nrm = sqrtf(s);
      d.scale[row] = d.g[row] / nrm;
      d.scale[d.rows + row] = nrm;
    }
    __syncthreads();
  }
}

// Tiled permutation between the torch weight and a prepared layout, through shared memory so
// that BOTH sides move in contiguous runs (torch side: runs of inner-dim x taps; prepared side:
// runs of 32 columns).  MODE 0: torch -> 'fwd' layout [K][G/m][a_pad][b_pad] (rows = A, cols = B)
//                        MODE 1: torch -> 'bwd' layout [K][G/m][b_pad][a_pad] (rows = B, cols = A)
//                        MODE 2: fp32 dWp ('fwd' layout) -> dv (torch layout), overwriting
constexpr int PT = 32;   // tile edge
constexpr int PK = 8;    // taps per tile
// Tiles of ALL layers form one flat index space (descs[i].tile_begin = exclusive prefix of the per-layer
// tile counts, computed by the host with wperm_tiles()), so the grid is balanced over the big layers.
template <int MODE>
__global__ void __launch_bounds__(256) wperm_kernel(const artic_wdesc_t* __restrict__ descs, int n_layers,
                                                    long long total_tiles) {
  __shared__ float tile[PK][PT][PT + 1];
  for (long long gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
  int lo = 0, hi = n_layers - 1;                       // last layer with tile_begin <= gt
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (descs[mid].tile_begin <= gt) lo = mid; else hi = mid - 1;
  }
  const artic_wdesc_t& d = descs[lo];
  void* outp = MODE == 0 ? d.out_f : MODE == 1 ? d.out_b : (void*)d.dv;
  if (outp == nullptr || (MODE == 2 && d.dWp == nullptr)) continue;
  const bool swap = MODE == 1 || (MODE == 2 && d.dw_swapped != 0);
  const int Rn = swap ? d.B : d.A, Cn = swap ? d.A : d.B;          // rows / cols of the prepared matrix
  const int64_t sr = swap ? d.sb : d.sa, sc = swap ? d.sa : d.sb;  // their strides in the torch weight
  const int r_pad = swap ? d.b_pad : d.a_pad, c_pad = swap ? d.a_pad : d.b_pad;
  const int m = d.merge, Gs = d.G / m;
  const int kc = d.K < PK ? d.K : PK;
  const int n_rt = (Rn + PT - 1) / PT, n_ct = (Cn + PT - 1) / PT, n_kt = (d.K + kc - 1) / kc;
  const bool r_inner = sr < sc;                                      // which matrix dim is contiguous-ish in torch
  const int dtype = MODE == 0 ? d.dtype_f : d.dtype_b;
  {
    int64_t w = gt - d.tile_begin;
    const int ct = (int)(w % n_ct); w /= n_ct;
    const int rt = (int)(w % n_rt); w /= n_rt;
    const int kt = (int)(w % n_kt);
    const int g = (int)(w / n_kt);
    const int r0 = rt * PT, c0 = ct * PT, k0 = kt * kc;
    const int kn = min(kc, d.K - k0);
    const int64_t tbase = (int64_t)g * d.sg + (int64_t)k0 * d.sk;
    // prepared-side offset of element (kk, r, c)
    auto poff = [&](int kk, int r, int c) -> int64_t {
      return (((int64_t)(k0 + kk) * Gs + g / m) * r_pad + (g % m) * Rn + r0 + r) * (int64_t)c_pad + (g % m) * Cn + c0 + c;
    };
    // thread mapping without divisions: lane l = tid % 32 runs along the contiguous side of each phase
    // (torch side: the torch-inner matrix dim, each thread walking the kn contiguous taps; prepared
    // side: the column), w8 = tid / 32 strides the other matrix dim in steps of 8.
    const int l = threadIdx.x & 31, w8 = threadIdx.x >> 5;
    if (MODE != 2) {
      for (int o = w8; o < PT; o += 8) {
        const int r = r_inner ? l : o, c = r_inner ? o : l;
        const bool ok = r0 + r < Rn && c0 + c < Cn;
        const int64_t src = tbase + (int64_t)(r0 + r) * sr + (int64_t)(c0 + c) * sc;
        const float sc_row = (ok && d.g != nullptr) ? d.scale[src / d.row_len] : 1.f;
        for (int kk = 0; kk < kn; ++kk) tile[kk][r][c] = ok ? d.v[src + (int64_t)kk * d.sk] * sc_row : 0.f;
      }
      __syncthreads();
      if (c0 + l < Cn) {
        for (int kk = 0; kk < kn; ++kk)
          for (int r = w8; r < PT && r0 + r < Rn; r += 8) {
            const int64_t o = poff(kk, r, l);
            if (dtype == ARTIC_BF16) reinterpret_cast<__nv_bfloat16*>(outp)[o] = __float2bfloat16_rn(tile[kk][r][l]);
            else reinterpret_cast<float*>(outp)[o] = tile[kk][r][l];
          }
      }
    } else {
      if (c0 + l < Cn) {
        for (int kk = 0; kk < kn; ++kk)
          for (int r = w8; r < PT && r0 + r < Rn; r += 8) tile[kk][r][l] = d.dWp[poff(kk, r, l)];
      }
      __syncthreads();
      for (int o = w8; o < PT; o += 8) {
        const int r = r_inner ? l : o, c = r_inner ? o : l;
        if (r0 + r < Rn && c0 + c < Cn) {
          const int64_t dst = tbase + (int64_t)(r0 + r) * sr + (int64_t)(c0 + c) * sc;
          for (int kk = 0; kk < kn; ++kk) d.dv[dst + (int64_t)kk * d.sk] = tile[kk][r][c];
        }
      }
    }
    __syncthreads();
  }
  }
}

// Weight-norm backward, in place on dv (which holds dL/dw in torch layout):
//   dot = <dw_row, v_row>;  dv = s * dw - (s * dot / nrm^2) * v;  dg = dot / nrm     (s = g / nrm)
__global__ void __launch_bounds__(256) wn_bwd_kernel(const artic_wdesc_t* __restrict__ descs) {
  __shared__ float red[32];
  __shared__ float s_dot;
  const artic_wdesc_t& d = descs[blockIdx.y];
  if (d.g == nullptr || d.dv == nullptr) return;
  for (int row = blockIdx.x; row < d.rows; row += gridDim.x) {
    const int64_t e0 = (int64_t)row * d.row_len;
    float dot = 0.f;
    for (int64_t e = threadIdx.x; e < d.row_len; e += blockDim.x) dot = fmaf(d.dv[e0 + e], d.v[e0 + e], dot);
    dot = block_sum(dot, red);
    if (threadIdx.x == 0) s_dot = dot;
    __syncthreads();
    dot = s_dot;
    const float s = d.scale[row], nrm = d.scale[d.rows + row];
    const float coef = s * dot / (nrm * nrm);
    for (int64_t e = threadIdx.x; e < d.row_len; e += blockDim.x) d.dv[e0 + e] = s * d.dv[e0 + e] - coef * d.v[e0 + e];
    if (threadIdx.x == 0) d.dg[row] = dot / nrm;
    __syncthreads();
  }
}

// ---- lean versions (the default) ----------------------------------------------------------
// wprep_kernel: ONE pass over the torch weight writes BOTH prepared layouts ('fwd' [K][G/m][a_pad][b_pad]
// and 'bwd' [K][G/m][b_pad][a_pad]); wunprep_kernel: the prepared fp32 gradient back to the torch layout.
// Same flat tile space as wperm_kernel ([<= 8 taps][32 a][32 b] per tile).  Against the generic three-pass
// version (debug key 12): 32-bit index arithmetic with per-tile bases (no 64-bit multiplies / divisions per
// element), the torch side of a tile moved as CONTIGUOUS runs of 32 * K floats whenever a tile holds all taps
// (K <= 8), bf16 pairs stored as 32-bit words, the descriptor read once per tile.  Measured on the
// discriminator (70.7 M parameters, tools/weights_bench.py): prep 705 -> 315 us, unprep + weight-norm backward
// 626 -> 480 us; still instruction- rather than HBM-bound (1.8 / 2.9 TB/s, profiles/r1_weights_bench_final.log).
constexpr int TS_A = PT + 1;                    // smem row pitch (floats)
constexpr int TS_K = PT * TS_A;                 // tap pitch, plus a per-layer skew (below)

struct TileGeom {
  uint32_t a0, b0, k0, kn, g;
  uint32_t tb;            // torch offset of (g, k0)
  uint32_t base_f, base_b, kstride;   // prepared-side offsets of the tile origin in the two layouts, tap stride
  uint32_t ks;            // smem tap pitch: TS_K + skew, skew = ceil(32 / K) keeps the run accesses (lanes along
                          // (inner, tap)) nearly bank-conflict free
};

__device__ __forceinline__ int find_layer(const artic_wdesc_t* __restrict__ descs, int n_layers, long long gt) {
  int lo = 0, hi = n_layers - 1;                       // last layer with tile_begin <= gt
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&descs[mid].tile_begin) <= gt) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ TileGeom tile_geom(const artic_wdesc_t& d, long long gt) {
  TileGeom t;
  const uint32_t kc = d.K < PK ? d.K : PK;
  const uint32_t n_at = (d.A + PT - 1) / PT, n_bt = (d.B + PT - 1) / PT, n_kt = (d.K + kc - 1) / kc;
  uint32_t w = (uint32_t)(gt - d.tile_begin);
  const uint32_t bt = w % n_bt; w /= n_bt;
  const uint32_t at = w % n_at; w /= n_at;
  const uint32_t kt = w % n_kt;
  t.g = w / n_kt;
  t.a0 = at * PT; t.b0 = bt * PT; t.k0 = kt * kc;
  t.kn = min(kc, (uint32_t)d.K - t.k0);
  t.tb = t.g * (uint32_t)d.sg + t.k0 * (uint32_t)d.sk;
  const uint32_t m = d.merge, Gs = d.G / m, gm = t.g % m, gd = t.g / m;
  t.kstride = Gs * (uint32_t)d.a_pad * (uint32_t)d.b_pad;
  t.base_f = ((t.k0 * Gs + gd) * d.a_pad + gm * d.A + t.a0) * d.b_pad + gm * d.B + t.b0;
  t.base_b = ((t.k0 * Gs + gd) * d.b_pad + gm * d.B + t.b0) * d.a_pad + gm * d.A + t.a0;
  t.ks = TS_K + (32 + d.K - 1) / d.K;
  return t;
}

__device__ __forceinline__ void store_out(void* out, int dtype, uint32_t o, float x) {
  if (dtype == ARTIC_BF16) reinterpret_cast<__nv_bfloat16*>(out)[o] = __float2bfloat16_rn(x);
  else reinterpret_cast<float*>(out)[o] = x;
}

// Fast path of both kernels (a tile holds all K <= 8 taps and the taps are the torch-inner index): the torch side
// of the tile is 32 runs of n_in * K CONTIGUOUS floats; they are copied to / from shared memory as they are
// (row pitch 32 * K + 1 floats), and the prepared side addresses element (a, b, kk) at a * sA + b * sB + kk.
__device__ __forceinline__ bool tile_is_runs(const artic_wdesc_t& d, const TileGeom& t, uint32_t s_in, uint32_t s_out) {
  return t.kn == (uint32_t)d.K && d.sk == 1 && s_in == (uint32_t)d.K &&
         (d.g == nullptr || (uint32_t)d.row_len % s_out == 0);   // a run stays inside ONE weight-norm row
}

__global__ void __launch_bounds__(256) wprep_kernel(const artic_wdesc_t* __restrict__ descs, int n_layers,
                                                    long long total_tiles) {
  __shared__ float tile[PK * (TS_K + 32)];
  const int l = threadIdx.x & 31, w8 = threadIdx.x >> 5;
  for (long long gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
    const artic_wdesc_t d = descs[find_layer(descs, n_layers, gt)];
    if (d.out_f == nullptr && d.out_b == nullptr) continue;
    const TileGeom t = tile_geom(d, gt);
    const uint32_t A = d.A, B = d.B, K = d.K;
    const bool a_inner = d.sa < d.sb || (d.sa == d.sb && d.A == 1);   // tie (one channel): keep the other dim outer
    const uint32_t s_in = (uint32_t)(a_inner ? d.sa : d.sb), s_out = (uint32_t)(a_inner ? d.sb : d.sa);
    const uint32_t in0 = a_inner ? t.a0 : t.b0, out0 = a_inner ? t.b0 : t.a0;
    const uint32_t n_in = min((uint32_t)PT, (a_inner ? A : B) - in0), n_out = min((uint32_t)PT, (a_inner ? B : A) - out0);
    const uint32_t row_len = (uint32_t)d.row_len;
    const bool runs = tile_is_runs(d, t, s_in, s_out);
    uint32_t sA, sB, sK;            // smem strides of (a, b, kk)
    // ---- torch -> smem (weight-norm scale applied)
    if (runs) {
      const uint32_t pitch = PT * K + 1, run = n_in * K;
      sA = a_inner ? K : pitch; sB = a_inner ? pitch : K; sK = 1;
      // all of a warp's loads (4 rows x <= 2 vectors per lane) are issued before the first use: the pass is
      // bound by the bytes in flight per SM, not by instruction issue
      const bool vec = (run & 3) == 0 && (reinterpret_cast<uintptr_t>(d.v + t.tb + in0 * K + out0 * s_out) & 15) == 0 && (s_out & 3) == 0;
      if (vec) {
        float4 x[4][2];
        float sc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t o = w8 + 8 * i;
          const uint32_t src0 = t.tb + in0 * K + (out0 + o) * s_out;
          sc[i] = (o < n_out && d.g != nullptr) ? d.scale[src0 / row_len] : 1.f;
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const uint32_t e = 4 * l + 128 * it;
            x[i][it] = (o < n_out && e < run) ? __ldg(reinterpret_cast<const float4*>(d.v + src0 + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const uint32_t o = w8 + 8 * i;
          float* dst = tile + o * pitch;
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const uint32_t e = 4 * l + 128 * it;
            if (o < n_out && e < run) {
              dst[e] = x[i][it].x * sc[i]; dst[e + 1] = x[i][it].y * sc[i]; dst[e + 2] = x[i][it].z * sc[i]; dst[e + 3] = x[i][it].w * sc[i];
            }
          }
        }
      } else {
        for (uint32_t o = w8; o < n_out; o += 8) {
          const uint32_t src0 = t.tb + in0 * K + (out0 + o) * s_out;
          const float sc = d.g != nullptr ? d.scale[src0 / row_len] : 1.f;
          const float* src = d.v + src0;
          float* dst = tile + o * pitch;
#pragma unroll 4
          for (uint32_t e = l; e < run; e += 32) dst[e] = __ldg(src + e) * sc;
        }
      }
    } else {
      sA = TS_A; sB = 1; sK = t.ks;
      const uint32_t sm_in = a_inner ? TS_A : 1, sm_out = a_inner ? 1 : TS_A;
      for (uint32_t o = w8; o < PT; o += 8) {
        if (o < n_out && (uint32_t)l < n_in) {
          const uint32_t src = t.tb + (in0 + l) * s_in + (out0 + o) * s_out;
          const float sc = d.g != nullptr ? d.scale[src / row_len] : 1.f;
          const float* vp = d.v + src;
          float* tp = tile + l * sm_in + o * sm_out;
          float xv[PK];
#pragma unroll
          for (uint32_t kk = 0; kk < PK; ++kk) xv[kk] = kk < t.kn ? __ldg(vp + kk * (uint32_t)d.sk) : 0.f;
#pragma unroll
          for (uint32_t kk = 0; kk < PK; ++kk)
            if (kk < t.kn) tp[kk * t.ks] = xv[kk] * sc;
        }
      }
    }
    __syncthreads();
    const uint32_t na = min((uint32_t)PT, A - t.a0), nb = min((uint32_t)PT, B - t.b0);
    const uint32_t j = threadIdx.x & 15, rr = threadIdx.x >> 4;       // pair index along the run, row
    // ---- 'fwd' layout: rows a, runs along b
    if (d.out_f != nullptr) {
      const bool pairs = d.dtype_f == ARTIC_BF16 && !(d.b_pad & 1) && !(t.base_f & 1);
      if (pairs) {
        const uint32_t b = 2 * j;
        if (b < nb) {
          const bool two = b + 1 < nb;
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out_f) + t.base_f + b;
          for (uint32_t a = rr; a < na; a += 16) {
            const float* s0 = tile + a * sA + b * sB;
            __nv_bfloat16* o = out + a * d.b_pad;
#pragma unroll 4
            for (uint32_t kk = 0; kk < t.kn; ++kk, s0 += sK, o += t.kstride) {
              if (two) *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(s0[0], s0[sB]);
              else *o = __float2bfloat16_rn(s0[0]);
            }
          }
        }
      } else {
        for (uint32_t kk = 0; kk < t.kn; ++kk)
          for (uint32_t a = w8; a < na; a += 8)
            if ((uint32_t)l < nb)
              store_out(d.out_f, d.dtype_f, t.base_f + kk * t.kstride + a * d.b_pad + l, tile[kk * sK + a * sA + l * sB]);
      }
    }
    // ---- 'bwd' layout: rows b, runs along a
    if (d.out_b != nullptr) {
      const bool pairs = d.dtype_b == ARTIC_BF16 && !(d.a_pad & 1) && !(t.base_b & 1);
      if (pairs) {
        const uint32_t a = 2 * j;
        if (a < na) {
          const bool two = a + 1 < na;
          __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(d.out_b) + t.base_b + a;
          for (uint32_t b = rr; b < nb; b += 16) {
            const float* s0 = tile + a * sA + b * sB;
            __nv_bfloat16* o = out + b * d.a_pad;
#pragma unroll 4
            for (uint32_t kk = 0; kk < t.kn; ++kk, s0 += sK, o += t.kstride) {
              if (two) *reinterpret_cast<__nv_bfloat162*>(o) = __floats2bfloat162_rn(s0[0], s0[sA]);
              else *o = __float2bfloat16_rn(s0[0]);
            }
          }
        }
      } else {
        for (uint32_t kk = 0; kk < t.kn; ++kk)
          for (uint32_t b = w8; b < nb; b += 8)
            if ((uint32_t)l < na)
              store_out(d.out_b, d.dtype_b, t.base_b + kk * t.kstride + b * d.a_pad + l, tile[kk * sK + l * sA + b * sB]);
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) wunprep_kernel(const artic_wdesc_t* __restrict__ descs, int n_layers,
                                                      long long total_tiles) {
  __shared__ float tile[PK * (TS_K + 32)];
  const int l = threadIdx.x & 31, w8 = threadIdx.x >> 5;
  for (long long gt = blockIdx.x; gt < total_tiles; gt += gridDim.x) {
    const artic_wdesc_t d = descs[find_layer(descs, n_layers, gt)];
    if (d.dv == nullptr || d.dWp == nullptr) continue;
    const TileGeom t = tile_geom(d, gt);
    const uint32_t A = d.A, B = d.B, K = d.K;
    const uint32_t na = min((uint32_t)PT, A - t.a0), nb = min((uint32_t)PT, B - t.b0);
    const bool a_inner = d.sa < d.sb || (d.sa == d.sb && d.A == 1);
    const uint32_t s_in = (uint32_t)(a_inner ? d.sa : d.sb), s_out = (uint32_t)(a_inner ? d.sb : d.sa);
    const uint32_t in0 = a_inner ? t.a0 : t.b0, out0 = a_inner ? t.b0 : t.a0;
    const uint32_t n_in = a_inner ? na : nb, n_out = a_inner ? nb : na;
    const bool runs = t.kn == K && d.sk == 1 && s_in == K;
    const uint32_t pitch = PT * K + 1;
    const uint32_t sA = runs ? (a_inner ? K : pitch) : TS_A, sB = runs ? (a_inner ? pitch : K) : 1, sK = runs ? 1 : t.ks;
    // ---- prepared gradient -> smem
    if (!d.dw_swapped) {
      if ((uint32_t)l < nb) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {      // 2 rows x 8 taps of loads in flight per lane
          float xv[2][PK];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t a = w8 + 8 * (2 * half + i);
            const float* src = d.dWp + t.base_f + a * d.b_pad + l;
#pragma unroll
            for (uint32_t kk = 0; kk < PK; ++kk) xv[i][kk] = (a < na && kk < t.kn) ? __ldg(src + kk * t.kstride) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t a = w8 + 8 * (2 * half + i);
            float* tp = tile + a * sA + l * sB;
#pragma unroll
            for (uint32_t kk = 0; kk < PK; ++kk)
              if (a < na && kk < t.kn) tp[kk * sK] = xv[i][kk];
          }
        }
      }
    } else {
      if ((uint32_t)l < na) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float xv[2][PK];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t b = w8 + 8 * (2 * half + i);
            const float* src = d.dWp + t.base_b + b * d.a_pad + l;
#pragma unroll
            for (uint32_t kk = 0; kk < PK; ++kk) xv[i][kk] = (b < nb && kk < t.kn) ? __ldg(src + kk * t.kstride) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const uint32_t b = w8 + 8 * (2 * half + i);
            float* tp = tile + l * sA + b * sB;
#pragma unroll
            for (uint32_t kk = 0; kk < PK; ++kk)
              if (b < nb && kk < t.kn) tp[kk * sK] = xv[i][kk];
          }
        }
      }
    }
    __syncthreads();
    // ---- smem -> torch layout
    if (runs) {
      const uint32_t run = n_in * K;
      for (uint32_t o = w8; o < n_out; o += 8) {
        float* dst = d.dv + t.tb + in0 * K + (out0 + o) * s_out;
        const float* src = tile + o * pitch;
        if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (run & 3) == 0) {
          for (uint32_t e = 4 * l; e < run; e += 128)
            *reinterpret_cast<float4*>(dst + e) = make_float4(src[e], src[e + 1], src[e + 2], src[e + 3]);
        } else {
          for (uint32_t e = l; e < run; e += 32) dst[e] = src[e];
        }
      }
    } else {
      const uint32_t sm_in = a_inner ? TS_A : 1, sm_out = a_inner ? 1 : TS_A;
      for (uint32_t o = w8; o < n_out; o += 8)
        if ((uint32_t)l < n_in) {
          float* dst = d.dv + t.tb + (in0 + l) * s_in + (out0 + o) * s_out;
          const float* tp = tile + l * sm_in + o * sm_out;
          for (uint32_t kk = 0; kk < t.kn; ++kk) dst[kk * (uint32_t)d.sk] = tp[kk * t.ks];
        }
    }
    __syncthreads();
  }
}

// Vector versions of the two weight-norm row kernels: rows are read with 128-bit loads when the row
// length and base allow, and the backward keeps the row in registers between the dot product and the update.
__global__ void __launch_bounds__(256) wn_scale4_kernel(const artic_wdesc_t* __restrict__ descs) {
  __shared__ float red[32];
  const artic_wdesc_t& d = descs[blockIdx.y];
  if (d.g == nullptr) return;
  const int64_t row_len = d.row_len;
  const bool vec = (row_len & 3) == 0 && (reinterpret_cast<uintptr_t>(d.v) & 15) == 0;
  for (int row = blockIdx.x; row < d.rows; row += gridDim.x) {
    const float* vr = d.v + (int64_t)row * row_len;
    float s = 0.f;
    if (vec) {
      const float4* v4 = reinterpret_cast<const float4*>(vr);
      for (int64_t e = threadIdx.x; e < (row_len >> 2); e += blockDim.x) {
        const float4 x = __ldg(v4 + e);
        s = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, s))));
      }
    } else {
      for (int64_t e = threadIdx.x; e < row_len; e += blockDim.x) {
        const float x = vr[e];
        s = fmaf(x, x, s);
      }
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
      const float nrm = sqrtf(s);
      d.scale[row] = d.g[row] / nrm;
      d.scale[d.rows + row] = nrm;
    }
    __syncthreads();
  }
}

constexpr int WNB_V = 8;     // float4 per thread kept in registers (rows up to 256 * 4 * 8 = 8192 elements)
__global__ void __launch_bounds__(256) wn_bwd4_kernel(const artic_wdesc_t* __restrict__ descs) {
  __shared__ float red[32];
  __shared__ float s_dot;
  const artic_wdesc_t& d = descs[blockIdx.y];
  if (d.g == nullptr || d.dv == nullptr) return;
  const int64_t row_len = d.row_len;
  const bool vec = (row_len & 3) == 0 && row_len <= 256 * 4 * WNB_V &&
                   ((reinterpret_cast<uintptr_t>(d.v) | reinterpret_cast<uintptr_t>(d.dv)) & 15) == 0;
  for (int row = blockIdx.x; row < d.rows; row += gridDim.x) {
    const int64_t e0 = (int64_t)row * row_len;
    const float s = d.scale[row], nrm = d.scale[d.rows + row];
    if (vec) {
      const float4* v4 = reinterpret_cast<const float4*>(d.v + e0);
      float4* g4 = reinterpret_cast<float4*>(d.dv + e0);
      const int n4 = (int)(row_len >> 2);
      float4 gv[WNB_V], vv[WNB_V];
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < WNB_V; ++i) {
        const int e = threadIdx.x + i * 256;
        if (e < n4) {
          gv[i] = g4[e]; vv[i] = __ldg(v4 + e);
          dot = fmaf(gv[i].x, vv[i].x, fmaf(gv[i].y, vv[i].y, fmaf(gv[i].z, vv[i].z, fmaf(gv[i].w, vv[i].w, dot))));
        }
      }
      dot = block_sum(dot, red);
      if (threadIdx.x == 0) s_dot = dot;
      __syncthreads();
      dot = s_dot;
      const float coef = s * dot / (nrm * nrm);
#pragma unroll
      for (int i = 0; i < WNB_V; ++i) {
        const int e = threadIdx.x + i * 256;
        if (e < n4)
          g4[e] = make_float4(s * gv[i].x - coef * vv[i].x, s * gv[i].y - coef * vv[i].y, s * gv[i].z - coef * vv[i].z,
                              s * gv[i].w - coef * vv[i].w);
      }
      if (threadIdx.x == 0) d.dg[row] = dot / nrm;
      __syncthreads();
    } else {
      float dot = 0.f;
      for (int64_t e = threadIdx.x; e < row_len; e += blockDim.x) dot = fmaf(d.dv[e0 + e], d.v[e0 + e], dot);
      dot = block_sum(dot, red);
      if (threadIdx.x == 0) s_dot = dot;
      __syncthreads();
      dot = s_dot;
      const float coef = s * dot / (nrm * nrm);
      for (int64_t e = threadIdx.x; e < row_len; e += blockDim.x) d.dv[e0 + e] = s * d.dv[e0 + e] - coef * d.v[e0 + e];
      if (threadIdx.x == 0) d.dg[row] = dot / nrm;
      __syncthreads();
    }
  }
}

// ---- row-run relayout kernels (the default) ------------------------------------------------------------------
// The torch weight of a conv / linear / transposed-conv layer is, per group, a matrix [outer][inner * K] whose rows are
// CONTIGUOUS (outer = torch dim 0 = the weight-norm row; conv: outer = out-ch, inner = in-ch; convT: outer = in-ch,
// inner = out-ch; the taps are the fastest index).  A tile is 32 outer rows x TI inner columns x ALL K taps
// (TI * K <= 352 floats, TI = 32 for K <= 11, 8 for K = 41): the torch side moves as whole-row runs with 128-bit
// accesses, the prepared side as 16-byte items (8 bf16 / 4 fp32) along the prepared layout's fastest index:
//   X1 = the outer-fastest layout  (conv: 'fwd' [k][a][b];  convT: 'bwd' [k][b][a])   item = (k, i, 8 outer)
//   X2 = the inner-fastest layout  (conv: 'bwd' [k][b][a];  convT: 'fwd' [k][a][b])   item = (k, o, 8 inner)
// Shared-memory tile: row pitch odd (run + 1), which makes both item gathers bank-conflict free for odd K.
// Against wprep_kernel / wunprep_kernel (32 x 32 x <= 8-tap tiles, 4-byte stores, per-element index arithmetic) this
// issues ~4x fewer instructions per element (ncu: the old kernels were issue-bound at 47 % of peak issue rate with
// 25 % of DRAM throughput).
constexpr int RT_O = 32;
constexpr int RT_RUN = 352;
constexpr int RT_SMEM = RT_O * (RT_RUN + 1);
constexpr int RT_CHUNK_MAX = 8;      // inner tiles per work unit (artic_wdesc_t.row_chunk: the unit's index arithmetic is done once)
constexpr int RT_WPL = RT_RUN / 32;  // 32-bit words per lane per row

__host__ __device__ __forceinline__ int rt_inner(int K) {
  int ti = 32;
  while (ti > 1 && ti * K > RT_RUN) ti >>= 1;
  return ti;
}

struct RUnit {                      // one work unit: 32 outer rows x up to RT_CHUNK inner tiles of one group
  uint32_t O, I, TI, K, g, o0, no, it0, it1, pitch, s_out, kstride;
  uint32_t tb0;                     // torch offset of (g, o0, inner 0)
  uint32_t base_f0, base_b0;        // prepared offsets of (k = 0, a/b origin of the unit with inner index 0)
  bool a_inner;
};

struct RTile {
  uint32_t no, ni, run, pitch;      // live outer rows / inner columns, ni * K, smem row pitch
  uint32_t tb, s_out;               // torch offset of the tile's first row, torch row stride
  uint32_t base1, pitch1, base2, pitch2, kstride;
  uint32_t lg_ti;                   // log2(TI) when the tile is full (32 outer x TI inner), else 0xffffffff
};

__device__ __forceinline__ RUnit rt_unit(const artic_wdesc_t& d, long long gt) {
  RUnit u;
  u.K = (uint32_t)d.K;
  u.a_inner = d.sa < d.sb || (d.sa == d.sb && d.A == 1);
  u.O = u.a_inner ? d.B : d.A; u.I = u.a_inner ? d.A : d.B;
  u.TI = (uint32_t)rt_inner(d.K);
  const uint32_t ch = (uint32_t)max(1, min(d.row_chunk, RT_CHUNK_MAX));
  const uint32_t n_it = (u.I + u.TI - 1) / u.TI, n_ic = (n_it + ch - 1) / ch, n_ot = (u.O + RT_O - 1) / RT_O;
  uint32_t w = (uint32_t)(gt - d.tile2_begin);
  const uint32_t ic = w % n_ic; w /= n_ic;
  const uint32_t ot = w % n_ot;
  u.g = w / n_ot;
  u.o0 = ot * RT_O;
  u.no = min((uint32_t)RT_O, u.O - u.o0);
  u.it0 = ic * ch; u.it1 = min(n_it, u.it0 + ch);
  u.pitch = (u.TI * u.K) | 1u;
  u.s_out = (uint32_t)(u.a_inner ? d.sb : d.sa);
  u.tb0 = u.g * (uint32_t)d.sg + u.o0 * u.s_out;
  const uint32_t m = d.merge, Gs = d.G / m, gm = u.g % m, gd = u.g / m;
  u.kstride = Gs * (uint32_t)d.a_pad * (uint32_t)d.b_pad;
  const uint32_t a0 = u.a_inner ? 0 : u.o0, b0 = u.a_inner ? u.o0 : 0;
  u.base_f0 = (gd * d.a_pad + gm * d.A + a0) * d.b_pad + gm * d.B + b0;   // [k][a][b]
  u.base_b0 = (gd * d.b_pad + gm * d.B + b0) * d.a_pad + gm * d.A + a0;   // [k][b][a]
  return u;
}

__device__ __forceinline__ RTile rt_tile(const artic_wdesc_t& d, const RUnit& u, uint32_t it) {
  RTile t;
  const uint32_t i0 = it * u.TI;
  t.no = u.no;
  t.ni = min(u.TI, u.I - i0);
  t.run = t.ni * u.K;
  t.pitch = u.pitch;
  t.s_out = u.s_out;
  t.tb = u.tb0 + i0 * u.K;
  t.kstride = u.kstride;
  if (u.a_inner) { t.base1 = u.base_f0 + i0 * d.b_pad; t.pitch1 = d.b_pad; t.base2 = u.base_b0 + i0; t.pitch2 = d.a_pad; }
  else           { t.base1 = u.base_b0 + i0 * d.a_pad; t.pitch1 = d.a_pad; t.base2 = u.base_f0 + i0; t.pitch2 = d.b_pad; }
  t.lg_ti = (t.no == RT_O && t.ni == u.TI) ? (uint32_t)(31 - __clz((int)u.TI)) : 0xffffffffu;
  return t;
}

__device__ __forceinline__ int rt_find(const artic_wdesc_t* __restrict__ descs, int n_layers, long long gt) {
  int lo = 0, hi = n_layers - 1;                       // last layer with tile2_begin <= gt
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(&descs[mid].tile2_begin) <= gt) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// One 16-byte item of a prepared layout: E consecutive elements read from shared memory at stride `ss`
template <typename T> struct RtItem;
template <> struct RtItem<__nv_bfloat16> {
  static constexpr int E = 8, LG = 3;
  static __device__ __forceinline__ void store_full(__nv_bfloat16* dst, const float* src, uint32_t ss) {
    uint4 q;
    q.x = pack_bf2(src[0], src[ss]); q.y = pack_bf2(src[2 * ss], src[3 * ss]);
    q.z = pack_bf2(src[4 * ss], src[5 * ss]); q.w = pack_bf2(src[6 * ss], src[7 * ss]);
    *reinterpret_cast<uint4*>(dst) = q;
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* dst, const float* src, uint32_t ss, uint32_t n) {
    if (n == 8 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      uint4 q;
      q.x = pack_bf2(src[0], src[ss]); q.y = pack_bf2(src[2 * ss], src[3 * ss]);
      q.z = pack_bf2(src[4 * ss], src[5 * ss]); q.w = pack_bf2(src[6 * ss], src[7 * ss]);
      *reinterpret_cast<uint4*>(dst) = q;
    } else {
      for (uint32_t j = 0; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j * ss]);
    }
  }
};
template <> struct RtItem<float> {
  static constexpr int E = 4, LG = 2;
  static __device__ __forceinline__ void store_full(float* dst, const float* src, uint32_t ss) {
    *reinterpret_cast<float4*>(dst) = make_float4(src[0], src[ss], src[2 * ss], src[3 * ss]);
  }
  static __device__ __forceinline__ void store(float* dst, const float* src, uint32_t ss, uint32_t n) {
    if (n == 4 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(src[0], src[ss], src[2 * ss], src[3 * ss]);
    } else {
      for (uint32_t j = 0; j < n; ++j) dst[j] = src[j * ss];
    }
  }
};

// tile (shared memory, [o][i * K + k]) -> X1 (outer-fastest): items (k, i, group of E outer rows), groups fastest
template <typename T>
__device__ __forceinline__ void rt_store_x1(T* out, const float* tile, const RTile& t, uint32_t K) {
  constexpr uint32_t E = RtItem<T>::E;
  if (t.lg_ti != 0xffffffffu) {                       // full tile: shifts instead of divisions
    constexpr uint32_t LGG = 5 - RtItem<T>::LG;       // log2(groups per (k, i)) = log2(32 / E)
    const uint32_t total = K << (t.lg_ti + LGG);
    // every item 16-byte aligned?  (uniform: decided once per tile)
    const bool al = ((t.base1 | t.kstride | t.pitch1) & (E - 1)) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (al) {
      for (uint32_t q = threadIdx.x; q < total; q += 256) {
        const uint32_t o = (q & ((1u << LGG) - 1)) * E, i = (q >> LGG) & ((1u << t.lg_ti) - 1), k = q >> (LGG + t.lg_ti);
        RtItem<T>::store_full(out + t.base1 + k * t.kstride + i * t.pitch1 + o, tile + o * t.pitch + i * K + k, t.pitch);
      }
    } else {
      for (uint32_t q = threadIdx.x; q < total; q += 256) {
        const uint32_t o = (q & ((1u << LGG) - 1)) * E, i = (q >> LGG) & ((1u << t.lg_ti) - 1), k = q >> (LGG + t.lg_ti);
        RtItem<T>::store(out + t.base1 + k * t.kstride + i * t.pitch1 + o, tile + o * t.pitch + i * K + k, t.pitch, E);
      }
    }
    return;
  }
  const uint32_t ng = (t.no + E - 1) / E;
  const uint32_t per_k = t.ni * ng, total = K * per_k;
  for (uint32_t q = threadIdx.x; q < total; q += 256) {
    const uint32_t k = q / per_k, r = q - k * per_k;
    const uint32_t i = r / ng, og = r - i * ng;
    const uint32_t o = og * E;
    RtItem<T>::store(out + t.base1 + k * t.kstride + i * t.pitch1 + o, tile + o * t.pitch + i * K + k, t.pitch, min(E, t.no - o));
  }
}

// tile -> X2 (inner-fastest): items (k, o, group of E inner columns), groups fastest
template <typename T>
__device__ __forceinline__ void rt_store_x2(T* out, const float* tile, const RTile& t, uint32_t K) {
  constexpr uint32_t E = RtItem<T>::E;
  if (t.lg_ti != 0xffffffffu && t.lg_ti >= RtItem<T>::LG) {
    const uint32_t lgg = t.lg_ti - RtItem<T>::LG;     // log2(groups per (k, o))
    const uint32_t total = K << (lgg + 5);
    const bool al = ((t.base2 | t.kstride | t.pitch2) & (E - 1)) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (al) {
      for (uint32_t q = threadIdx.x; q < total; q += 256) {
        const uint32_t i = (q & ((1u << lgg) - 1)) * E, o = (q >> lgg) & 31u, k = q >> (lgg + 5);
        RtItem<T>::store_full(out + t.base2 + k * t.kstride + o * t.pitch2 + i, tile + o * t.pitch + i * K + k, K);
      }
    } else {
      for (uint32_t q = threadIdx.x; q < total; q += 256) {
        const uint32_t i = (q & ((1u << lgg) - 1)) * E, o = (q >> lgg) & 31u, k = q >> (lgg + 5);
        RtItem<T>::store(out + t.base2 + k * t.kstride + o * t.pitch2 + i, tile + o * t.pitch + i * K + k, K, E);
      }
    }
    return;
  }
  const uint32_t ng = (t.ni + E - 1) / E;
  const uint32_t per_k = t.no * ng, total = K * per_k;
  for (uint32_t q = threadIdx.x; q < total; q += 256) {
    const uint32_t k = q / per_k, r = q - k * per_k;
    const uint32_t o = r / ng, ig = r - o * ng;
    const uint32_t i = ig * E;
    RtItem<T>::store(out + t.base2 + k * t.kstride + o * t.pitch2 + i, tile + o * t.pitch + i * K + k, K, min(E, t.ni - i));
  }
}

// Rows w8, w8 + 8, w8 + 16, w8 + 24 of the tile, NW words per lane and row; the loads of two rows are issued before
// their first use.
template <int NW>
__device__ __forceinline__ void rt_load_rows(float* tile, const float* __restrict__ v, const RTile& t, const float (&sc)[4],
                                             uint32_t lane, uint32_t w8) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float x[2][NW];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t r = w8 + 8 * (2 * h + i);
      const float* src = v + t.tb + r * t.s_out;
      const uint32_t lim = r < t.no ? t.run : 0u;
#pragma unroll
      for (int j = 0; j < NW; ++j) {
        const uint32_t e = lane + 32 * j;
        x[i][j] = e < lim ? __ldg(src + e) : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const uint32_t r = w8 + 8 * (2 * h + i);
      float* dst = tile + r * t.pitch;
      const uint32_t lim = r < t.no ? t.run : 0u;
#pragma unroll
      for (int j = 0; j < NW; ++j) {
        const uint32_t e = lane + 32 * j;
        if (e < lim) dst[e] = x[i][j] * sc[2 * h + i];
      }
    }
  }
}

__global__ void __launch_bounds__(256) wprep_rows_kernel(const artic_wdesc_t* __restrict__ descs, int n_layers,
                                                         long long total_units) {
  __shared__ float tile[RT_SMEM];
  const uint32_t lane = threadIdx.x & 31, w8 = threadIdx.x >> 5;
  for (long long gt = blockIdx.x; gt < total_units; gt += gridDim.x) {
    const artic_wdesc_t& d = descs[rt_find(descs, n_layers, gt)];
    const RUnit u = rt_unit(d, gt);
    const float* __restrict__ v = d.v;
    const float* __restrict__ scale = d.g != nullptr ? d.scale : nullptr;
    void* p1 = u.a_inner ? d.out_f : d.out_b;
    void* p2 = u.a_inner ? d.out_b : d.out_f;
    const int dt1 = u.a_inner ? d.dtype_f : d.dtype_b, dt2 = u.a_inner ? d.dtype_b : d.dtype_f;
    // weight-norm scale of this warp's four rows (a row of the tile lies inside ONE torch row)
    float sc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t r = w8 + 8 * i;
      sc[i] = (r < u.no && scale != nullptr) ? __ldg(scale + (u.tb0 + r * u.s_out) / (uint32_t)d.row_len) : 1.f;
    }
    for (uint32_t it = u.it0; it < u.it1; ++it) {
      const RTile t = rt_tile(d, u, it);
      // ---- torch rows -> shared memory, lane-consecutive words (conflict-free stores)
      const uint32_t nw = (t.run + 31) >> 5;
      if (nw <= 3) rt_load_rows<3>(tile, v, t, sc, lane, w8);
      else if (nw <= 5) rt_load_rows<5>(tile, v, t, sc, lane, w8);
      else if (nw <= 7) rt_load_rows<7>(tile, v, t, sc, lane, w8);
      else rt_load_rows<RT_WPL>(tile, v, t, sc, lane, w8);
      __syncthreads();
      if (p1 != nullptr) {
        if (dt1 == ARTIC_BF16) rt_store_x1(reinterpret_cast<__nv_bfloat16*>(p1), tile, t, u.K);
        else rt_store_x1(reinterpret_cast<float*>(p1), tile, t, u.K);
      }
      if (p2 != nullptr) {
        if (dt2 == ARTIC_BF16) rt_store_x2(reinterpret_cast<__nv_bfloat16*>(p2), tile, t, u.K);
        else rt_store_x2(reinterpret_cast<float*>(p2), tile, t, u.K);
      }
      __syncthreads();
    }
  }
}

// Prepared fp32 gradient -> torch layout.  The gradient is in the 'fwd' layout, or in the 'bwd' layout for
// dw_swapped layers: outer-fastest (X1) for conv / linear and for swapped transposed convs, inner-fastest (X2) for
// transposed convs in the CUDA-core modes.
__global__ void __launch_bounds__(256) wunprep_rows_kernel(const artic_wdesc_t* __restrict__ descs, int n_layers,
                                                           long long total_units) {
  __shared__ float tile[RT_SMEM];
  const uint32_t lane = threadIdx.x & 31, w8 = threadIdx.x >> 5;
  for (long long gt = blockIdx.x; gt < total_units; gt += gridDim.x) {
    const artic_wdesc_t& d = descs[rt_find(descs, n_layers, gt)];
    if (d.dv == nullptr || d.dWp == nullptr) continue;
    const RUnit u = rt_unit(d, gt);
    const uint32_t K = u.K;
    const float* __restrict__ src = d.dWp;
    float* __restrict__ dv = d.dv;
    // which of the two layouts holds the gradient, seen from the tile: 'fwd' = X1 iff a_inner
    const bool x1 = (d.dw_swapped == 0) == u.a_inner;
    for (uint32_t it = u.it0; it < u.it1; ++it) {
      const RTile t = rt_tile(d, u, it);
      if (x1) {
        if (t.lg_ti != 0xffffffffu) {
          const uint32_t total = K << (t.lg_ti + 3);               // items (k, i, 4 outer rows): 8 groups per (k, i)
          for (uint32_t q = threadIdx.x; q < total; q += 256) {
            const uint32_t o = (q & 7u) * 4, i = (q >> 3) & ((1u << t.lg_ti) - 1), k = q >> (3 + t.lg_ti);
            const float* p = src + t.base1 + k * t.kstride + i * t.pitch1 + o;
            float* dst = tile + o * t.pitch + i * K + k;
            if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
              const float4 x = __ldg(reinterpret_cast<const float4*>(p));
              dst[0] = x.x; dst[t.pitch] = x.y; dst[2 * t.pitch] = x.z; dst[3 * t.pitch] = x.w;
            } else {
              for (uint32_t j = 0; j < 4; ++j) dst[j * t.pitch] = __ldg(p + j);
            }
          }
        } else {
          const uint32_t ng = (t.no + 3) >> 2, per_k = t.ni * ng, total = K * per_k;
          for (uint32_t q = threadIdx.x; q < total; q += 256) {
            const uint32_t k = q / per_k, r = q - k * per_k;
            const uint32_t i = r / ng, o = (r - i * ng) * 4;
            const float* p = src + t.base1 + k * t.kstride + i * t.pitch1 + o;
            float* dst = tile + o * t.pitch + i * K + k;
            const uint32_t n = min(4u, t.no - o);
            if (n == 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
              const float4 x = __ldg(reinterpret_cast<const float4*>(p));
              dst[0] = x.x; dst[t.pitch] = x.y; dst[2 * t.pitch] = x.z; dst[3 * t.pitch] = x.w;
            } else {
              for (uint32_t j = 0; j < n; ++j) dst[j * t.pitch] = __ldg(p + j);
            }
          }
        }
      } else {
        const uint32_t ng = (t.ni + 3) >> 2, per_k = t.no * ng, total = K * per_k;
        for (uint32_t q = threadIdx.x; q < total; q += 256) {
          const uint32_t k = q / per_k, r = q - k * per_k;
          const uint32_t o = r / ng, i = (r - o * ng) * 4;
          const float* p = src + t.base2 + k * t.kstride + o * t.pitch2 + i;
          float* dst = tile + o * t.pitch + i * K + k;
          const uint32_t n = min(4u, t.ni - i);
          if (n == 4 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(p));
            dst[0] = x.x; dst[K] = x.y; dst[2 * K] = x.z; dst[3 * K] = x.w;
          } else {
            for (uint32_t j = 0; j < n; ++j) dst[j * K] = __ldg(p + j);
          }
        }
      }
      __syncthreads();
      // ---- shared memory -> torch rows, lane-consecutive words (conflict-free loads, 128-byte store segments)
      for (uint32_t r = w8; r < t.no; r += 8) {
        float* dst = dv + t.tb + r * t.s_out;
        const float* s = tile + r * t.pitch;
#pragma unroll 4
        for (uint32_t e = lane; e < t.run; e += 32) dst[e] = s[e];
      }
      __syncthreads();
    }
  }
}

// Warp-per-row versions (the default): no block barrier between a row's load, reduction and store phases, eight rows
// in flight per block, eight independent 128-bit loads in flight per lane.  The backward makes two passes over its row
// (dot product, then update); the second one hits L1 / L2, so DRAM still sees 12 bytes per element.

__global__ void __launch_bounds__(256) wn_scale_rows_kernel(const artic_wdesc_t* __restrict__ descs) {
  const artic_wdesc_t& d = descs[blockIdx.y];
  if (d.g == nullptr) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row_len = d.row_len;
  const bool vec = (row_len & 3) == 0 && (reinterpret_cast<uintptr_t>(d.v) & 15) == 0;
  for (int row = blockIdx.x * 8 + w; row < d.rows; row += gridDim.x * 8) {
    const float* vr = d.v + (int64_t)row * row_len;
    float s = 0.f;
    if (vec) {
      const float4* v4 = reinterpret_cast<const float4*>(vr);
      const int n4 = (int)(row_len >> 2);
      for (int e0 = 0; e0 < n4; e0 += 256) {
        float4 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = e0 + lane + 32 * i;
          x[i] = e < n4 ? __ldg(v4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(x[i].x, x[i].x, fmaf(x[i].y, x[i].y, fmaf(x[i].z, x[i].z, fmaf(x[i].w, x[i].w, s))));
      }
    } else {
      for (int64_t e = lane; e < row_len; e += 32) {
        const float x = __ldg(vr + e);
        s = fmaf(x, x, s);
      }
    }
    s = warp_sum(s);
    if (lane == 0) {
      const float nrm = sqrtf(s);
      d.scale[row] = d.g[row] / nrm;
      d.scale[d.rows + row] = nrm;
    }
  }
}

__global__ void __launch_bounds__(256) wn_bwd_rows_kernel(const artic_wdesc_t* __restrict__ descs) {
  const artic_wdesc_t& d = descs[blockIdx.y];
  if (d.g == nullptr || d.dv == nullptr) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row_len = d.row_len;
  const bool vec = (row_len & 3) == 0 && ((reinterpret_cast<uintptr_t>(d.v) | reinterpret_cast<uintptr_t>(d.dv)) & 15) == 0;
  for (int row = blockIdx.x * 8 + w; row < d.rows; row += gridDim.x * 8) {
    const int64_t e0 = (int64_t)row * row_len;
    const float s = d.scale[row], nrm = d.scale[d.rows + row];
    float dot = 0.f;
    if (vec) {
      const float4* v4 = reinterpret_cast<const float4*>(d.v + e0);
      float4* g4 = reinterpret_cast<float4*>(d.dv + e0);
      const int n4 = (int)(row_len >> 2);
      if (n4 <= 256) {                      // the row fits in registers (8 x 128 bits per lane per array): one pass
        float4 gv[8], vv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = lane + 32 * i;
          gv[i] = e < n4 ? g4[e] : make_float4(0.f, 0.f, 0.f, 0.f);
          vv[i] = e < n4 ? __ldg(v4 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) dot = fmaf(gv[i].x, vv[i].x, fmaf(gv[i].y, vv[i].y, fmaf(gv[i].z, vv[i].z, fmaf(gv[i].w, vv[i].w, dot))));
        dot = warp_sum(dot);
        const float coef = s * dot / (nrm * nrm);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int e = lane + 32 * i;
          if (e < n4)
            g4[e] = make_float4(s * gv[i].x - coef * vv[i].x, s * gv[i].y - coef * vv[i].y, s * gv[i].z - coef * vv[i].z,
                                s * gv[i].w - coef * vv[i].w);
        }
      } else {
        for (int c0 = 0; c0 < n4; c0 += 128) {
          float4 gv[4], vv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = c0 + lane + 32 * i;
            gv[i] = e < n4 ? g4[e] : make_float4(0.f, 0.f, 0.f, 0.f);
            vv[i] = e < n4 ? v4[e] : make_float4(0.f, 0.f, 0.f, 0.f);     // plain loads: the second pass re-reads them
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) dot = fmaf(gv[i].x, vv[i].x, fmaf(gv[i].y, vv[i].y, fmaf(gv[i].z, vv[i].z, fmaf(gv[i].w, vv[i].w, dot))));
        }
        dot = warp_sum(dot);
        const float coef = s * dot / (nrm * nrm);
        for (int c0 = 0; c0 < n4; c0 += 128) {
          float4 gv[4], vv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = c0 + lane + 32 * i;
            gv[i] = e < n4 ? g4[e] : make_float4(0.f, 0.f, 0.f, 0.f);
            vv[i] = e < n4 ? v4[e] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int e = c0 + lane + 32 * i;
            if (e < n4)
              g4[e] = make_float4(s * gv[i].x - coef * vv[i].x, s * gv[i].y - coef * vv[i].y, s * gv[i].z - coef * vv[i].z,
                                  s * gv[i].w - coef * vv[i].w);
          }
        }
      }
    } else {
      for (int64_t e = lane; e < row_len; e += 32) dot = fmaf(d.dv[e0 + e], d.v[e0 + e], dot);
      dot = warp_sum(dot);
      const float coef = s * dot / (nrm * nrm);
      for (int64_t e = lane; e < row_len; e += 32) d.dv[e0 + e] = s * d.dv[e0 + e] - coef * d.v[e0 + e];
    }
    if (lane == 0) d.dg[row] = dot / nrm;
  }
}

__device__ __forceinline__ float4 adam_ld4(const float* g, int64_t i) { return reinterpret_cast<const float4*>(g)[i]; }
__device__ __forceinline__ float4 adam_ld4(const __nv_bfloat16* g, int64_t i) {
  const uint2 q = reinterpret_cast<const uint2*>(g)[i];
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&q.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float adam_ld1(const float* g, int64_t i) { return g[i]; }
__device__ __forceinline__ float adam_ld1(const __nv_bfloat16* g, int64_t i) { return __bfloat162float(g[i]); }

// TG = float: the local gradient; TG = bf16: the data-parallel wire buffer (the reduced gradient is consumed as it came
// off the wire, without a cast back to fp32)
template <typename TG>
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const TG* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   const artic_adam_hyper_t* __restrict__ hyper) {
  const artic_adam_hyper_t h = *hyper;
  int passed = 0;
  for (int i = 0; i < h.n_milestones; ++i) passed += (h.step >= h.milestones[i]) ? 1 : 0;
  const float lr = h.lr0 * powf(h.gamma, (float)passed);
  const double t = (double)(h.step + 1);
  const float bc1 = (float)(1.0 - pow((double)h.beta1, t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)h.beta2, t));
  const float step_size = lr / bc1;
  // 128-bit accesses over the 16-byte aligned body (the flat buffers come from the torch allocator), scalar tail
  const int64_t n4 = (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                        reinterpret_cast<uintptr_t>(v)) & 15) == 0) ? (n >> 2) : 0;
  const float b1 = h.beta1, b2 = h.beta2, eps = h.eps;
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    mi = mi + (gi - mi) * (1.f - b1);              // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * b2 + (1.f - b2) * gi * gi;           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    pi -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  };
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], mm = m4[i], vv = v4[i];
    const float4 gg = adam_ld4(g, i);
    upd(pp.x, gg.x, mm.x, vv.x);
    upd(pp.y, gg.y, mm.y, vv.y);
    upd(pp.z, gg.z, mm.z, vv.z);
    upd(pp.w, gg.w, mm.w, vv.w);
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i], mi = m[i], vi = v[i];
    upd(pi, adam_ld1(g, i), mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

__global__ void adam_tick_kernel(artic_adam_hyper_t* hyper) { hyper->step += 1; }

}  // namespace artic

using namespace artic;

// One tile per block up to a generous cap: short-lived blocks let a low-priority weight prep yield SMs to the
// critical-path kernels it runs under (persistent blocks would hold them for the whole pass).
static unsigned wperm_grid(int64_t total_tiles) {
  const int64_t cap = 1LL << 20;
  return (unsigned)(total_tiles < cap ? (total_tiles > 0 ? total_tiles : 1) : cap);
}

extern "C" int64_t artic_wperm_tiles(int32_t K, int32_t G, int32_t A, int32_t B) {
  const int kc = K < PK ? K : PK;
  return (int64_t)((A + PT - 1) / PT) * ((B + PT - 1) / PT) * ((K + kc - 1) / kc) * G;
}

/* see include/artic.h */
extern "C" int64_t artic_wrow_tiles(int32_t K, int32_t G, int32_t A, int32_t B, int64_t sk, int64_t sa, int64_t sb,
                                    int32_t row_chunk) {
  if (K < 1 || G < 1 || A < 1 || B < 1 || K > RT_RUN) return 0;
  const bool a_inner = sa < sb || (sa == sb && A == 1);
  if ((a_inner ? sa : sb) != K || !(sk == 1 || K == 1)) return 0;      // taps must be the torch-inner index
  const int O = a_inner ? B : A, I = a_inner ? A : B, TI = rt_inner(K);
  const int n_it = (I + TI - 1) / TI;
  const int ch = row_chunk < 1 ? 1 : (row_chunk > RT_CHUNK_MAX ? RT_CHUNK_MAX : row_chunk);
  return (int64_t)G * ((O + RT_O - 1) / RT_O) * ((n_it + ch - 1) / ch);    // work units of <= row_chunk inner tiles
}

extern "C" int artic_weights_prep(const artic_wdesc_t* descs, int32_t n, int32_t any_norm, int64_t total_tiles,
                                  int64_t total_tiles2, void* stream) {
  ARTIC_CHECK_ARG(descs != nullptr && n >= 0 && total_tiles >= 0 && total_tiles2 >= 0, "bad descriptor table");
  if (n == 0) return ARTIC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (tc::g_debug[12] == 1) {          // debug: the generic three-pass version
    if (any_norm) wn_scale_kernel<<<dim3(128, (unsigned)n), 256, 0, st>>>(descs);
    wperm_kernel<0><<<wperm_grid(total_tiles), 256, 0, st>>>(descs, n, total_tiles);
    wperm_kernel<1><<<wperm_grid(total_tiles), 256, 0, st>>>(descs, n, total_tiles);
  } else {
    if (any_norm) {
      if (tc::g_debug[27] == 1) wn_scale4_kernel<<<dim3(128, (unsigned)n), 256, 0, st>>>(descs);
      else wn_scale_rows_kernel<<<dim3(128, (unsigned)n), 256, 0, st>>>(descs);
    }
    if (total_tiles > 0) wprep_kernel<<<wperm_grid(total_tiles), 256, 0, st>>>(descs, n, total_tiles);
    if (total_tiles2 > 0) {
      // debug key 31: cap on the grid (the kernel strides): a prep that runs BESIDE tensor-core launches (D's weights
      // under the generator forward) then holds only that many SMs — its 45 KB blocks cannot share an SM with a
      // 200+ KB tensor-core CTA, and an uncapped grid starves those CTAs until it has drained
      unsigned grid = wperm_grid(total_tiles2);
      if (tc::g_debug[31] > 0 && grid > (unsigned)tc::g_debug[31]) grid = (unsigned)tc::g_debug[31];
      wprep_rows_kernel<<<grid, 256, 0, st>>>(descs, n, total_tiles2);
    }
  }
  tc::note_weights_written(st);   // the next tensor-core conv on `st` must not prefetch weights early
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_weights_unprep(const artic_wdesc_t* descs, int32_t n, int32_t any_norm, int64_t total_tiles,
                                    int64_t total_tiles2, void* stream) {
  ARTIC_CHECK_ARG(descs != nullptr && n >= 0 && total_tiles >= 0 && total_tiles2 >= 0, "bad descriptor table");
  if (n == 0) return ARTIC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (tc::g_debug[12] == 1) {
    wperm_kernel<2><<<wperm_grid(total_tiles), 256, 0, st>>>(descs, n, total_tiles);
    if (any_norm) wn_bwd_kernel<<<dim3(128, (unsigned)n), 256, 0, st>>>(descs);
  } else {
    if (total_tiles > 0) wunprep_kernel<<<wperm_grid(total_tiles), 256, 0, st>>>(descs, n, total_tiles);
    if (total_tiles2 > 0) wunprep_rows_kernel<<<wperm_grid(total_tiles2), 256, 0, st>>>(descs, n, total_tiles2);
    if (any_norm) {
      if (tc::g_debug[27] == 1) wn_bwd4_kernel<<<dim3(128, (unsigned)n), 256, 0, st>>>(descs);
      else wn_bwd_rows_kernel<<<dim3(128, (unsigned)n), 256, 0, st>>>(descs);
    }
  }
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_adam_step(float* p, const float* g, float* m, float* v, int64_t n,
                               const artic_adam_hyper_t* hyper, void* stream) {
  ARTIC_CHECK_ARG(p && g && m && v && hyper, "null pointer");
  if (n == 0) return ARTIC_OK;
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  adam_kernel<float><<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, hyper);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_adam_step_wire(float* p, const void* g, int32_t g_dtype, float* m, float* v, int64_t n,
                                    const artic_adam_hyper_t* hyper, void* stream) {
  ARTIC_CHECK_ARG(p && g && m && v && hyper, "null pointer");
  ARTIC_CHECK_ARG(g_dtype == ARTIC_F32 || g_dtype == ARTIC_BF16, "bad gradient dtype");
  if (g_dtype == ARTIC_F32) return artic_adam_step(p, reinterpret_cast<const float*>(g), m, v, n, hyper, stream);
  if (n == 0) return ARTIC_OK;
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 16LL * num_sms()) blocks = 16LL * num_sms();
  adam_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, reinterpret_cast<const __nv_bfloat16*>(g), m, v, n, hyper);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_adam_tick(artic_adam_hyper_t* hyper, void* stream) {
  ARTIC_CHECK_ARG(hyper, "null pointer");
  adam_tick_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(hyper);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}
