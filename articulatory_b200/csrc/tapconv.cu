// Tap-gather GEMM on CUDA cores (fp32 accumulate) — the shape-generic implementation of
// artic_tapconv / artic_tapconv_wgrad / artic_colsum (see include/artic.h).
//
// It carries every conv-like contraction of the path in the fp32 ("accurate") mode and
// the odd shapes (C_in = 1, C_out = 1, tiny groups) in the bf16 mode; the dense
// stride-1 bf16 contractions are taken by the tcgen05 kernel in tapconv_tc.cu.
//
// Tiling: the M dimension is the FLATTENED (sequence, q) index so that short sequences
// (discriminator tails, L <= 53) still fill 128-row tiles; the K dimension is the flattened
// (tap, in-channel) index so that C_in/groups < 16 does not waste the K chunk.
#include <stdarg.h>

#include <stdlib.h>

#include "common.cuh"

namespace artic {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

constexpr int TM = 128;  // rows (flattened n,q) per CTA
constexpr int TK = 16;   // flattened (tap, ci) per chunk
constexpr int NT = 256;  // threads

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

// ------------------------------------------------------------------------------------
// forward / dgrad
// ------------------------------------------------------------------------------------
// TNT = output channels per thread (TN = 16 * TNT); VEC = 4-wide loads/stores allowed.
template <typename T, typename TO, int TNT, bool VEC>
__global__ void __launch_bounds__(NT) tapconv_kernel(const __grid_constant__ artic_tapconv_t p) {
  constexpr int TN = 16 * TNT;
  __shared__ __align__(16) float Xs[TK][TM];
  __shared__ __align__(16) float Ws[TK][TN];
  __shared__ int64_t row_xbase[TM];  // element offset of (n, row 0) in X; -1 = row inactive
  __shared__ int row_xpos[TM];       // q * si
  __shared__ int64_t row_ybase[TM];  // element offset of (n, out row) in Y; -1 = not stored

  const int tid = threadIdx.x;
  const int g = blockIdx.y / ((p.Cog + TN - 1) / TN);
  const int co0 = (blockIdx.y % ((p.Cog + TN - 1) / TN)) * TN;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int64_t Mtot = (int64_t)p.N * p.nq;

  if (tid < TM) {
    const int64_t m = m0 + tid;
    int64_t xb = -1, yb = -1;
    int xp = 0;
    if (m < Mtot) {
      const int n = (int)(m / p.nq);
      const int q = p.q0 + (int)(m % p.nq);
      xb = seq_base(p.x, n);
      xp = q * p.si;
      const int row = q * p.so + p.ro;
      if (row >= 0 && row < p.y.len) yb = seq_base(p.y, n) + (int64_t)row * p.y.s_row;
    }
    row_xbase[tid] = xb;
    row_xpos[tid] = xp;
    row_ybase[tid] = yb;
  }
  __syncthreads();

  const T* __restrict__ X = reinterpret_cast<const T*>(p.X);
  const T* __restrict__ W = reinterpret_cast<const T*>(p.W);

  float acc[8][TNT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TNT; ++j) acc[i][j] = 0.f;

  const int ty = tid >> 4;  // rows ty*8 .. +8
  const int tx = tid & 15;  // cols tx*TNT .. +TNT
  const int Ktot = p.ntaps * p.Cig;

  // X loader mapping: row = tid % 128, k-quad = tid / 128 + 2*i
  const int lr = tid & (TM - 1);
  const int64_t lxb = row_xbase[lr];
  const int lxp = row_xpos[lr];

  for (int k0 = 0; k0 < Ktot; k0 += TK) {
    // ---- stage X chunk: Xs[kk][row]
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int kq = (tid >> 7) + 2 * i;  // 0..3
      const int kidx = k0 + kq * 4;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (VEC) {
        if (kidx < Ktot && lxb >= 0) {
          const int tap = kidx / p.Cig;
          const int ci = kidx - tap * p.Cig;
          const int pos = lxp + p.off[tap];
          if (pos >= 0 && pos < p.x.len)
            Vec4<T>::load(X + lxb + (int64_t)pos * p.x.s_row + (int64_t)g * p.Cig + ci, v);
        }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ku = kidx + u;
          if (ku < Ktot && lxb >= 0) {
            const int tap = ku / p.Cig;
            const int ci = ku - tap * p.Cig;
            const int pos = lxp + p.off[tap];
            if (pos >= 0 && pos < p.x.len)
              v[u] = ld_f(X + lxb + (int64_t)pos * p.x.s_row + (int64_t)g * p.Cig + ci);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) Xs[kq * 4 + u][lr] = v[u];
    }
    // ---- stage W chunk: Ws[kk][c]
    for (int e = tid; e < TK * TN; e += NT) {
      const int kk = e / TN, c = e % TN;
      const int kidx = k0 + kk;
      float w = 0.f;
      if (kidx < Ktot && co0 + c < p.Cog) {
        const int tap = kidx / p.Cig;
        const int ci = kidx - tap * p.Cig;
        w = ld_f(W + (((int64_t)p.widx[tap] * p.G + g) * p.Cig + ci) * p.Cog + co0 + c);
      }
      Ws[kk][c] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[8], b[TNT];
      const float4 a0 = *reinterpret_cast<const float4*>(&Xs[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Xs[kk][ty * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      if (TNT == 4) {
        const float4 b0 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
        b[0] = b0.x; b[1 % TNT] = b0.y; b[2 % TNT] = b0.z; b[3 % TNT] = b0.w;
      } else {
#pragma unroll
        for (int j = 0; j < TNT; ++j) b[j] = Ws[kk][tx * TNT + j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TNT; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue
  const float* __restrict__ bias = p.bias;
  const TO* __restrict__ res_pre = reinterpret_cast<const TO*>(p.res_pre);
  const TO* __restrict__ mask = reinterpret_cast<const TO*>(p.mask);
  const TO* __restrict__ res = reinterpret_cast<const TO*>(p.res);
  const TO* __restrict__ res2 = reinterpret_cast<const TO*>(p.res2);
  TO* __restrict__ Y = reinterpret_cast<TO*>(p.Y);
  TO* __restrict__ Y2 = reinterpret_cast<TO*>(p.Y2);
  const int cbase = co0 + tx * TNT;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t yb = row_ybase[ty * 8 + i];
    if (yb < 0) continue;
    const int64_t o = yb + (int64_t)g * p.Cog + cbase;
    float v[TNT];
#pragma unroll
    for (int j = 0; j < TNT; ++j) {
      v[j] = p.alpha * acc[i][j];
      if (bias != nullptr && cbase + j < p.Cog) v[j] += __ldg(bias + g * p.Cog + cbase + j);
    }
    if (VEC && TNT == 4 && cbase + 3 < p.Cog) {
      float t[4];
      float vv[4] = {v[0], v[1 % TNT], v[2 % TNT], v[3 % TNT]};
      if (res_pre) { Vec4<TO>::load(res_pre + o, t); for (int j = 0; j < 4; ++j) vv[j] += t[j]; }
      if (mask) { Vec4<TO>::load(mask + o, t); for (int j = 0; j < 4; ++j) vv[j] *= (t[j] > 0.f ? 1.f : p.mask_slope); }
      if (res) { Vec4<TO>::load(res + o, t); for (int j = 0; j < 4; ++j) vv[j] += t[j]; }
      if (res2) { Vec4<TO>::load(res2 + o, t); for (int j = 0; j < 4; ++j) vv[j] += t[j]; }
      if (Y) Vec4<TO>::store(Y + o, vv);
      if (Y2) {
        for (int j = 0; j < 4; ++j) {
          if (p.act == ARTIC_ACT_LRELU) vv[j] = vv[j] > 0.f ? vv[j] : p.act_slope * vv[j];
          else if (p.act == ARTIC_ACT_TANH) vv[j] = tanhf(vv[j]);
        }
        Vec4<TO>::store(Y2 + o, vv);
      }
    } else {
#pragma unroll
      for (int j = 0; j < TNT; ++j) {
        if (cbase + j >= p.Cog) continue;
        float x = v[j];
        if (res_pre) x += ld_f(res_pre + o + j);
        if (mask) x *= (ld_f(mask + o + j) > 0.f ? 1.f : p.mask_slope);
        if (res) x += ld_f(res + o + j);
        if (res2) x += ld_f(res2 + o + j);
        if (Y) st_f(Y + o + j, x);
        if (Y2) {
          if (p.act == ARTIC_ACT_LRELU) x = x > 0.f ? x : p.act_slope * x;
          else if (p.act == ARTIC_ACT_TANH) x = tanhf(x);
          st_f(Y2 + o + j, x);
        }
      }
    }
  }
}

static bool seq_vec_ok(const artic_seq_t& s, int esize) {
  (void)esize;
  return (s.s_outer % 4 == 0) && (s.s_inner % 4 == 0) && (s.s_row % 4 == 0);
}
static bool ptr_vec_ok(const void* p, int esize) {
  return p == nullptr || (reinterpret_cast<uintptr_t>(p) % (4 * esize) == 0);
}

template <typename T, typename TO>
static int launch_tapconv(const artic_tapconv_t& p, cudaStream_t st) {
  const int es = (int)sizeof(T), eo = (int)sizeof(TO);
  const bool vec_in = (p.Cig % 4 == 0) && seq_vec_ok(p.x, es) && ptr_vec_ok(p.X, es);
  const bool vec_out = (p.Cog % 4 == 0) && seq_vec_ok(p.y, eo) && ptr_vec_ok(p.Y, eo) && ptr_vec_ok(p.Y2, eo) &&
                       ptr_vec_ok(p.res, eo) && ptr_vec_ok(p.res2, eo) && ptr_vec_ok(p.res_pre, eo) &&
                       ptr_vec_ok(p.mask, eo);
  const bool vec = vec_in && vec_out;
  const int64_t Mtot = (int64_t)p.N * p.nq;
  const int tnt = p.Cog > 16 ? 4 : 1;
  const int TN = 16 * tnt;
  dim3 grid((unsigned)((Mtot + TM - 1) / TM), (unsigned)(((p.Cog + TN - 1) / TN) * p.G), 1);
  if (grid.y > 65535) { set_error("artic_tapconv: too many channel tiles"); return ARTIC_ENOSUP; }
  if (tnt == 4) {
    if (vec) tapconv_kernel<T, TO, 4, true><<<grid, NT, 0, st>>>(p);
    else tapconv_kernel<T, TO, 4, false><<<grid, NT, 0, st>>>(p);
  } else {
    // narrow outputs: VEC epilogue is never used (TNT == 1); input vector loads still help
    if (vec_in) tapconv_kernel<T, TO, 1, true><<<grid, NT, 0, st>>>(p);
    else tapconv_kernel<T, TO, 1, false><<<grid, NT, 0, st>>>(p);
  }
  return ARTIC_OK;
}

// ------------------------------------------------------------------------------------
// wgrad
// ------------------------------------------------------------------------------------
constexpr int WT = 64;  // ci tile and co tile
constexpr int WR = 16;  // rows per chunk

template <typename T, typename TY>
__global__ void __launch_bounds__(NT) tapwgrad_kernel(const __grid_constant__ artic_tapwgrad_t p, int rows_per_split) {
  __shared__ __align__(16) float Xs[WR][WT];
  __shared__ __align__(16) float Ys[WR][WT];
  const int tid = threadIdx.x;
  const int n_ci_t = (p.Cig + WT - 1) / WT;
  const int n_co_t = (p.Cog + WT - 1) / WT;
  const int ci0 = (blockIdx.x % n_ci_t) * WT;
  const int co0 = (blockIdx.x / n_ci_t) * WT;
  (void)n_co_t;
  const int tap = blockIdx.y / p.G;
  const int g = blockIdx.y % p.G;
  const int64_t Mtot = (int64_t)p.N * p.nq;
  const int64_t mb = (int64_t)blockIdx.z * rows_per_split;
  const int64_t me = min(Mtot, mb + rows_per_split);
  const T* __restrict__ X = reinterpret_cast<const T*>(p.X);
  const TY* __restrict__ dY = reinterpret_cast<const TY*>(p.dY);
  const int xoff = p.off[tap], yoff = p.yoff[tap];

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int ty = tid >> 4, tx = tid & 15;
  const int lc = tid & 63;   // channel within tile
  const int lr0 = tid >> 6;  // 0..3

  for (int64_t mc = mb; mc < me; mc += WR) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = lr0 + 4 * i;
      const int64_t m = mc + r;
      float xv = 0.f, yv = 0.f;
      if (m < me) {
        const int n = (int)(m / p.nq);
        const int q = p.q0 + (int)(m % p.nq);
        const int xpos = q * p.si + xoff;
        const int ypos = q * p.so + yoff;
        if (xpos >= 0 && xpos < p.x.len && ypos >= 0 && ypos < p.y.len) {
          if (ci0 + lc < p.Cig)
            xv = ld_f(X + seq_base(p.x, n) + (int64_t)xpos * p.x.s_row + (int64_t)g * p.Cig + ci0 + lc);
          if (co0 + lc < p.Cog)
            yv = ld_f(dY + seq_base(p.y, n) + (int64_t)ypos * p.y.s_row + (int64_t)g * p.Cog + co0 + lc);
        }
      }
      Xs[r][lc] = xv;
      Ys[r][lc] = yv;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WR; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&Xs[r][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ys[r][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* __restrict__ dW = p.dW + (((int64_t)p.widx[tap] * p.G + g) * p.Cig) * p.Cog;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty * 4 + i;
    if (ci >= p.Cig) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < p.Cog) atomicAdd(dW + (int64_t)ci * p.Cog + co, acc[i][j]);
    }
  }
}

template <typename T, typename TY>
static int launch_tapwgrad(const artic_tapwgrad_t& p, cudaStream_t st) {
  const int n_ci_t = (p.Cig + WT - 1) / WT, n_co_t = (p.Cog + WT - 1) / WT;
  const int64_t Mtot = (int64_t)p.N * p.nq;
  const int64_t base_ctas = (int64_t)n_ci_t * n_co_t * p.ntaps * p.G;
  // split the row (reduction) dimension so that the grid has ~8 waves of CTAs
  int64_t want = (int64_t)num_sms() * 8;
  int64_t splits = (want + base_ctas - 1) / base_ctas;
  const int64_t max_splits = (Mtot + 4 * WR - 1) / (4 * WR);  // at least 64 rows per split
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  int64_t rps = (Mtot + splits - 1) / splits;
  rps = ((rps + WR - 1) / WR) * WR;
  splits = (Mtot + rps - 1) / rps;
  dim3 grid((unsigned)(n_ci_t * n_co_t), (unsigned)(p.ntaps * p.G), (unsigned)splits);
  if (grid.y > 65535) { set_error("artic_tapconv_wgrad: taps*groups too large"); return ARTIC_ENOSUP; }
  tapwgrad_kernel<T, TY><<<grid, NT, 0, st>>>(p, (int)rps);
  return ARTIC_OK;
}

// ------------------------------------------------------------------------------------
// column sum (bias gradient)
// ------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ dY, artic_seq_t y, int N, int C,
                                                     int rows_per_block, float* __restrict__ out) {
  __shared__ float part[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  const int64_t Mtot = (int64_t)N * y.len;
  const int64_t mb = (int64_t)blockIdx.y * rows_per_block;
  const int64_t me = min(Mtot, mb + rows_per_block);
  float s = 0.f;
  if (c < C) {
    for (int64_t m = mb + rl; m < me; m += 8) {
      const int n = (int)(m / y.len);
      const int row = (int)(m % y.len);
      s += ld_f(dY + seq_base(y, n) + (int64_t)row * y.s_row + c);
    }
  }
  part[rl][cl] = s;
  __syncthreads();
  if (rl == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][cl];
    atomicAdd(out + c, t);
  }
}

// Dense fast path: dY is a contiguous [M][C] bf16 matrix with C % 8 == 0 and (C/8) | 256.
// C/8 threads cover one row with 128-bit loads; the block's row slab is reduced in shared
// memory and flushed with one atomicAdd per channel.
__global__ void __launch_bounds__(256) colsum_dense_bf16_kernel(const __nv_bfloat16* __restrict__ dY, int64_t M, int C,
                                                                int rows_per_block, float* __restrict__ out) {
  __shared__ float part[256][9];
  const int tpr = C >> 3, rpp = 256 / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  const int64_t mb = (int64_t)blockIdx.x * rows_per_block;
  const int64_t me = min(M, mb + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const uint4* __restrict__ src = reinterpret_cast<const uint4*>(dY);
  int64_t m = mb + rl;
  for (; m + 3 * rpp < me; m += 4 * rpp) {
    uint4 u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) u[j] = __ldg(src + (m + (int64_t)j * rpp) * tpr + cg);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += __uint_as_float(w[i] << 16);
        acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
      }
    }
  }
  for (; m < me; m += rpp) {
    const uint4 u = __ldg(src + m * tpr + cg);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] += __uint_as_float(w[i] << 16);
      acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) part[threadIdx.x][i] = acc[i];
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    const int g8 = c >> 3, i = c & 7;
    float t = 0.f;
    for (int r = 0; r < rpp; ++r) t += part[r * tpr + g8][i];
    atomicAdd(out + c, t);
  }
}

}  // namespace artic

using namespace artic;

extern "C" const char* artic_last_error(void) { return artic::g_err; }

namespace artic { long long g_path_counts[12] = {0}; }
extern "C" int artic_path_counts(int64_t* h_out, int32_t reset) {
  ARTIC_CHECK_ARG(h_out != nullptr, "null pointer");
  for (int i = 0; i < 12; ++i) h_out[i] = artic::g_path_counts[i];
  if (reset) for (int i = 0; i < 12; ++i) artic::g_path_counts[i] = 0;
  return ARTIC_OK;
}
extern "C" int artic_version(void) { return 100; }
extern "C" const char* artic_arch(void) { return "sm_100a"; }

static int check_seq(const artic_seq_t& s) { return s.n_inner >= 1 && s.len >= 0; }

// implemented in tapconv_tc.cu: returns 1 if it took the launch, 0 if the shape is not
// eligible (fall through to the generic kernel), <0 on error.
int artic_tapconv_tc_multi(const artic_tapconv_t* ps, int n, int* taken, cudaStream_t st);   // tapconv_tc.cu
int artic_tapwgrad_tc_try(const artic_tapwgrad_t* p, cudaStream_t st, int* bias_fused);  // tapwgrad_tc.cu, same convention
int artic_tapconv_co1_try(const artic_tapconv_t* p, cudaStream_t st);    // smallc.cu
int artic_tapconv_ci1_try(const artic_tapconv_t* p, cudaStream_t st);    // smallc.cu
int artic_tapwgrad_ci1_try(const artic_tapwgrad_t* p, cudaStream_t st);  // smallc.cu

static int check_tapconv(const artic_tapconv_t* p) {
  ARTIC_CHECK_ARG(p != nullptr, "null params");
  ARTIC_CHECK_ARG(p->X && p->W && (p->Y || p->Y2), "X, W and one of Y/Y2 are required");
  ARTIC_CHECK_ARG(check_seq(p->x) && check_seq(p->y), "bad sequence descriptor");
  ARTIC_CHECK_ARG(p->N >= 0 && p->G >= 1 && p->Cig >= 1 && p->Cog >= 1, "bad dims");
  ARTIC_CHECK_ARG(p->ntaps >= 1 && p->ntaps <= ARTIC_MAX_TAPS, "ntaps out of range");
  ARTIC_CHECK_ARG(p->si >= 1 && p->so >= 1 && p->nq >= 0, "bad q mapping");
  ARTIC_CHECK_ARG(p->dtype == ARTIC_F32 || p->dtype == ARTIC_BF16, "bad dtype");
  ARTIC_CHECK_ARG(p->out_dtype == ARTIC_F32 || p->out_dtype == ARTIC_BF16, "bad out_dtype");
  return ARTIC_OK;
}

// one problem on the CUDA-core kernels (channel-1 special cases first)
static int tapconv_fallback(const artic_tapconv_t* p, cudaStream_t st) {
  if (p->N == 0 || p->nq == 0) return ARTIC_OK;
  if (artic_tapconv_co1_try(p, st) == 1 || artic_tapconv_ci1_try(p, st) == 1) {
    ARTIC_LAUNCH_CHECK();
    ++g_path_counts[PATH_CONV_C1];
    return ARTIC_OK;
  }
  ++g_path_counts[PATH_CONV_GENERIC];
  {   // ARTIC_LOG_GENERIC=1: name the shapes that run on the generic CUDA-core kernel (stderr)
    static const bool log = getenv("ARTIC_LOG_GENERIC") != nullptr;
    if (log)
      fprintf(stderr, "[artic] generic conv: N=%d nq=%d G=%d Cig=%d Cog=%d taps=%d si=%d so=%d dtype=%d/%d x_sp=%d\n", p->N, p->nq,
              p->G, p->Cig, p->Cog, p->ntaps, p->si, p->so, p->dtype, p->out_dtype, p->X_sp != nullptr);
  }
  const bool ob = p->out_dtype == ARTIC_BF16;
  int rc;
  if (p->dtype == ARTIC_BF16) rc = ob ? launch_tapconv<__nv_bfloat16, __nv_bfloat16>(*p, st) : launch_tapconv<__nv_bfloat16, float>(*p, st);
  else rc = ob ? launch_tapconv<float, __nv_bfloat16>(*p, st) : launch_tapconv<float, float>(*p, st);
  if (rc != ARTIC_OK) return rc;
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_tapconv_multi(const artic_tapconv_t* ps, int32_t n, void* stream) {
  ARTIC_CHECK_ARG(ps != nullptr && n >= 0 && n <= 64, "bad problem list");
  for (int i = 0; i < n; ++i) {
    const int rc = check_tapconv(&ps[i]);
    if (rc != ARTIC_OK) return rc;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int taken[64];
  bool cand[64];
  artic_tapconv_t tcp[64];
  int map[64], m = 0;
  for (int i = 0; i < n; ++i) {     // channel-1 problems never go to the tensor cores
    cand[i] = !(ps[i].Cog == 1 && ps[i].G == 1 && ps[i].Cig >= 32) && ps[i].Cig != 1;
    if (cand[i]) { tcp[m] = ps[i]; map[m++] = i; }
  }
  int tk[64];
  int rc = artic_tapconv_tc_multi(tcp, m, tk, st);
  if (rc != ARTIC_OK) return rc;
  for (int i = 0; i < n; ++i) taken[i] = 0;
  for (int k = 0; k < m; ++k) taken[map[k]] = tk[k];
  for (int i = 0; i < n; ++i) {
    if (taken[i]) continue;
    rc = tapconv_fallback(&ps[i], st);
    if (rc != ARTIC_OK) return rc;
  }
  // bf16x3 mode: split copies requested from a problem that ran on a CUDA-core kernel (the tcgen05 epilogue
  // writes them itself).  The phases of one layer share their output tensor: split it once, after the last.
  for (int i = 0; i < n; ++i) {
    if (taken[i] || ps[i].N == 0 || ps[i].nq == 0 || ps[i].out_dtype != ARTIC_F32) continue;
    const void* outs[2] = {ps[i].Y, ps[i].Y2};
    void* sps[2] = {ps[i].Y_sp, ps[i].Y2_sp};
    for (int k = 0; k < 2; ++k) {
      if (outs[k] == nullptr || sps[k] == nullptr) continue;
      bool later = false;
      for (int j = i + 1; j < n && !later; ++j)
        later = !taken[j] && ps[j].N != 0 && ps[j].nq != 0 && (ps[j].Y == outs[k] || ps[j].Y2 == outs[k]);
      if (later) continue;
      const artic_seq_t& y = ps[i].y;
      const int64_t total = (int64_t)((ps[i].N + y.n_inner - 1) / y.n_inner) * y.s_outer;
      rc = artic_split(reinterpret_cast<const float*>(outs[k]), sps[k], ps[i].y_plane, total, stream);
      if (rc != ARTIC_OK) return rc;
    }
  }
  return ARTIC_OK;
}

extern "C" int artic_tapconv(const artic_tapconv_t* p, void* stream) { return artic_tapconv_multi(p, 1, stream); }

extern "C" int artic_tapconv_wgrad(const artic_tapwgrad_t* p, void* stream) {
  ARTIC_CHECK_ARG(p != nullptr, "null params");
  ARTIC_CHECK_ARG(p->X && p->dY && p->dW, "X, dY, dW are required");
  ARTIC_CHECK_ARG(check_seq(p->x) && check_seq(p->y), "bad sequence descriptor");
  ARTIC_CHECK_ARG(p->N >= 0 && p->G >= 1 && p->Cig >= 1 && p->Cog >= 1, "bad dims");
  ARTIC_CHECK_ARG(p->ntaps >= 1 && p->ntaps <= ARTIC_MAX_TAPS, "ntaps out of range");
  ARTIC_CHECK_ARG(p->si >= 1 && p->so >= 1 && p->nq >= 0, "bad q mapping");
  ARTIC_CHECK_ARG(p->dtype == ARTIC_F32 || p->dtype == ARTIC_BF16, "bad dtype");
  ARTIC_CHECK_ARG(p->y_dtype == ARTIC_F32 || p->y_dtype == ARTIC_BF16, "bad y_dtype");
  if (p->N == 0 || p->nq == 0) return ARTIC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const bool yb = p->y_dtype == ARTIC_BF16;
  if (p->dbias != nullptr)
    ARTIC_CHECK_ARG(p->q0 + p->yoff[0] == 0 && p->nq == p->y.len && p->so == 1, "dbias needs a launch over every dY row");
  if (artic_tapwgrad_ci1_try(p, st) == 1) {
    ARTIC_LAUNCH_CHECK();
    ++g_path_counts[PATH_WGRAD_C1];
    if (p->dbias != nullptr) return artic_colsum(p->dY, &p->y, p->N, p->G * p->Cog, p->y_dtype, p->dbias, stream);
    return ARTIC_OK;
  }
  int bias_fused = 0;
  int rc = artic_tapwgrad_tc_try(p, st, &bias_fused);
  if (rc < 0) return rc;
  if (p->dbias != nullptr && !bias_fused) {
    const int brc = artic_colsum(p->dY, &p->y, p->N, p->G * p->Cog, p->y_dtype, p->dbias, stream);
    if (brc != ARTIC_OK) return brc;
  }
  if (rc == 1) return ARTIC_OK;
  ++g_path_counts[PATH_WGRAD_GENERIC];
  {
    static const bool log = getenv("ARTIC_LOG_GENERIC") != nullptr;
    if (log)
      fprintf(stderr, "[artic] generic wgrad: N=%d nq=%d G=%d Cig=%d Cog=%d taps=%d si=%d so=%d dtype=%d/%d\n", p->N, p->nq, p->G,
              p->Cig, p->Cog, p->ntaps, p->si, p->so, p->dtype, p->y_dtype);
  }
  if (p->dtype == ARTIC_BF16) rc = yb ? launch_tapwgrad<__nv_bfloat16, __nv_bfloat16>(*p, st) : launch_tapwgrad<__nv_bfloat16, float>(*p, st);
  else rc = yb ? launch_tapwgrad<float, __nv_bfloat16>(*p, st) : launch_tapwgrad<float, float>(*p, st);
  if (rc != ARTIC_OK) return rc;
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_colsum(const void* dY, const artic_seq_t* y, int32_t N, int32_t C, int32_t dtype,
                            float* out, void* stream) {
  ARTIC_CHECK_ARG(dY && y && out, "null pointer");
  ARTIC_CHECK_ARG(check_seq(*y) && N >= 0 && C >= 1, "bad dims");
  ARTIC_CHECK_ARG(dtype == ARTIC_F32 || dtype == ARTIC_BF16, "bad dtype");
  const int64_t Mtot = (int64_t)N * y->len;
  if (Mtot == 0) return ARTIC_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  {
    const bool dense = y->s_row == (int64_t)C * y->n_inner && (y->n_inner == 1 || y->s_inner == C) &&
                       y->s_outer == (int64_t)y->len * y->s_row && N % y->n_inner == 0;
    const int tpr = C / 8;
    if (dtype == ARTIC_BF16 && dense && C % 8 == 0 && tpr >= 1 && tpr <= 256 && 256 % tpr == 0 &&
        (reinterpret_cast<uintptr_t>(dY) & 15) == 0) {
      const int rpp = 256 / tpr;
      int64_t blocks = 4LL * num_sms();
      int64_t rpb = (Mtot + blocks - 1) / blocks;
      rpb = ((rpb + rpp - 1) / rpp) * rpp;
      if (rpb < 4 * rpp) rpb = 4 * rpp;
      blocks = (Mtot + rpb - 1) / rpb;
      colsum_dense_bf16_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dY), Mtot, C,
                                                                (int)rpb, out);
      ARTIC_LAUNCH_CHECK();
      return ARTIC_OK;
    }
  }
  const int ctiles = (C + 31) / 32;
  int64_t blocks_y = (4LL * num_sms() + ctiles - 1) / ctiles;
  int64_t rpb = (Mtot + blocks_y - 1) / blocks_y;
  if (rpb < 64) rpb = 64;
  blocks_y = (Mtot + rpb - 1) / rpb;
  dim3 grid((unsigned)ctiles, (unsigned)blocks_y);
  if (dtype == ARTIC_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dY), *y, N, C, (int)rpb, out);
  else
    colsum_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(dY), *y, N, C, (int)rpb, out);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}
