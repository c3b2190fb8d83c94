// Channel-1 ends of the networks, where the tap-gather contraction degenerates and the GEMM
// tilings waste the machine (these launches are HBM / latency bound):
//
//   * Cog == 1  (logit convs 1024->1, the generator's output conv 32->1 + tanh, and the
//     data gradient of every first layer): one warp per output row, lanes across the
//     (tap, channel) reduction with 128-bit loads, warp-shuffle reduction, fused epilogue.
//   * Cig == 1 weight gradient (first layers on the raw fp32 signal, k = 15 / 5):
//     dW[t][co] = sum_{n,q} x[n, q*si + off_t] * dY[n, q, co]; threads own output channels
//     (coalesced dY rows), the taps live in registers, one atomicAdd per (tap, co) per CTA.
#include "common.cuh"
#include "tc_common.cuh"

namespace artic {

// ------------------------------------------------------------------------------------
// Cog == 1 forward / dgrad
// ------------------------------------------------------------------------------------
template <typename T> struct Ld8;
template <> struct Ld8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <> struct Ld8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&f)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

// One warp per output row.  VEC: Cig % 8 == 0 and 16/32-byte aligned rows -> 8-wide loads.
template <typename T, typename TO, bool VEC>
__global__ void __launch_bounds__(256) tapconv_co1_kernel(const __grid_constant__ artic_tapconv_t p) {
  const int lane = threadIdx.x & 31;
  const int64_t Mtot = (int64_t)p.N * p.nq;
  const int64_t m = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= Mtot) return;
  const int n = (int)(m / p.nq);
  const int q = p.q0 + (int)(m % p.nq);
  const int row = q * p.so + p.ro;
  if (row < 0 || row >= p.y.len) return;
  const T* __restrict__ X = reinterpret_cast<const T*>(p.X) + seq_base(p.x, n);
  const T* __restrict__ W = reinterpret_cast<const T*>(p.W);
  float acc = 0.f;
  if (VEC) {
    // lanes run over the flattened (tap, 8-channel chunk) index, so narrow layers (the generator's
    // 32 -> 1 output conv: 7 taps x 4 chunks) still use 28 of 32 lanes
    const int cpt = p.Cig >> 3;
    for (int e = lane; e < p.ntaps * cpt; e += 32) {
      const int t = e / cpt, c = (e - t * cpt) << 3;
      const int pos = q * p.si + p.off[t];
      if (pos < 0 || pos >= p.x.len) continue;
      float a[8], b[8];
      Ld8<T>::load(X + (int64_t)pos * p.x.s_row + c, a);
      Ld8<T>::load(W + (int64_t)p.widx[t] * p.Cig + c, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(a[i], b[i], acc);
    }
  } else {
    for (int t = 0; t < p.ntaps; ++t) {
      const int pos = q * p.si + p.off[t];
      if (pos < 0 || pos >= p.x.len) continue;
      const T* xr = X + (int64_t)pos * p.x.s_row;
      const T* wr = W + (int64_t)p.widx[t] * p.Cig;   // [K][G=1][Cig][Cog=1]
      for (int c = lane; c < p.Cig; c += 32) acc = fmaf(ld_f(xr + c), ld_f(wr + c), acc);
    }
  }
  acc = warp_sum(acc);
  if (lane != 0) return;
  const int64_t o = seq_base(p.y, n) + (int64_t)row * p.y.s_row;
  float v = p.alpha * acc;
  if (p.bias) v += __ldg(p.bias);
  if (p.res_pre) v += ld_f(reinterpret_cast<const TO*>(p.res_pre) + o);
  if (p.mask) v *= (ld_f(reinterpret_cast<const TO*>(p.mask) + o) > 0.f ? 1.f : p.mask_slope);
  if (p.res) v += ld_f(reinterpret_cast<const TO*>(p.res) + o);
  if (p.res2) v += ld_f(reinterpret_cast<const TO*>(p.res2) + o);
  if (p.Y) st_f(reinterpret_cast<TO*>(p.Y) + o, v);
  if (p.Y2) {
    if (p.act == ARTIC_ACT_LRELU) v = v > 0.f ? v : p.act_slope * v;
    else if (p.act == ARTIC_ACT_TANH) v = tanhf(v);
    st_f(reinterpret_cast<TO*>(p.Y2) + o, v);
  }
}

// Staged variant for bf16 inputs of moderate width (Cig <= 256): a CTA produces 256 consecutive output rows,
// ONE THREAD PER ROW, from one copy of its input window in shared memory (row pitch Cig * 2 + 16 bytes: the
// 16-byte reads of 8 neighbouring rows cover all 32 banks) and a broadcast fp32 copy of the weights.  The plain
// kernel re-reads every input row once per tap through L1/L2 and pays a shuffle reduction plus a serial
// epilogue per row (136 us for the k = 15 first-layer data gradient of a 16 x 8512 x 128 tensor).
constexpr int CO1_ROWS = 256;
template <typename TO>
__global__ void __launch_bounds__(256) tapconv_co1_staged_kernel(const __grid_constant__ artic_tapconv_t p, int min_off, int span) {
  extern __shared__ uint4 xs4[];
  using T = __nv_bfloat16;
  const int n = blockIdx.y;
  const int qa = blockIdx.x * CO1_ROWS, qb = min(p.nq, qa + CO1_ROWS);
  const int cpt = p.Cig >> 3, pitch = cpt + 1;
  const T* __restrict__ X = reinterpret_cast<const T*>(p.X) + seq_base(p.x, n);
  const T* __restrict__ W = reinterpret_cast<const T*>(p.W);
  const int x0 = (p.q0 + qa) * p.si + min_off;
  const int nrows = (qb - qa - 1) * p.si + span + 1;
  float4* wsm = reinterpret_cast<float4*>(xs4 + (size_t)((CO1_ROWS - 1) * p.si + span + 1) * pitch);   // [tap][cpt][2] float4
  for (int i = threadIdx.x; i < nrows * cpt; i += 256) {
    const int r = i / cpt, c = i - r * cpt;
    const int pos = x0 + r;
    xs4[r * pitch + c] = (pos >= 0 && pos < p.x.len) ? __ldg(reinterpret_cast<const uint4*>(X + (int64_t)pos * p.x.s_row + 8 * c))
                                                     : make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = threadIdx.x; i < p.ntaps * cpt; i += 256) {
    const int t = i / cpt, c = i - t * cpt;
    float f[8];
    Ld8<T>::load(W + (int64_t)p.widx[t] * p.Cig + 8 * c, f);
    wsm[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
    wsm[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
  __syncthreads();
  const int q = qa + threadIdx.x;
  if (q >= qb) return;
  const int row = (p.q0 + q) * p.so + p.ro;
  if (row < 0 || row >= p.y.len) return;
  const uint4* xr = xs4 + (size_t)(threadIdx.x * p.si) * pitch;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int t = 0; t < p.ntaps; ++t) {
    const uint4* xt = xr + (p.off[t] - min_off) * pitch;
    const float4* wt = wsm + 2 * t * cpt;
#pragma unroll 4
    for (int c = 0; c < cpt; ++c) {
      const uint4 u = xt[c];
      const float4 w0 = wt[2 * c], w1 = wt[2 * c + 1];
      acc[0] = fmaf(__uint_as_float(u.x << 16), w0.x, acc[0]);
      acc[1] = fmaf(__uint_as_float(u.x & 0xffff0000u), w0.y, acc[1]);
      acc[2] = fmaf(__uint_as_float(u.y << 16), w0.z, acc[2]);
      acc[3] = fmaf(__uint_as_float(u.y & 0xffff0000u), w0.w, acc[3]);
      acc[0] = fmaf(__uint_as_float(u.z << 16), w1.x, acc[0]);
      acc[1] = fmaf(__uint_as_float(u.z & 0xffff0000u), w1.y, acc[1]);
      acc[2] = fmaf(__uint_as_float(u.w << 16), w1.z, acc[2]);
      acc[3] = fmaf(__uint_as_float(u.w & 0xffff0000u), w1.w, acc[3]);
    }
  }
  const int64_t o = seq_base(p.y, n) + (int64_t)row * p.y.s_row;
  float v = p.alpha * ((acc[0] + acc[1]) + (acc[2] + acc[3]));
  if (p.bias) v += __ldg(p.bias);
  if (p.res_pre) v += ld_f(reinterpret_cast<const TO*>(p.res_pre) + o);
  if (p.mask) v *= (ld_f(reinterpret_cast<const TO*>(p.mask) + o) > 0.f ? 1.f : p.mask_slope);
  if (p.res) v += ld_f(reinterpret_cast<const TO*>(p.res) + o);
  if (p.res2) v += ld_f(reinterpret_cast<const TO*>(p.res2) + o);
  if (p.Y) st_f(reinterpret_cast<TO*>(p.Y) + o, v);
  if (p.Y2) {
    if (p.act == ARTIC_ACT_LRELU) v = v > 0.f ? v : p.act_slope * v;
    else if (p.act == ARTIC_ACT_TANH) v = tanhf(v);
    st_f(reinterpret_cast<TO*>(p.Y2) + o, v);
  }
}

template <typename T, typename TO>
static void launch_co1(const artic_tapconv_t& p, cudaStream_t st) {
  const int64_t Mtot = (int64_t)p.N * p.nq;
  const unsigned grid = (unsigned)((Mtot + 7) / 8);
  const int es = (int)sizeof(T);
  const bool vec = (p.Cig % 8 == 0) && (p.x.s_row % 8 == 0) && (p.x.s_outer % 8 == 0) && (p.x.s_inner % 8 == 0) &&
                   (reinterpret_cast<uintptr_t>(p.X) % (8 * es) == 0) && (reinterpret_cast<uintptr_t>(p.W) % (8 * es) == 0);
  if (vec && sizeof(T) == 2 && p.Cig <= 256 && p.N <= 65535 && p.nq >= 64) {
    int min_off = p.off[0], max_off = p.off[0];
    for (int t = 1; t < p.ntaps; ++t) { min_off = min(min_off, p.off[t]); max_off = max(max_off, p.off[t]); }
    const int span = max_off - min_off;
    const size_t smem = (size_t)((CO1_ROWS - 1) * p.si + span + 1) * ((p.Cig >> 3) + 1) * 16 + (size_t)p.ntaps * p.Cig * 4;
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(tapconv_co1_staged_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
      attr_set = true;
    }
    if (min_off > -(1 << 20) && smem <= 160 * 1024) {
      dim3 g2((unsigned)((p.nq + CO1_ROWS - 1) / CO1_ROWS), (unsigned)p.N);
      tapconv_co1_staged_kernel<TO><<<g2, 256, smem, st>>>(p, min_off, span);
      return;
    }
  }
  if (vec) tapconv_co1_kernel<T, TO, true><<<grid, 256, 0, st>>>(p);
  else tapconv_co1_kernel<T, TO, false><<<grid, 256, 0, st>>>(p);
}

// ------------------------------------------------------------------------------------
// Cig == 1 forward (first layer of every discriminator chain: raw fp32 signal -> C channels)
// ------------------------------------------------------------------------------------
// HBM-bound on the output write: the CTA stages its input window and the whole [taps][C] weight in
// shared memory; Cog/8 threads produce one output row with a single 128-bit store each.
constexpr int CI1_ROWS = 512;     // output positions per CTA
constexpr int CI1_XS = 2176;      // staged input samples
constexpr int CI1_WS = 4096;      // staged weights (taps * Cog)

// NT = register-resident taps: each thread keeps the weights of its 8 output channels for all (<= NT) taps
// in registers for the CTA's lifetime, so a row costs ntaps broadcast loads of x and 8 * ntaps FMAs (the
// first version re-read two 16-byte weight vectors per tap and row and was bound by shared-memory bandwidth:
// 150 us for the 1 -> 128, k = 15 layer on a 32 x 8512 batch).  NT = 0: weights stay in shared memory.
template <typename T, int NT>
__global__ void __launch_bounds__(256) tapconv_ci1_kernel(const __grid_constant__ artic_tapconv_t p, int min_off, int span) {
  __shared__ float xs[CI1_XS];
  __shared__ __align__(16) float wsm[CI1_WS];
  __shared__ __align__(16) float bsm[2048];       // Cog <= 8 * 256
  const int tpr = p.Cog >> 3, rpp = 256 / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  const int n = blockIdx.y;
  const int qa = blockIdx.x * CI1_ROWS;
  const int qb = min(p.nq, qa + CI1_ROWS);
  const T* __restrict__ X = reinterpret_cast<const T*>(p.X) + seq_base(p.x, n);
  const float* __restrict__ W = reinterpret_cast<const float*>(p.W);      // [K][1][1][Cog] fp32
  const int x0 = (p.q0 + qa) * p.si + min_off;
  const int nx = (qb - qa - 1) * p.si + span + 1;
  for (int i = threadIdx.x; i < nx; i += 256) {
    const int pos = x0 + i;
    xs[i] = (pos >= 0 && pos < p.x.len) ? ld_f(X + (int64_t)pos * p.x.s_row) : 0.f;
  }
  for (int i = threadIdx.x; i < p.ntaps * p.Cog; i += 256)
    wsm[i] = __ldg(W + (int64_t)p.widx[i / p.Cog] * p.Cog + (i % p.Cog));
  for (int i = threadIdx.x; i < p.Cog; i += 256) bsm[i] = p.bias != nullptr ? __ldg(p.bias + i) : 0.f;
  __syncthreads();
  float wr[NT > 0 ? NT : 1][8];
  int offr[NT > 0 ? NT : 1];
  if (NT > 0) {
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const bool on = t < p.ntaps;
      offr[t] = on ? p.off[t] - min_off : 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) wr[t][i] = on ? wsm[t * p.Cog + cg * 8 + i] : 0.f;
    }
  }
  float bias8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bias8[i] = bsm[cg * 8 + i];
  __nv_bfloat16* __restrict__ Y = reinterpret_cast<__nv_bfloat16*>(p.Y);
  __nv_bfloat16* __restrict__ Y2 = reinterpret_cast<__nv_bfloat16*>(p.Y2);
  const int64_t ybase = seq_base(p.y, n);
  const __nv_bfloat16* __restrict__ res_pre = reinterpret_cast<const __nv_bfloat16*>(p.res_pre);
  const __nv_bfloat16* __restrict__ mask = reinterpret_cast<const __nv_bfloat16*>(p.mask);
  for (int q = qa + rl; q < qb; q += rpp) {
    const int row = p.q0 + q + p.ro;                // so == 1
    if (row < 0 || row >= p.y.len) continue;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (NT > 0) {
      const float* xp = xs + (q - qa) * p.si;
      float xv[NT > 0 ? NT : 1];
#pragma unroll
      for (int t = 0; t < NT; ++t) xv[t] = xp[offr[t]];      // all tap samples in flight first; padded taps: weight 0, offset 0
#pragma unroll
      for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fmaf(xv[t], wr[t][i], acc[i]);
    } else {
      const int xb = (q - qa) * p.si - min_off;
      for (int t = 0; t < p.ntaps; ++t) {
        const float xv = xs[xb + p.off[t]];
        const float4 w0 = *reinterpret_cast<const float4*>(&wsm[t * p.Cog + cg * 8]);
        const float4 w1 = *reinterpret_cast<const float4*>(&wsm[t * p.Cog + cg * 8 + 4]);
        acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
        acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
        acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
        acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
      }
    }
    uint32_t o1[4], o2[4];
    const int64_t o = ybase + (int64_t)row * p.y.s_row + cg * 8;
    // data gradient of a 1-channel logit conv: + upstream feature-matching gradient, x LeakyReLU' of the saved activation
    float rp[8], mk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { rp[i] = 0.f; mk[i] = 1.f; }
    if (res_pre != nullptr) tc::unpack8<__nv_bfloat16>(__ldg(reinterpret_cast<const uint4*>(res_pre + o)), rp);
    if (mask != nullptr) {
      float m[8];
      tc::unpack8<__nv_bfloat16>(__ldg(reinterpret_cast<const uint4*>(mask + o)), m);
#pragma unroll
      for (int i = 0; i < 8; ++i) mk[i] = m[i] > 0.f ? 1.f : p.mask_slope;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float a = (p.alpha * acc[2 * i] + bias8[2 * i] + rp[2 * i]) * mk[2 * i];
      float b = (p.alpha * acc[2 * i + 1] + bias8[2 * i + 1] + rp[2 * i + 1]) * mk[2 * i + 1];
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      o1[i] = *reinterpret_cast<uint32_t*>(&h);
      if (p.act == ARTIC_ACT_LRELU) { a = a > 0.f ? a : p.act_slope * a; b = b > 0.f ? b : p.act_slope * b; }
      else if (p.act == ARTIC_ACT_TANH) { a = tanhf(a); b = tanhf(b); }
      h = __floats2bfloat162_rn(a, b);
      o2[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    if (Y) *reinterpret_cast<uint4*>(Y + o) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    if (Y2) *reinterpret_cast<uint4*>(Y2 + o) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  }
}

// ------------------------------------------------------------------------------------
// Cig == 1 and Cog == 1 weight gradients
// ------------------------------------------------------------------------------------
constexpr int C1_TAPS = 16;   // taps per pass (registers)
constexpr int C1_ROWS = 512;  // positions per CTA
constexpr int C1_SMEM = 2304; // staged single-channel samples per CTA

// dW[t][co] = sum_q x[q*si + off_t] * dY[q*so + yoff_t][co]   (x single channel, staged in smem).
// grid = (chunks per sequence, N, channel tiles); threads = cw channel lanes x 256/cw position lanes.
template <typename T, typename TY>
__global__ void __launch_bounds__(256) tapwgrad_ci1_kernel(const __grid_constant__ artic_tapwgrad_t p, int cw, int tap0,
                                                           int min_off, int span) {
  __shared__ float xs[C1_SMEM];
  __shared__ float red[256];
  const int cl = threadIdx.x % cw, pl = threadIdx.x / cw, npl = 256 / cw;
  const int c = blockIdx.z * cw + cl;
  const int n = blockIdx.y;
  const int qa = blockIdx.x * C1_ROWS;                 // relative to q0
  const int qb = min(p.nq, qa + C1_ROWS);
  const T* __restrict__ X = reinterpret_cast<const T*>(p.X) + seq_base(p.x, n);
  const TY* __restrict__ dY = reinterpret_cast<const TY*>(p.dY) + seq_base(p.y, n);
  const int x0 = (p.q0 + qa) * p.si + min_off;         // first staged sample
  const int nx = (qb - qa - 1) * p.si + span + 1;
  for (int i = threadIdx.x; i < nx; i += 256) {
    const int pos = x0 + i;
    xs[i] = (pos >= 0 && pos < p.x.len) ? ld_f(X + (int64_t)pos * p.x.s_row) : 0.f;
  }
  __syncthreads();
  const int nt = min(C1_TAPS, p.ntaps - tap0);
  float acc[C1_TAPS];
#pragma unroll
  for (int t = 0; t < C1_TAPS; ++t) acc[t] = 0.f;
  const int yo = p.yoff[tap0];
  if (c < p.Cog) {
    for (int q = qa + pl; q < qb; q += npl) {
      const int ypos = (p.q0 + q) * p.so + yo;
      if (ypos < 0 || ypos >= p.y.len) continue;
      const float dy = ld_f(dY + (int64_t)ypos * p.y.s_row + c);
      const int xb = (q - qa) * p.si - min_off;
#pragma unroll
      for (int t = 0; t < C1_TAPS; ++t)
        if (t < nt) acc[t] = fmaf(xs[xb + p.off[tap0 + t]], dy, acc[t]);
    }
  }
  for (int t = 0; t < nt; ++t) {
    __syncthreads();
    red[threadIdx.x] = acc[t];
    __syncthreads();
    if (pl == 0 && c < p.Cog) {
      float s = 0.f;
      for (int i = 0; i < npl; ++i) s += red[i * cw + cl];
      atomicAdd(p.dW + (int64_t)p.widx[tap0 + t] * p.Cog + c, s);   // [K][G=1][Cig=1][Cog]
    }
  }
}

// Vector variant of the above for bf16 dY rows (Cog % 8 == 0): Cog/8 threads cover one dY row with
// 128-bit loads, each thread keeps [taps][8 channels] partial sums in registers, the position loop
// is unrolled 4x so that four independent row loads are in flight (the scalar kernel above is
// latency-bound: one 2-byte load per thread per iteration).
constexpr int C1V_ROWS = 256;   // positions per CTA of the vector kernel (twice the CTAs of the scalar ones)
template <typename T, typename TY = __nv_bfloat16>
__global__ void __launch_bounds__(256) tapwgrad_ci1_vec_kernel(const __grid_constant__ artic_tapwgrad_t p, int tap0,
                                                               int min_off, int span, int dbg) {
  constexpr bool YF = sizeof(TY) == 4;          // fp32 dY (bf16x3 / fp32 modes): two 128-bit loads per 8 channels
  __shared__ float xs[C1_SMEM];
  __shared__ float red[256][17];
  const int tpr = p.Cog >> 3, rpp = 256 / tpr;
  const int cg = threadIdx.x % tpr, rl = threadIdx.x / tpr;
  const int n = blockIdx.y;
  const int qa = blockIdx.x * C1V_ROWS;
  const int qb = min(p.nq, qa + C1V_ROWS);
  const T* __restrict__ X = reinterpret_cast<const T*>(p.X) + seq_base(p.x, n);
  const TY* __restrict__ dY = reinterpret_cast<const TY*>(p.dY) + seq_base(p.y, n);
  const int x0 = (p.q0 + qa) * p.si + min_off;
  const int nx = (qb - qa - 1) * p.si + span + 1;
  for (int i = threadIdx.x; i < nx; i += 256) {
    const int pos = x0 + i;
    xs[i] = (pos >= 0 && pos < p.x.len) ? ld_f(X + (int64_t)pos * p.x.s_row) : 0.f;
  }
  __syncthreads();
  const int nt = min(C1_TAPS, p.ntaps - tap0);
  float acc[C1_TAPS][8];
#pragma unroll
  for (int t = 0; t < C1_TAPS; ++t)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[t][i] = 0.f;
  const int yo = p.yoff[tap0];
  // tap offsets in registers (padded taps read offset 0 and are skipped), and the dY rows double-buffered:
  // the next four row loads are in flight while the current four are consumed (eight warps per SM cannot
  // hide a global-load round trip per batch otherwise; the loop was 125 of the kernel's 174 us)
  int offr[C1_TAPS];
#pragma unroll
  for (int t = 0; t < C1_TAPS; ++t) offr[t] = t < nt ? p.off[tap0 + t] - min_off : 0;
  constexpr int NV = YF ? 2 : 1;                // 128-bit loads per row and thread
  auto load4 = [&](uint4 (&u)[4 * NV], int q) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int qq = q + j * rpp;
      const int ypos = (p.q0 + qq) * p.so + yo;
      const bool ok = qq < qb && ypos >= 0 && ypos < p.y.len;
#pragma unroll
      for (int v = 0; v < NV; ++v)
        u[j * NV + v] = ok ? __ldg(reinterpret_cast<const uint4*>(dY + (int64_t)ypos * p.y.s_row) + cg * NV + v)
                           : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 u[4 * NV], un[4 * NV];
  if (!(dbg & 2)) load4(u, qa + rl);
  for (int q = qa + rl; q < qb && !(dbg & 2); q += 4 * rpp) {
    load4(un, q + 4 * rpp);                      // rows past qb load zeros
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int qq = q + j * rpp;
      if (qq >= qb) break;
      float dy[8];
      if (YF) {
        const uint4 a = u[j * NV], b = u[j * NV + NV - 1];
        dy[0] = __uint_as_float(a.x); dy[1] = __uint_as_float(a.y); dy[2] = __uint_as_float(a.z); dy[3] = __uint_as_float(a.w);
        dy[4] = __uint_as_float(b.x); dy[5] = __uint_as_float(b.y); dy[6] = __uint_as_float(b.z); dy[7] = __uint_as_float(b.w);
      } else {
        const uint32_t w[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          dy[2 * i] = __uint_as_float(w[i] << 16);
          dy[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
      }
      // all tap samples first (16 independent shared-memory loads in flight: with two warps per scheduler a
      // load -> 8 dependent FMAs chain per tap stalls on the shared-memory latency), padded taps read offset 0
      // into accumulators that are never written out
      const float* xp = xs + (qq - qa) * p.si;
      float xv[C1_TAPS];
#pragma unroll
      for (int t = 0; t < C1_TAPS; ++t) xv[t] = xp[offr[t]];
#pragma unroll
      for (int t = 0; t < C1_TAPS; ++t)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[t][i] = fmaf(xv[t], dy[i], acc[t][i]);
    }
#pragma unroll
    for (int j = 0; j < 4 * NV; ++j) u[j] = un[j];
  }
#pragma unroll
  for (int t = 0; t < C1_TAPS; t += 2) {       // two taps per reduction round
    if (t < nt) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) { red[threadIdx.x][i] = acc[t][i]; red[threadIdx.x][8 + i] = acc[t + 1][i]; }
      __syncthreads();
      for (int e = threadIdx.x; e < 2 * p.Cog; e += 256) {
        const int tt = e / p.Cog, c = e - tt * p.Cog;
        if (t + tt < nt) {
          float sum = 0.f;
          for (int r = 0; r < rpp; ++r) sum += red[r * tpr + (c >> 3)][8 * tt + (c & 7)];
          if (!(dbg & 1)) atomicAdd(p.dW + (int64_t)p.widx[tap0 + t + tt] * p.Cog + c, sum);   // [K][G=1][Cig=1][Cog]
        }
      }
    }
  }
}

// dW[t][ci] = sum_q X[q + off_t][ci] * dy[q + yoff]   (dy single channel, staged in smem; si = so = 1).
// Every X element is read once and feeds all taps.
template <typename T, typename TY>
__global__ void __launch_bounds__(256) tapwgrad_co1_kernel(const __grid_constant__ artic_tapwgrad_t p, int cw, int tap0,
                                                           int min_off, int span) {
  __shared__ float ys[C1_SMEM];
  __shared__ float red[256];
  const int cl = threadIdx.x % cw, pl = threadIdx.x / cw, npl = 256 / cw;
  const int c = blockIdx.z * cw + cl;
  const int n = blockIdx.y;
  // this CTA owns X rows [ra, rb) and needs dy[q] for q = r - off_t, i.e. q in [ra - max_off, rb - 1 - min_off]
  const int ra = p.q0 + min_off + blockIdx.x * C1_ROWS;
  const int rb = min(p.q0 + p.nq + min_off + span, ra + C1_ROWS);
  const T* __restrict__ X = reinterpret_cast<const T*>(p.X) + seq_base(p.x, n);
  const TY* __restrict__ dY = reinterpret_cast<const TY*>(p.dY) + seq_base(p.y, n);
  const int y0 = ra - (min_off + span);                // first staged q
  const int ny = (rb - ra) + span;
  const int yo = p.yoff[tap0];
  for (int i = threadIdx.x; i < ny; i += 256) {
    const int q = y0 + i;
    const int ypos = q + yo;
    ys[i] = (q >= p.q0 && q < p.q0 + p.nq && ypos >= 0 && ypos < p.y.len) ? ld_f(dY + (int64_t)ypos * p.y.s_row) : 0.f;
  }
  __syncthreads();
  const int nt = min(C1_TAPS, p.ntaps - tap0);
  float acc[C1_TAPS];
#pragma unroll
  for (int t = 0; t < C1_TAPS; ++t) acc[t] = 0.f;
  if (c < p.Cig) {
    // four independent row loads in flight per thread (one 2-byte load per iteration is latency-bound)
    for (int r = ra + pl; r < rb; r += 4 * npl) {
      float x[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = r + j * npl;
        x[j] = (rr < rb && rr >= 0 && rr < p.x.len) ? ld_f(X + (int64_t)rr * p.x.s_row + c) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int rr = r + j * npl;
        if (rr < rb) {
          const int yb = rr - y0;                      // index of q = r in ys
#pragma unroll
          for (int t = 0; t < C1_TAPS; ++t)
            if (t < nt) acc[t] = fmaf(x[j], ys[yb - p.off[tap0 + t]], acc[t]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < C1_TAPS; ++t) {      // static indexing: acc[] stays in registers
    if (t < nt) {
      __syncthreads();
      red[threadIdx.x] = acc[t];
      __syncthreads();
      if (pl == 0 && c < p.Cig) {
        float s = 0.f;
        for (int i = 0; i < npl; ++i) s += red[i * cw + cl];
        atomicAdd(p.dW + (int64_t)p.widx[tap0 + t] * p.Cig + c, s);   // [K][G=1][Cig][Cog=1]
      }
    }
  }
}

}  // namespace artic

using namespace artic;

// returns 1 if taken, 0 if not eligible
int artic_tapconv_ci1_try(const artic_tapconv_t* pp, cudaStream_t st) {
  const artic_tapconv_t& p = *pp;
  if (p.Cig != 1 || p.G != 1 || p.dtype != ARTIC_F32 || p.out_dtype != ARTIC_BF16) return 0;
  if (p.so != 1 || p.res || p.res2 || p.N > 65535) return 0;
  if ((p.res_pre && (reinterpret_cast<uintptr_t>(p.res_pre) & 15)) || (p.mask && (reinterpret_cast<uintptr_t>(p.mask) & 15))) return 0;
  const int tpr = p.Cog / 8;
  if (p.Cog % 8 != 0 || tpr < 1 || tpr > 256 || 256 % tpr != 0 || p.ntaps * p.Cog > CI1_WS) return 0;
  if ((p.y.s_row % 8) || (p.y.s_outer % 8) || (p.y.n_inner > 1 && (p.y.s_inner % 8))) return 0;
  if ((p.Y && (reinterpret_cast<uintptr_t>(p.Y) & 15)) || (p.Y2 && (reinterpret_cast<uintptr_t>(p.Y2) & 15))) return 0;
  int min_off = p.off[0], max_off = p.off[0];
  for (int t = 1; t < p.ntaps; ++t) { min_off = min(min_off, p.off[t]); max_off = max(max_off, p.off[t]); }
  const int span = max_off - min_off;
  if ((CI1_ROWS - 1) * p.si + span + 1 > CI1_XS) return 0;
  dim3 grid((unsigned)((p.nq + CI1_ROWS - 1) / CI1_ROWS), (unsigned)p.N);
  if (p.ntaps <= 5) tapconv_ci1_kernel<float, 5><<<grid, 256, 0, st>>>(p, min_off, span);
  else if (p.ntaps <= 16) tapconv_ci1_kernel<float, 16><<<grid, 256, 0, st>>>(p, min_off, span);
  else tapconv_ci1_kernel<float, 0><<<grid, 256, 0, st>>>(p, min_off, span);
  return 1;
}

int artic_tapconv_co1_try(const artic_tapconv_t* pp, cudaStream_t st) {
  const artic_tapconv_t& p = *pp;
  if (p.Cog != 1 || p.G != 1 || p.Cig < 32) return 0;
  const bool ob = p.out_dtype == ARTIC_BF16;
  if (p.dtype == ARTIC_BF16) {
    if (ob) launch_co1<__nv_bfloat16, __nv_bfloat16>(p, st);
    else launch_co1<__nv_bfloat16, float>(p, st);
  } else {
    if (ob) launch_co1<float, __nv_bfloat16>(p, st);
    else launch_co1<float, float>(p, st);
  }
  return 1;
}

int artic_tapwgrad_ci1_try(const artic_tapwgrad_t* pp, cudaStream_t st) {
  const artic_tapwgrad_t& p = *pp;
  if (p.G != 1 || p.N > 65535 || p.ntaps < 1) return 0;
  const bool ci1 = p.Cig == 1;
  const bool co1 = !ci1 && p.Cog == 1 && p.si == 1 && p.so == 1;
  if (!ci1 && !co1) return 0;
  int min_off = p.off[0], max_off = p.off[0];
  for (int t = 0; t < p.ntaps; ++t) {
    min_off = min(min_off, p.off[t]);
    max_off = max(max_off, p.off[t]);
    if (p.yoff[t] != p.yoff[0]) return 0;
  }
  const int span = max_off - min_off;
  if ((C1_ROWS - 1) * p.si + span + 1 > C1_SMEM || C1_ROWS + span > C1_SMEM) return 0;
  const int C = ci1 ? p.Cog : p.Cig;
  int cw = 32;
  while (cw < 256 && cw < C) cw <<= 1;
  const int rows = ci1 ? p.nq : p.nq + span;
  dim3 grid((unsigned)((rows + C1_ROWS - 1) / C1_ROWS), (unsigned)p.N, (unsigned)((C + cw - 1) / cw));
  const bool yb = p.y_dtype == ARTIC_BF16, xb = p.dtype == ARTIC_BF16;
  for (int tap0 = 0; tap0 < p.ntaps; tap0 += C1_TAPS) {
#define ARTIC_C1_LAUNCH(K)                                                                                        \
  do {                                                                                                            \
    if (xb && yb) K<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(p, cw, tap0, min_off, span);               \
    else if (xb) K<__nv_bfloat16, float><<<grid, 256, 0, st>>>(p, cw, tap0, min_off, span);                        \
    else if (yb) K<float, __nv_bfloat16><<<grid, 256, 0, st>>>(p, cw, tap0, min_off, span);                        \
    else K<float, float><<<grid, 256, 0, st>>>(p, cw, tap0, min_off, span);                                        \
  } while (0)
    const int tpr = p.Cog / 8;
    const bool vec = ci1 && p.Cog % 8 == 0 && tpr <= 256 && 256 % tpr == 0 && p.y.s_row % 8 == 0 &&
                     p.y.s_outer % 8 == 0 && (p.y.n_inner == 1 || p.y.s_inner % 8 == 0) &&
                     (reinterpret_cast<uintptr_t>(p.dY) & 15) == 0 && (yb || !xb);
    if (vec) {
      dim3 vgrid((unsigned)((p.nq + C1V_ROWS - 1) / C1V_ROWS), grid.y, 1);
      if (!yb) tapwgrad_ci1_vec_kernel<float, float><<<vgrid, 256, 0, st>>>(p, tap0, min_off, span, tc::g_debug[20]);
      else if (xb) tapwgrad_ci1_vec_kernel<__nv_bfloat16><<<vgrid, 256, 0, st>>>(p, tap0, min_off, span, tc::g_debug[20]);
      else tapwgrad_ci1_vec_kernel<float><<<vgrid, 256, 0, st>>>(p, tap0, min_off, span, tc::g_debug[20]);
    } else if (ci1) ARTIC_C1_LAUNCH(tapwgrad_ci1_kernel);
    else ARTIC_C1_LAUNCH(tapwgrad_co1_kernel);
#undef ARTIC_C1_LAUNCH
  }
  return 1;
}
