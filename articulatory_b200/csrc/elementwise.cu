// Small fused elementwise / reduction kernels of the path (HBM- or latency-bound).
#include "common.cuh"

namespace artic {

static inline int grid_for(int64_t n, int per_block = 256) {
  int64_t b = (n + per_block - 1) / per_block;
  const int64_t cap = 16LL * num_sms();
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

#define GRID_STRIDE(i, n) \
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n); i += (int64_t)gridDim.x * blockDim.x)

template <typename T>
__global__ void gen_input_kernel(const float* __restrict__ c, const T* __restrict__ ar, T* __restrict__ out, int B,
                                 int Cc, int Ca, int C, int Tn) {
  const int64_t n = (int64_t)B * Tn * C;
  GRID_STRIDE(i, n) {
    const int ch = (int)(i % C);
    const int64_t r = i / C;
    const int t = (int)(r % Tn);
    const int b = (int)(r / Tn);
    float v = 0.f;   // channels >= Cc + Ca are zero padding
    if (ch < Cc) v = c[((int64_t)b * Cc + ch) * Tn + t];
    else if (ch < Cc + Ca) v = ld_f(ar + (int64_t)b * Ca + (ch - Cc));
    st_f(out + i, v);
  }
}

template <typename T>
__global__ void gen_input_bwd_kernel(const T* __restrict__ dX, float* __restrict__ d_ar, int B, int Cc, int Ca, int C, int Tn) {
  // one thread per (b, a); Tn is small (<= a few hundred frames)
  const int64_t n = (int64_t)B * Ca;
  GRID_STRIDE(i, n) {
    const int a = (int)(i % Ca);
    const int b = (int)(i / Ca);
    float s = 0.f;
    for (int t = 0; t < Tn; ++t) s += ld_f(dX + ((int64_t)b * Tn + t) * C + Cc + a);
    d_ar[i] = s;
  }
}

template <typename T, typename TO>
__global__ void mean3_act_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c,
                                 TO* __restrict__ out, int64_t n, float slope) {
  GRID_STRIDE(i, n) {
    float v = (ld_f(a + i) + ld_f(b + i) + ld_f(c + i)) / 3.0f;
    st_f(out + i, v > 0.f ? v : slope * v);
  }
}

template <typename T>
__global__ void sum3_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c, T* __restrict__ out,
                            int64_t n) {
  GRID_STRIDE(i, n) { st_f(out + i, ld_f(a + i) + ld_f(b + i) + ld_f(c + i)); }
}

template <typename T>
__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, T* __restrict__ dpre, int64_t n) {
  GRID_STRIDE(i, n) { st_f(dpre + i, dy[i] * (1.f - y[i] * y[i])); }
}

// bf16x3 operand split, 4 elements per thread (128-bit load, two 64-bit stores)
__global__ void split4_kernel(const float4* __restrict__ s, uint2* __restrict__ hi, uint2* __restrict__ lo, int64_t n4) {
  GRID_STRIDE(i, n4) {
    const float4 v = __ldg(s + i);
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
    hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
  }
}
__global__ void split1_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                              int64_t n) {
  GRID_STRIDE(i, n) {
    const float v = __ldg(s + i);
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// Training windows cut on the device from a dataset that lives in HBM (SpeechCollater's index plan, bin/train.py:1009-1027,
// 1082-1097, bit for bit): one thread per output element of x | y | ar.
__global__ void cut_windows_kernel(const float* __restrict__ audio, const int64_t* __restrict__ aoff,
                                   const float* __restrict__ art, const int64_t* __restrict__ toff,
                                   const int32_t* __restrict__ pick, int B, int C, int frames, int aux, int hop, int ar_len,
                                   float* __restrict__ x, float* __restrict__ y, float* __restrict__ ar) {
  const int fx = frames + 2 * aux, T = frames * hop;
  const int64_t nx = (int64_t)B * C * fx, ny = (int64_t)B * T, na = (int64_t)B * ar_len;
  GRID_STRIDE(i, nx + ny + na) {
    if (i < nx) {
      const int t = (int)(i % fx);
      const int c = (int)((i / fx) % C);
      const int b = (int)(i / ((int64_t)fx * C));
      const int u = pick[2 * b], start = pick[2 * b + 1];
      x[i] = __ldg(art + toff[u] + (int64_t)(start - aux + t) * C + c);
    } else if (i < nx + ny) {
      const int64_t j = i - nx;
      const int b = (int)(j / T);
      const int u = pick[2 * b], start = pick[2 * b + 1];
      y[j] = __ldg(audio + aoff[u] + (int64_t)start * hop + (j % T));
    } else {
      const int64_t j = i - nx - ny;
      const int b = (int)(j / ar_len);
      const int u = pick[2 * b], start = pick[2 * b + 1];
      const int64_t src = (int64_t)start * hop - ar_len + (j % ar_len);      // left zero padding where the past runs out
      ar[j] = src >= 0 ? __ldg(audio + aoff[u] + src) : 0.f;
    }
  }
}

// PastFCEncoder (layers/pytorch_layers.py:426-460) in ONE launch: Linear -> LeakyReLU x (n - 1) -> Linear for one batch item
// per CTA.  Five dependent GEMV-sized layers (0.36 M weights) are pure launch latency as separate launches (5 x ~13 us at
// the head of every generator forward and of every decode chunk); here the activations stay in shared memory, the
// weights stream from L2 with coalesced reads (threads along the output features, the input features split over
// 1024 / cout thread groups).  Every layer's (storage-rounded) output is also written out: the backward needs them.
__device__ __forceinline__ float mlp_f(float v) { return v; }
__device__ __forceinline__ float mlp_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T mlp_t(float v);
template <> __device__ __forceinline__ float mlp_t<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 mlp_t<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
constexpr int MLP_THREADS = 1024;
constexpr int MLP_MAX_DIM = 1024;
__device__ __forceinline__ void mlp_ld8(const float* w, float (&x)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(w)), b = __ldg(reinterpret_cast<const float4*>(w) + 1);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
__device__ __forceinline__ void mlp_ld8(const __nv_bfloat16* w, float (&x)[8]) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(w));
  const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[i]));
    x[2 * i] = f.x; x[2 * i + 1] = f.y;
  }
}

// Each thread owns 8 consecutive output features (one 16- or 32-byte weight load per input feature) and a slice of
// the input features; cout / 8 threads cover a weight row, 1024 / (cout / 8) slices split the reduction, whose
// partial sums meet in shared memory.  A layer is 4 .. 16 dependent load rounds instead of 64 .. 128.
template <typename T>
__global__ void __launch_bounds__(MLP_THREADS) mlp_fwd_kernel(const __grid_constant__ artic_mlp_t p) {
  __shared__ float h[MLP_MAX_DIM];
  __shared__ __align__(16) float part[8 * MLP_THREADS];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < p.dims[0]; i += MLP_THREADS) {
    const T v = mlp_t<T>(__ldg(p.in + (int64_t)b * p.dims[0] + i));   // the input enters in the storage type, as the tape keeps it
    if (p.act0 != nullptr) reinterpret_cast<T*>(p.act0)[(int64_t)b * p.dims[0] + i] = v;
    h[i] = mlp_f(v);
  }
  __syncthreads();
  for (int l = 0; l < p.n_layers; ++l) {
    const int cin = p.dims[l], cout = p.dims[l + 1];
    const T* __restrict__ W = reinterpret_cast<const T*>(p.W[l]);            // prepared layout [cin][cout]
    const int tpr = cout >> 3;                                                // threads per weight row (host: cout % 8 == 0, tpr | 1024)
    const int groups = MLP_THREADS / tpr;
    const int jt = threadIdx.x % tpr, g = threadIdx.x / tpr;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 4
    for (int i = g; i < cin; i += groups) {
      float w[8];
      mlp_ld8(W + (int64_t)i * cout + jt * 8, w);
      const float x = h[i];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaf(x, w[c], acc[c]);
    }
    float* pp = part + (size_t)g * cout + jt * 8;
    *reinterpret_cast<float4*>(pp) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(pp + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    __syncthreads();
    if (threadIdx.x < cout) {
      float v = p.bias[l] != nullptr ? __ldg(p.bias[l] + threadIdx.x) : 0.f;
      for (int q = 0; q < groups; ++q) v += part[q * cout + threadIdx.x];
      if (l + 1 < p.n_layers) v = v > 0.f ? v : p.slope * v;
      const T o = mlp_t<T>(v);
      if (p.outs[l] != nullptr) reinterpret_cast<T*>(p.outs[l])[(int64_t)b * cout + threadIdx.x] = o;
      h[threadIdx.x] = mlp_f(o);
    }
    __syncthreads();
  }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, int64_t n) {
  GRID_STRIDE(i, n) { st_f(d + i, ld_f(s + i)); }
}

template <typename T>
__global__ void concat_time_kernel(const float* __restrict__ ar, const float* __restrict__ y, T* __restrict__ out, int B,
                                   int La, int Ly, int64_t pitch) {
  const int L = La + Ly;
  const int64_t n = (int64_t)B * L;
  GRID_STRIDE(i, n) {
    const int l = (int)(i % L);
    const int b = (int)(i / L);
    const float v = l < La ? ar[(int64_t)b * La + l] : y[(int64_t)b * Ly + (l - La)];
    st_f(out + (int64_t)b * pitch + l, v);
  }
}

template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int L, int Lout, int k, int stride,
                               int pad) {
  const int64_t n = (int64_t)B * Lout;
  GRID_STRIDE(i, n) {
    const int o = (int)(i % Lout);
    const int b = (int)(i / Lout);
    float s = 0.f;
    for (int j = 0; j < k; ++j) {
      const int l = o * stride - pad + j;
      if (l >= 0 && l < L) s += ld_f(x + (int64_t)b * L + l);
    }
    st_f(y + i, s / (float)k);  // count_include_pad=True: divisor is always k
  }
}

// Discriminator input preamble in ONE launch (one thread block per batch row, the row staged in shared memory): the
// signal [past samples | waveform] (models/hifigan.py:809-811 via bin/train.py:345-346), the AvgPool1d pyramid of the
// scale discriminators (:733-736) and the right reflect-padded copies of the period discriminators (:413-416).  As
// seven dependent 4-10 us launches this chain was a ~100 us bubble without a single tensor-core CTA at the head of
// every discriminator forward (profiles/r2_gaps_final.txt).  Arithmetic identical to the single kernels.
__global__ void __launch_bounds__(1024) disc_prep_kernel(const __grid_constant__ artic_disc_prep_t p) {
  extern __shared__ float sm[];
  const int n = blockIdx.x, T = p.T;
  if (p.x != nullptr) {
    for (int t = threadIdx.x; t < T; t += blockDim.x) sm[t] = __ldg(p.x + (int64_t)n * T + t);
  } else {
    const int b = n % p.B, Ty = T - p.La;
    const float* __restrict__ y = p.y[n / p.B] + (int64_t)b * Ty;
    const float* __restrict__ ar = p.La > 0 ? p.ar + (int64_t)b * p.La : nullptr;
    for (int t = threadIdx.x; t < T; t += blockDim.x) sm[t] = t < p.La ? __ldg(ar + t) : __ldg(y + (t - p.La));
  }
  __syncthreads();
  if (p.x_out != nullptr)
    for (int t = threadIdx.x; t < T; t += blockDim.x) p.x_out[(int64_t)n * T + t] = sm[t];
  for (int j = 0; j < p.n_xp; ++j) {
    const int Lp = p.xp_len[j];
    float* __restrict__ o = p.xp[j] + (int64_t)n * Lp;
    for (int t = threadIdx.x; t < Lp; t += blockDim.x) o[t] = sm[t < T ? t : 2 * (T - 1) - t];
  }
  const float* prev = sm;
  int lp = T;
  float* cur = sm + T;
  for (int l = 0; l < p.n_pool; ++l) {
    const int lo = p.pool_len[l];
    float* __restrict__ o = p.pool[l] + (int64_t)n * lo;
    for (int u = threadIdx.x; u < lo; u += blockDim.x) {
      float s = 0.f;
      for (int j = 0; j < p.k; ++j) {
        const int q = u * p.stride - p.pad + j;
        if (q >= 0 && q < lp) s += prev[q];
      }
      const float v = s / (float)p.k;      // count_include_pad = True: the divisor is always k
      cur[u] = v;
      o[u] = v;
    }
    __syncthreads();
    prev = cur; lp = lo; cur += lo;
  }
}

template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int L, int Lout, int k,
                                   int stride, int pad, int accumulate) {
  const int64_t n = (int64_t)B * L;
  GRID_STRIDE(i, n) {
    const int l = (int)(i % L);
    const int b = (int)(i / L);
    // outputs o with o*stride - pad <= l <= o*stride - pad + k - 1
    float s = 0.f;
    const int lo_num = l + pad - (k - 1);
    int o_lo = lo_num <= 0 ? 0 : (lo_num + stride - 1) / stride;
    int o_hi = (l + pad) / stride;
    if (o_hi > Lout - 1) o_hi = Lout - 1;
    for (int o = o_lo; o <= o_hi; ++o) s += ld_f(dy + (int64_t)b * Lout + o);
    s /= (float)k;
    if (accumulate) s += ld_f(dx + i);
    st_f(dx + i, s);
  }
}

template <typename T>
__global__ void reflect_pad_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int L, int Lp) {
  const int64_t n = (int64_t)B * Lp;
  GRID_STRIDE(i, n) {
    const int l = (int)(i % Lp);
    const int b = (int)(i / Lp);
    const int src = l < L ? l : 2 * (L - 1) - l;
    y[i] = x[(int64_t)b * L + src];
  }
}

template <typename T>
__global__ void reflect_pad_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int L, int Lp, int accumulate) {
  const int64_t n = (int64_t)B * L;
  GRID_STRIDE(i, n) {
    const int l = (int)(i % L);
    const int b = (int)(i / L);
    float s = ld_f(dy + (int64_t)b * Lp + l);
    const int m = 2 * (L - 1) - l;  // padded index that mirrors onto l
    if (m >= L && m < Lp) s += ld_f(dy + (int64_t)b * Lp + m);
    if (accumulate) s += ld_f(dx + i);
    st_f(dx + i, s);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) sqerr_sum_kernel(const T* __restrict__ x, int64_t n, float target, float scale,
                                                        float* __restrict__ slot) {
  __shared__ float red[32];
  float s = 0.f;
  GRID_STRIDE(i, n) {
    const float d = ld_f(x + i) - target;
    s = fmaf(d, d, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(slot, s * scale);
}

template <typename T>
__global__ void sqerr_bwd_kernel(const T* __restrict__ x, int64_t n, float target, float scale, T* __restrict__ dx,
                                 int accumulate) {
  GRID_STRIDE(i, n) {
    float g = 2.f * scale * (ld_f(x + i) - target);
    if (accumulate) g += ld_f(dx + i);
    st_f(dx + i, g);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) l1_sum_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n,
                                                     float scale, float* __restrict__ slot) {
  __shared__ float red[32];
  float s = 0.f;
  GRID_STRIDE(i, n) s += fabsf(ld_f(a + i) - ld_f(b + i));
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(slot, s * scale);
}

template <typename T>
__global__ void l1_bwd_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n, float scale, T* __restrict__ da,
                              int accumulate) {
  GRID_STRIDE(i, n) {
    const float d = ld_f(a + i) - ld_f(b + i);
    float g = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
    if (accumulate) g += ld_f(da + i);
    st_f(da + i, g);
  }
}

// Fused feature-matching term: slot += sum_scale * sum |a - b| and da = grad_scale * sign(a - b) in one
// pass over the two feature maps (8 bf16 per thread and iteration when VEC).
template <bool VEC>
__global__ void __launch_bounds__(256) l1_fused_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b, int64_t n,
                                                            float sum_scale, float* __restrict__ slot, float grad_scale,
                                                            __nv_bfloat16* __restrict__ da) {
  __shared__ float red[32];
  float s = 0.f;
  if (VEC) {
    const int64_t n8 = n >> 3;
    const uint4* a4 = reinterpret_cast<const uint4*>(a);
    const uint4* b4 = reinterpret_cast<const uint4*>(b);
    uint4* d4 = reinterpret_cast<uint4*>(da);
    GRID_STRIDE(i, n8) {
      const uint4 ua = __ldg(a4 + i), ub = __ldg(b4 + i);
      const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
      uint32_t wo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d0 = __uint_as_float(wa[j] << 16) - __uint_as_float(wb[j] << 16);
        const float d1 = __uint_as_float(wa[j] & 0xffff0000u) - __uint_as_float(wb[j] & 0xffff0000u);
        s += fabsf(d0) + fabsf(d1);
        const __nv_bfloat162 h = __floats2bfloat162_rn(d0 > 0.f ? grad_scale : (d0 < 0.f ? -grad_scale : 0.f),
                                                       d1 > 0.f ? grad_scale : (d1 < 0.f ? -grad_scale : 0.f));
        wo[j] = *reinterpret_cast<const uint32_t*>(&h);
      }
      d4[i] = make_uint4(wo[0], wo[1], wo[2], wo[3]);
    }
  } else {
    GRID_STRIDE(i, n) {
      const float d = ld_f(a + i) - ld_f(b + i);
      s += fabsf(d);
      st_f(da + i, d > 0.f ? grad_scale : (d < 0.f ? -grad_scale : 0.f));
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(slot, s * sum_scale);
}

__global__ void add_rows_kernel(const float* __restrict__ src, int64_t sp, float* __restrict__ dst, int64_t dp, int rows,
                                int cols) {
  const int64_t n = (int64_t)rows * cols;
  GRID_STRIDE(i, n) {
    const int c = (int)(i % cols);
    const int r = (int)(i / cols);
    dst[(int64_t)r * dp + c] += src[(int64_t)r * sp + c];
  }
}

__global__ void train_log_kernel(const float* __restrict__ slots, const float* __restrict__ sums,
                                 const float* __restrict__ numel, int R, float la, float ladv, float lfm,
                                 float* __restrict__ vals, float* __restrict__ running) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float sc = 0.f, mag = 0.f;
  for (int r = 0; r < R; ++r) {
    sc += sqrtf(sums[3 * r + 0]) / sqrtf(sums[3 * r + 1]);
    mag += sums[3 * r + 2] / numel[r];
  }
  if (R > 0) { sc /= (float)R; mag /= (float)R; }
  const float mel = slots[0], adv = slots[1], fm = slots[2], real = slots[3], fake = slots[4];
  const float gen = la * (sc + mag + mel) + ladv * (adv + lfm * fm);
  const float v[9] = {sc, mag, mel, adv, fm, gen, real, fake, real + fake};
  for (int i = 0; i < 9; ++i) {
    vals[i] = v[i];
    running[i] += v[i];
  }
}

}  // namespace artic

using namespace artic;
typedef __nv_bfloat16 bf16;

#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define DISPATCH(dtype, CALL_F32, CALL_BF16)                             \
  do {                                                                   \
    ARTIC_CHECK_ARG(dtype == ARTIC_F32 || dtype == ARTIC_BF16, "bad dtype"); \
    if (dtype == ARTIC_BF16) { CALL_BF16; } else { CALL_F32; }           \
    ARTIC_LAUNCH_CHECK();                                                \
    return ARTIC_OK;                                                     \
  } while (0)

extern "C" int artic_gen_input(const float* c, const void* ar_feats, void* out, int32_t B, int32_t Cc, int32_t Ca,
                               int32_t Cpad, int32_t T, int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(c && out && (ar_feats || Ca == 0), "null pointer");
  ARTIC_CHECK_ARG(Cpad >= Cc + Ca, "row pitch smaller than the channel count");
  const int64_t n = (int64_t)B * T * Cpad;
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (gen_input_kernel<float><<<grid_for(n), 256, 0, ST(stream)>>>(c, (const float*)ar_feats, (float*)out, B, Cc, Ca, Cpad, T)),
           (gen_input_kernel<bf16><<<grid_for(n), 256, 0, ST(stream)>>>(c, (const bf16*)ar_feats, (bf16*)out, B, Cc, Ca, Cpad, T)));
}

extern "C" int artic_gen_input_bwd(const void* dX, float* d_ar, int32_t B, int32_t Cc, int32_t Ca, int32_t Cpad,
                                   int32_t T, int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(dX && d_ar, "null pointer");
  ARTIC_CHECK_ARG(Cpad >= Cc + Ca, "row pitch smaller than the channel count");
  const int64_t n = (int64_t)B * Ca;
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (gen_input_bwd_kernel<float><<<grid_for(n, 64), 64, 0, ST(stream)>>>((const float*)dX, d_ar, B, Cc, Ca, Cpad, T)),
           (gen_input_bwd_kernel<bf16><<<grid_for(n, 64), 64, 0, ST(stream)>>>((const bf16*)dX, d_ar, B, Cc, Ca, Cpad, T)));
}

extern "C" int artic_mean3_act(const void* a, const void* b, const void* c, void* out_act, int64_t n, float slope,
                               int32_t dtype, int32_t out_dtype, void* stream) {
  ARTIC_CHECK_ARG(a && b && c && out_act, "null pointer");
  ARTIC_CHECK_ARG(out_dtype == ARTIC_F32 || out_dtype == ARTIC_BF16, "bad out dtype");
  if (n == 0) return ARTIC_OK;
  if (out_dtype == ARTIC_F32) {
    DISPATCH(dtype, (mean3_act_kernel<float, float><<<grid_for(n), 256, 0, ST(stream)>>>((const float*)a, (const float*)b, (const float*)c, (float*)out_act, n, slope)),
             (mean3_act_kernel<bf16, float><<<grid_for(n), 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, (const bf16*)c, (float*)out_act, n, slope)));
  }
  DISPATCH(dtype, (mean3_act_kernel<float, bf16><<<grid_for(n), 256, 0, ST(stream)>>>((const float*)a, (const float*)b, (const float*)c, (bf16*)out_act, n, slope)),
           (mean3_act_kernel<bf16, bf16><<<grid_for(n), 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, (const bf16*)c, (bf16*)out_act, n, slope)));
}

extern "C" int artic_sum3(const void* a, const void* b, const void* c, void* out, int64_t n, int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(a && b && c && out, "null pointer");
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (sum3_kernel<float><<<grid_for(n), 256, 0, ST(stream)>>>((const float*)a, (const float*)b, (const float*)c, (float*)out, n)),
           (sum3_kernel<bf16><<<grid_for(n), 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, (const bf16*)c, (bf16*)out, n)));
}

extern "C" int artic_tanh_bwd(const float* dy, const float* y, void* dpre, int64_t n, int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(dy && y && dpre, "null pointer");
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (tanh_bwd_kernel<float><<<grid_for(n), 256, 0, ST(stream)>>>(dy, y, (float*)dpre, n)),
           (tanh_bwd_kernel<bf16><<<grid_for(n), 256, 0, ST(stream)>>>(dy, y, (bf16*)dpre, n)));
}

extern "C" int artic_mlp_fwd(const artic_mlp_t* p, void* stream) {
  ARTIC_CHECK_ARG(p != nullptr && p->in != nullptr, "null pointer");
  ARTIC_CHECK_ARG(p->n_layers >= 1 && p->n_layers <= 8 && p->B >= 0, "1..8 layers");
  ARTIC_CHECK_ARG(p->dtype == ARTIC_F32 || p->dtype == ARTIC_BF16, "bad dtype");
  for (int l = 0; l <= p->n_layers; ++l) ARTIC_CHECK_ARG(p->dims[l] >= 1 && p->dims[l] <= MLP_MAX_DIM, "layer width out of range");
  for (int l = 0; l < p->n_layers; ++l) {
    ARTIC_CHECK_ARG(p->W[l] != nullptr, "null weight");
    const int co = p->dims[l + 1];
    if (co % 8 != 0 || MLP_THREADS % (co / 8) != 0 || (reinterpret_cast<uintptr_t>(p->W[l]) & 15) != 0) {
      set_error("artic_mlp_fwd: output widths must be 8 x a divisor of %d, weights 16-byte aligned", MLP_THREADS);
      return ARTIC_ENOSUP;
    }
  }
  if (p->B == 0) return ARTIC_OK;
  if (p->dtype == ARTIC_BF16) mlp_fwd_kernel<__nv_bfloat16><<<p->B, MLP_THREADS, 0, ST(stream)>>>(*p);
  else mlp_fwd_kernel<float><<<p->B, MLP_THREADS, 0, ST(stream)>>>(*p);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_cut_windows(const float* audio, const int64_t* audio_off, const float* art, const int64_t* art_off,
                                 const int32_t* pick, int32_t B, int32_t C, int32_t frames, int32_t aux, int32_t hop,
                                 int32_t ar_len, float* x, float* y, float* ar, void* stream) {
  ARTIC_CHECK_ARG(audio && audio_off && art && art_off && pick && x && y && (ar || ar_len == 0), "null pointer");
  ARTIC_CHECK_ARG(B >= 0 && C >= 1 && frames >= 1 && aux >= 0 && hop >= 1 && ar_len >= 0, "bad dims");
  const int64_t n = (int64_t)B * ((int64_t)C * (frames + 2 * aux) + (int64_t)frames * hop + ar_len);
  if (n == 0) return ARTIC_OK;
  cut_windows_kernel<<<grid_for(n), 256, 0, ST(stream)>>>(audio, audio_off, art, art_off, pick, B, C, frames, aux, hop, ar_len,
                                                         x, y, ar);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_split(const float* src, void* hi, int64_t plane, int64_t n, void* stream) {
  ARTIC_CHECK_ARG(src && hi && plane >= n && n >= 0, "bad arguments");
  if (n == 0) return ARTIC_OK;
  __nv_bfloat16* h = reinterpret_cast<__nv_bfloat16*>(hi);
  if (n % 4 == 0 && plane % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(hi) & 7) == 0)
    split4_kernel<<<grid_for(n / 4), 256, 0, ST(stream)>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(h),
                                                          reinterpret_cast<uint2*>(h + plane), n / 4);
  else
    split1_kernel<<<grid_for(n), 256, 0, ST(stream)>>>(src, h, h + plane, n);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_cast(const void* src, int32_t sd, void* dst, int32_t dd, int64_t n, void* stream) {
  ARTIC_CHECK_ARG(src && dst, "null pointer");
  ARTIC_CHECK_ARG((sd == ARTIC_F32 || sd == ARTIC_BF16) && (dd == ARTIC_F32 || dd == ARTIC_BF16), "bad dtype");
  if (n == 0) return ARTIC_OK;
  const int g = grid_for(n);
  if (sd == ARTIC_F32 && dd == ARTIC_F32) cast_kernel<float, float><<<g, 256, 0, ST(stream)>>>((const float*)src, (float*)dst, n);
  else if (sd == ARTIC_F32) cast_kernel<float, bf16><<<g, 256, 0, ST(stream)>>>((const float*)src, (bf16*)dst, n);
  else if (dd == ARTIC_F32) cast_kernel<bf16, float><<<g, 256, 0, ST(stream)>>>((const bf16*)src, (float*)dst, n);
  else cast_kernel<bf16, bf16><<<g, 256, 0, ST(stream)>>>((const bf16*)src, (bf16*)dst, n);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_concat_time(const float* ar, const float* y, void* out, int32_t B, int32_t La, int32_t Ly,
                                 int64_t out_pitch, int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(y && out && (ar || La == 0), "null pointer");
  ARTIC_CHECK_ARG(out_pitch >= (int64_t)La + Ly, "pitch too small");
  const int64_t n = (int64_t)B * (La + Ly);
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (concat_time_kernel<float><<<grid_for(n), 256, 0, ST(stream)>>>(ar, y, (float*)out, B, La, Ly, out_pitch)),
           (concat_time_kernel<bf16><<<grid_for(n), 256, 0, ST(stream)>>>(ar, y, (bf16*)out, B, La, Ly, out_pitch)));
}

extern "C" int artic_disc_prep(const artic_disc_prep_t* p, void* stream) {
  ARTIC_CHECK_ARG(p != nullptr, "null params");
  ARTIC_CHECK_ARG(p->N >= 0 && p->T >= 1 && p->n_pool >= 0 && p->n_pool <= 4 && p->n_xp >= 0 && p->n_xp <= 8, "bad dims");
  ARTIC_CHECK_ARG(p->x != nullptr || (p->B >= 1 && p->N % p->B == 0 && p->N / p->B <= 2 && p->La >= 0 && p->La < p->T &&
                                      (p->La == 0 || p->ar != nullptr) && p->y[0] != nullptr && (p->N / p->B < 2 || p->y[1] != nullptr)),
                  "assembling the signal needs ar (La > 0) and one waveform batch per B rows");
  int64_t fl = p->T;
  int lp = p->T;
  for (int l = 0; l < p->n_pool; ++l) {
    ARTIC_CHECK_ARG(p->pool[l] != nullptr && p->k >= 1 && p->stride >= 1 && p->pad >= 0 &&
                    p->pool_len[l] == (lp + 2 * p->pad - p->k) / p->stride + 1, "bad pooling geometry");
    lp = p->pool_len[l];
    fl += lp;
  }
  for (int j = 0; j < p->n_xp; ++j)
    ARTIC_CHECK_ARG(p->xp[j] != nullptr && p->xp_len[j] >= p->T && p->xp_len[j] - p->T < p->T, "reflect pad must be smaller than the input");
  if (p->N == 0) return ARTIC_OK;
  const size_t smem = (size_t)fl * sizeof(float);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    if (smem > 200 * 1024 ||
        cudaFuncSetAttribute(disc_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
      cudaGetLastError();
      set_error("artic_disc_prep: a row of %lld floats does not fit in shared memory", (long long)fl);
      return ARTIC_ENOSUP;
    }
    smem_set = 200 * 1024;
  }
  disc_prep_kernel<<<(unsigned)p->N, 1024, smem, ST(stream)>>>(*p);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_avgpool1d(const void* x, void* y, int32_t B, int32_t L, int32_t Lout, int32_t k, int32_t stride,
                               int32_t pad, int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(x && y, "null pointer");
  ARTIC_CHECK_ARG(k >= 1 && stride >= 1 && pad >= 0 && Lout == (L + 2 * pad - k) / stride + 1, "bad pooling geometry");
  const int64_t n = (int64_t)B * Lout;
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (avgpool_kernel<float><<<grid_for(n), 256, 0, ST(stream)>>>((const float*)x, (float*)y, B, L, Lout, k, stride, pad)),
           (avgpool_kernel<bf16><<<grid_for(n), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)y, B, L, Lout, k, stride, pad)));
}

extern "C" int artic_avgpool1d_bwd(const void* dy, void* dx, int32_t B, int32_t L, int32_t Lout, int32_t k,
                                   int32_t stride, int32_t pad, int32_t accumulate, int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(dy && dx, "null pointer");
  ARTIC_CHECK_ARG(k >= 1 && stride >= 1 && pad >= 0 && Lout == (L + 2 * pad - k) / stride + 1, "bad pooling geometry");
  const int64_t n = (int64_t)B * L;
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (avgpool_bwd_kernel<float><<<grid_for(n), 256, 0, ST(stream)>>>((const float*)dy, (float*)dx, B, L, Lout, k, stride, pad, accumulate)),
           (avgpool_bwd_kernel<bf16><<<grid_for(n), 256, 0, ST(stream)>>>((const bf16*)dy, (bf16*)dx, B, L, Lout, k, stride, pad, accumulate)));
}

extern "C" int artic_reflect_pad_right(const void* x, void* y, int32_t B, int32_t L, int32_t Lp, int32_t dtype,
                                       void* stream) {
  ARTIC_CHECK_ARG(x && y, "null pointer");
  ARTIC_CHECK_ARG(Lp >= L && Lp - L < L, "reflect pad must be smaller than the input");
  const int64_t n = (int64_t)B * Lp;
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (reflect_pad_kernel<float><<<grid_for(n), 256, 0, ST(stream)>>>((const float*)x, (float*)y, B, L, Lp)),
           (reflect_pad_kernel<bf16><<<grid_for(n), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)y, B, L, Lp)));
}

extern "C" int artic_reflect_pad_right_bwd(const void* dy, void* dx, int32_t B, int32_t L, int32_t Lp,
                                           int32_t accumulate, int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(dy && dx, "null pointer");
  ARTIC_CHECK_ARG(Lp >= L && Lp - L < L, "reflect pad must be smaller than the input");
  const int64_t n = (int64_t)B * L;
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (reflect_pad_bwd_kernel<float><<<grid_for(n), 256, 0, ST(stream)>>>((const float*)dy, (float*)dx, B, L, Lp, accumulate)),
           (reflect_pad_bwd_kernel<bf16><<<grid_for(n), 256, 0, ST(stream)>>>((const bf16*)dy, (bf16*)dx, B, L, Lp, accumulate)));
}

extern "C" int artic_sqerr_sum(const void* x, int64_t n, float target, float scale, float* slot, int32_t dtype,
                               void* stream) {
  ARTIC_CHECK_ARG(x && slot, "null pointer");
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (sqerr_sum_kernel<float><<<grid_for(n, 1024), 256, 0, ST(stream)>>>((const float*)x, n, target, scale, slot)),
           (sqerr_sum_kernel<bf16><<<grid_for(n, 1024), 256, 0, ST(stream)>>>((const bf16*)x, n, target, scale, slot)));
}

extern "C" int artic_sqerr_bwd(const void* x, int64_t n, float target, float scale, void* dx, int32_t accumulate,
                               int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(x && dx, "null pointer");
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (sqerr_bwd_kernel<float><<<grid_for(n), 256, 0, ST(stream)>>>((const float*)x, n, target, scale, (float*)dx, accumulate)),
           (sqerr_bwd_kernel<bf16><<<grid_for(n), 256, 0, ST(stream)>>>((const bf16*)x, n, target, scale, (bf16*)dx, accumulate)));
}

extern "C" int artic_l1_sum(const void* a, const void* b, int64_t n, float scale, float* slot, int32_t dtype,
                            void* stream) {
  ARTIC_CHECK_ARG(a && b && slot, "null pointer");
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (l1_sum_kernel<float><<<grid_for(n, 1024), 256, 0, ST(stream)>>>((const float*)a, (const float*)b, n, scale, slot)),
           (l1_sum_kernel<bf16><<<grid_for(n, 1024), 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, n, scale, slot)));
}

extern "C" int artic_l1_bwd(const void* a, const void* b, int64_t n, float scale, void* da, int32_t accumulate,
                            int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(a && b && da, "null pointer");
  if (n == 0) return ARTIC_OK;
  DISPATCH(dtype, (l1_bwd_kernel<float><<<grid_for(n), 256, 0, ST(stream)>>>((const float*)a, (const float*)b, n, scale, (float*)da, accumulate)),
           (l1_bwd_kernel<bf16><<<grid_for(n), 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, n, scale, (bf16*)da, accumulate)));
}

extern "C" int artic_l1_sum_bwd(const void* a, const void* b, int64_t n, float sum_scale, float* slot, float grad_scale,
                                void* da, int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(a && b && slot && da, "null pointer");
  ARTIC_CHECK_ARG(dtype == ARTIC_BF16, "the fused kernel is bf16-only (use artic_l1_sum + artic_l1_bwd for fp32)");
  if (n == 0) return ARTIC_OK;
  const bool vec = (n % 8 == 0) && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(da)) & 15) == 0;
  if (vec) l1_fused_bf16_kernel<true><<<grid_for(n / 8, 1024), 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, n, sum_scale, slot, grad_scale, (bf16*)da);
  else l1_fused_bf16_kernel<false><<<grid_for(n, 1024), 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b, n, sum_scale, slot, grad_scale, (bf16*)da);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_add_rows(const float* src, int64_t src_pitch, float* dst, int64_t dst_pitch, int32_t rows,
                              int32_t cols, void* stream) {
  ARTIC_CHECK_ARG(src && dst, "null pointer");
  ARTIC_CHECK_ARG(rows >= 0 && cols >= 0 && src_pitch >= cols && dst_pitch >= cols, "bad geometry");
  const int64_t n = (int64_t)rows * cols;
  if (n == 0) return ARTIC_OK;
  add_rows_kernel<<<grid_for(n), 256, 0, ST(stream)>>>(src, src_pitch, dst, dst_pitch, rows, cols);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_train_log(const float* slots, const float* stft_sums, const float* stft_numel, int32_t R,
                               float lambda_aux, float lambda_adv, float lambda_fm, float* vals, float* running,
                               void* stream) {
  ARTIC_CHECK_ARG(slots && vals && running && (R == 0 || (stft_sums && stft_numel)), "null pointer");
  train_log_kernel<<<1, 32, 0, ST(stream)>>>(slots, stft_sums, stft_numel, R, lambda_aux, lambda_adv, lambda_fm, vals,
                                            running);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}
