// Weight preparation (weight-norm + relayout), its backward, and the fused Adam step.
#include "common.cuh"

namespace artic {

// scale[row] = g[row] / ||v[row]||, scale[rows + row] = ||v[row]||
__global__ void __launch_bounds__(256) wn_scale_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                                       float* __restrict__ scale, int rows, int64_t row_len) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const float* vr = v + (int64_t)row * row_len;
  float s = 0.f;
  for (int64_t e = threadIdx.x; e < row_len; e += blockDim.x) {
    const float x = vr[e];
    s = fmaf(x, x, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const float nrm = sqrtf(s);
    scale[row] = g[row] / nrm;
    scale[rows + row] = nrm;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) prep_permute_kernel(const float* __restrict__ v, const float* __restrict__ scale,
                                                           int64_t row_len, int K, int G, int A, int B, int64_t sk,
                                                           int64_t sg, int64_t sa, int64_t sb, int mg,
                                                           T* __restrict__ out) {
  const int64_t total = (int64_t)K * G * A * B;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(o % B);
    int64_t r = o / B;
    const int a = (int)(r % A);
    r /= A;
    const int g = (int)(r % G);
    const int k = (int)(r / G);
    const int64_t src = k * sk + g * sg + a * sa + b * sb;
    float w = v[src];
    if (scale != nullptr) w *= scale[src / row_len];
    // merge > 1: block-diagonal layout [K][G/mg][mg*A][mg*B] (off-diagonal blocks stay zero)
    const int64_t oo = mg == 1 ? o
                               : ((((int64_t)k * (G / mg) + g / mg) * (mg * A) + (g % mg) * A + a) * (int64_t)(mg * B) +
                                  (g % mg) * B + b);
    st_f(out + oo, w);
  }
}

// One block per torch row. dWp is [K][G][A][B]; (k,g,a,b) are recovered from the source
// element index by mixed-radix decomposition over the dims sorted by decreasing stride.
struct Unperm {
  int64_t stride[4];  // sorted descending
  int32_t dim_id[4];  // 0=k 1=g 2=a 3=b
  int32_t n;          // number of dims with extent > 1
};

__global__ void __launch_bounds__(256) unprep_kernel(const float* __restrict__ dWp, const float* __restrict__ v,
                                                     const float* __restrict__ scale, int rows, int64_t row_len, int G,
                                                     int A, int B, int mg, Unperm up, float* __restrict__ dv,
                                                     float* __restrict__ dg) {
  __shared__ float red[32];
  __shared__ float s_dot;
  const int row = blockIdx.x;
  const int64_t e0 = (int64_t)row * row_len;
  auto perm = [&](int64_t e) -> int64_t {
    int idx[4] = {0, 0, 0, 0};
    int64_t rem = e;
    for (int i = 0; i < up.n; ++i) {
      idx[up.dim_id[i]] = (int)(rem / up.stride[i]);
      rem -= (int64_t)idx[up.dim_id[i]] * up.stride[i];
    }
    if (mg == 1) return (((int64_t)idx[0] * G + idx[1]) * A + idx[2]) * B + idx[3];
    return (((int64_t)idx[0] * (G / mg) + idx[1] / mg) * (mg * A) + (idx[1] % mg) * A + idx[2]) * (int64_t)(mg * B) +
           (idx[1] % mg) * B + idx[3];
  };
  if (scale == nullptr) {
    for (int64_t e = threadIdx.x; e < row_len; e += blockDim.x) dv[e0 + e] += dWp[perm(e0 + e)];
    return;
  }
  float dot = 0.f;
  for (int64_t e = threadIdx.x; e < row_len; e += blockDim.x) dot = fmaf(dWp[perm(e0 + e)], v[e0 + e], dot);
  dot = block_sum(dot, red);
  if (threadIdx.x == 0) s_dot = dot;
  __syncthreads();
  dot = s_dot;
  const float s = scale[row], nrm = scale[rows + row];
  const float coef = s * dot / (nrm * nrm);
  for (int64_t e = threadIdx.x; e < row_len; e += blockDim.x)
    dv[e0 + e] += s * dWp[perm(e0 + e)] - coef * v[e0 + e];
  if (threadIdx.x == 0) dg[row] += dot / nrm;
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
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
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (gi - m[i]) * (1.f - h.beta1);      // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * h.beta2 + (1.f - h.beta2) * gi * gi; // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + h.eps;
    p[i] -= step_size * (mi / denom);
  }
}

__global__ void adam_tick_kernel(artic_adam_hyper_t* hyper) { hyper->step += 1; }

}  // namespace artic

using namespace artic;

extern "C" int artic_weight_prep(const float* v, const float* g, float* scale, int32_t rows, int64_t row_len,
                                 int32_t K, int32_t G, int32_t A, int32_t B, int64_t sk, int64_t sg, int64_t sa,
                                 int64_t sb, int32_t merge, void* out, int32_t dtype, void* stream) {
  ARTIC_CHECK_ARG(v && out, "null pointer");
  ARTIC_CHECK_ARG(merge >= 1 && G % merge == 0, "merge must divide the group count");
  ARTIC_CHECK_ARG(g == nullptr || scale != nullptr, "scale buffer required with weight norm");
  ARTIC_CHECK_ARG(rows >= 1 && row_len >= 1 && K >= 1 && G >= 1 && A >= 1 && B >= 1, "bad dims");
  ARTIC_CHECK_ARG((int64_t)rows * row_len == (int64_t)K * G * A * B, "element count mismatch");
  ARTIC_CHECK_ARG(dtype == ARTIC_F32 || dtype == ARTIC_BF16, "bad dtype");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (g != nullptr) wn_scale_kernel<<<rows, 256, 0, st>>>(v, g, scale, rows, row_len);
  const int64_t total = (int64_t)K * G * A * B;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  const float* sc = g != nullptr ? scale : nullptr;
  if (dtype == ARTIC_BF16)
    prep_permute_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(v, sc, row_len, K, G, A, B, sk, sg, sa, sb, merge,
                                                                reinterpret_cast<__nv_bfloat16*>(out));
  else
    prep_permute_kernel<float><<<blocks, 256, 0, st>>>(v, sc, row_len, K, G, A, B, sk, sg, sa, sb, merge,
                                                        reinterpret_cast<float*>(out));
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_weight_unprep(const float* dWp, const float* v, const float* g, const float* scale, int32_t rows,
                                   int64_t row_len, int32_t K, int32_t G, int32_t A, int32_t B, int64_t sk, int64_t sg,
                                   int64_t sa, int64_t sb, int32_t merge, float* dv, float* dg, void* stream) {
  ARTIC_CHECK_ARG(dWp && v && dv, "null pointer");
  ARTIC_CHECK_ARG(merge >= 1 && G % merge == 0, "merge must divide the group count");
  ARTIC_CHECK_ARG(g == nullptr || (scale != nullptr && dg != nullptr), "scale and dg required with weight norm");
  ARTIC_CHECK_ARG((int64_t)rows * row_len == (int64_t)K * G * A * B, "element count mismatch");
  Unperm up;
  const int64_t strides[4] = {sk, sg, sa, sb};
  const int32_t ext[4] = {K, G, A, B};
  up.n = 0;
  for (int d = 0; d < 4; ++d)
    if (ext[d] > 1) { up.stride[up.n] = strides[d]; up.dim_id[up.n] = d; ++up.n; }
  for (int i = 0; i < up.n; ++i)       // sort by decreasing stride
    for (int j = i + 1; j < up.n; ++j)
      if (up.stride[j] > up.stride[i]) {
        int64_t ts = up.stride[i]; up.stride[i] = up.stride[j]; up.stride[j] = ts;
        int32_t td = up.dim_id[i]; up.dim_id[i] = up.dim_id[j]; up.dim_id[j] = td;
      }
  for (int i = up.n; i < 4; ++i) { up.stride[i] = 1; up.dim_id[i] = 0; }
  // nestedness check: each stride must be the product of the extents of the smaller-stride dims
  int64_t expect = 1;
  for (int i = up.n - 1; i >= 0; --i) {
    ARTIC_CHECK_ARG(up.stride[i] == expect, "source layout is not a permutation of a contiguous tensor");
    expect *= ext[up.dim_id[i]];
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  unprep_kernel<<<rows, 256, 0, st>>>(dWp, v, g != nullptr ? scale : nullptr, rows, row_len, G, A, B, merge, up, dv, dg);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_adam_step(float* p, const float* g, float* m, float* v, int64_t n,
                               const artic_adam_hyper_t* hyper, void* stream) {
  ARTIC_CHECK_ARG(p && g && m && v && hyper, "null pointer");
  if (n == 0) return ARTIC_OK;
  int64_t blocks = (n + 1023) / 1024;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  adam_kernel<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, g, m, v, n, hyper);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_adam_tick(artic_adam_hyper_t* hyper, void* stream) {
  ARTIC_CHECK_ARG(hyper, "null pointer");
  adam_tick_kernel<<<1, 1, 0, reinterpret_cast<cudaStream_t>(stream)>>>(hyper);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}
