// Fused spectral losses: framing (reflect pad, centred window) -> shared-memory Stockham
// FFT of z = w*(x + i*y) (one complex FFT yields both spectra) -> magnitude -> loss partial
// sums, and the matching backward (spectrogram recomputed, adjoint FFT, overlap-add).
// No spectrogram ever touches HBM.  Algorithmic HBM bytes: read x, y; write dL/dx.
#include "common.cuh"

namespace artic {

struct FrameGeom {
  int T, N, hop, win, lpad;  // lpad = (N - win) / 2 : torch.stft centres the window in n_fft
};

__device__ __forceinline__ int reflect_index(int t, int T) {
  if (t < 0) t = -t;
  if (t >= T) t = 2 * (T - 1) - t;
  return t;
}

// Radix-2 Stockham autosort FFT (forward, e^{-i...}), natural-order in and out.
// Source is (r0,i0); returns 0 if the result ends in (r0,i0), 1 if in (r1,i1).
template <typename T, typename TW>
__device__ __forceinline__ int fft_stockham(T* r0, T* i0, T* r1, T* i1, const TW* twr, const TW* twi, int N) {
  int cur = 0;
  const int half = N >> 1;
  for (int s = 1; s < N; s <<= 1) {
    const T* xr = cur ? r1 : r0;
    const T* xi = cur ? i1 : i0;
    T* yr = cur ? r0 : r1;
    T* yi = cur ? i0 : i1;
    for (int b = threadIdx.x; b < half; b += blockDim.x) {
      const int q = b & (s - 1);
      const int ps = b - q;
      const T ar = xr[b], ai = xi[b];
      const T br = xr[b + half], bi = xi[b + half];
      const T wr = (T)twr[ps], wi = (T)twi[ps];
      const int o = q + 2 * ps;
      yr[o] = ar + br;
      yi[o] = ai + bi;
      const T dr = ar - br, di = ai - bi;
      yr[o + s] = dr * wr - di * wi;
      yi[o + s] = dr * wi + di * wr;
    }
    __syncthreads();
    cur ^= 1;
  }
  return cur;
}

__device__ __forceinline__ void make_twiddles(double* twr, double* twi, int N) {
  for (int k = threadIdx.x; k < (N >> 1); k += blockDim.x) {
    double s, c;
    sincospi(2.0 * (double)k / (double)N, &s, &c);
    twr[k] = c;
    twi[k] = -s;
  }
}
__device__ __forceinline__ void make_twiddles(float* twr, float* twi, int N) {
  for (int k = threadIdx.x; k < (N >> 1); k += blockDim.x) {
    float s, c;
    sincospif(2.0f * (float)k / (float)N, &s, &c);
    twr[k] = c;
    twi[k] = -s;
  }
}

// Load frame f of (x, y) into (re, im) with window and reflect padding.
template <typename T>
__device__ __forceinline__ void load_frame(const float* __restrict__ x, const float* __restrict__ y,
                                           const float* __restrict__ window, const FrameGeom& g, int f, T* re,
                                           T* im) {
  const int start = f * g.hop - (g.N >> 1);
  for (int n = threadIdx.x; n < g.N; n += blockDim.x) {
    T a = 0, b = 0;
    const int wn = n - g.lpad;
    if (wn >= 0 && wn < g.win) {
      const T w = (T)__ldg(window + wn);
      const int t = reflect_index(start + n, g.T);
      a = w * (T)__ldg(x + t);
      b = w * (T)__ldg(y + t);
    }
    re[n] = a;
    im[n] = b;
  }
}

// Split Z = FFT(x + i y) into the two Hermitian spectra at bin k (0 <= k <= N/2).
template <typename T>
__device__ __forceinline__ void split_bin(const T* zr, const T* zi, int k, int N, T& xr, T& xi, T& yr, T& yi) {
  const int k2 = (N - k) & (N - 1);
  const T ar = zr[k], ai = zi[k], br = zr[k2], bi = zi[k2];
  xr = (T)0.5 * (ar + br);
  xi = (T)0.5 * (ai - bi);
  yr = (T)0.5 * (ai + bi);
  yi = (T)-0.5 * (ar - br);
}

// smem layout (floats): r0[N] i0[N] r1[N] i1[N] twr[N/2] twi[N/2] red[32] extra[...]
struct Smem {
  float *r0, *i0, *r1, *i1, *twr, *twi, *red, *extra;
  __device__ Smem(float* base, int N) {
    r0 = base; i0 = r0 + N; r1 = i0 + N; i1 = r1 + N; twr = i1 + N; twi = twr + (N >> 1);
    red = twi + (N >> 1); extra = red + 32;
  }
};

__global__ void __launch_bounds__(512) stft_loss_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            FrameGeom g, const float* __restrict__ window, float eps,
                                                            float* __restrict__ sums) {
  extern __shared__ __align__(16) float smem_f[];
  Smem sm(smem_f, g.N);
  const int f = blockIdx.x, b = blockIdx.y;
  make_twiddles(sm.twr, sm.twi, g.N);
  load_frame(x + (int64_t)b * g.T, y + (int64_t)b * g.T, window, g, f, sm.r0, sm.i0);
  __syncthreads();
  const int cur = fft_stockham(sm.r0, sm.i0, sm.r1, sm.i1, sm.twr, sm.twi, g.N);
  const float* zr = cur ? sm.r1 : sm.r0;
  const float* zi = cur ? sm.i1 : sm.i0;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
  for (int k = threadIdx.x; k <= (g.N >> 1); k += blockDim.x) {
    float xr, xi, yr, yi;
    split_bin(zr, zi, k, g.N, xr, xi, yr, yi);
    const float xm = sqrtf(fmaxf(xr * xr + xi * xi, eps));
    const float ym = sqrtf(fmaxf(yr * yr + yi * yi, eps));
    const float d = ym - xm;
    s0 = fmaf(d, d, s0);
    s1 = fmaf(ym, ym, s1);
    s2 += fabsf(logf(ym) - logf(xm));
  }
  s0 = block_sum(s0, sm.red);
  s1 = block_sum(s1, sm.red);
  s2 = block_sum(s2, sm.red);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 0, s0);
    atomicAdd(sums + 1, s1);
    atomicAdd(sums + 2, s2);
  }
}

// Adjoint of the framing + rFFT: the caller has written conj(G[k]) for k <= N/2 (zeros above)
// into (gr, gi); result dx_frame[n] = w[n] * Re(FFT(conj G))[n] is overlap-added into dx.
template <typename TW>
__device__ __forceinline__ void adjoint_to_dx(float* gr, float* gi, float* orr, float* oi, const TW* twr, const TW* twi,
                                              const FrameGeom& g, const float* __restrict__ window, int f,
                                              float* __restrict__ dx) {
  const int cur = fft_stockham(gr, gi, orr, oi, twr, twi, g.N);
  const float* rr = cur ? orr : gr;
  const int start = f * g.hop - (g.N >> 1);
  for (int wn = threadIdx.x; wn < g.win; wn += blockDim.x) {
    const int n = wn + g.lpad;
    const int t = reflect_index(start + n, g.T);
    atomicAdd(dx + t, __ldg(window + wn) * rr[n]);
  }
}

// The forward transform of the backward pass runs in fp64: the log-magnitude gradient scales
// like 1/|X_k|^2, so the absolute rounding error of an fp32 FFT (~1e-7 * ||frame||) at the few
// near-silent bins dominates the whole gradient (measured 1.9e-3 relative vs 2e-4 with an fp64
// forward transform; torch's fp32 path sits at ~1e-4).  The adjoint transform stays fp32.
// smem (doubles): r0[N] i0[N] r1[N] i1[N] twr[N/2] twi[N/2]; the idle double pair is re-used as
// four fp32 arrays for the adjoint.
__global__ void __launch_bounds__(512) stft_loss_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                            FrameGeom g, const float* __restrict__ window, float eps,
                                                            const float* __restrict__ sums, float w_sc, float w_mag,
                                                            float inv_numel, float* __restrict__ dx) {
  extern __shared__ __align__(16) double smem_d[];
  const int N = g.N;
  double *r0 = smem_d, *i0 = r0 + N, *r1 = i0 + N, *i1 = r1 + N, *twr = i1 + N, *twi = twr + (N >> 1);
  const int f = blockIdx.x, b = blockIdx.y;
  make_twiddles(twr, twi, N);
  load_frame(x + (int64_t)b * g.T, y + (int64_t)b * g.T, window, g, f, r0, i0);
  __syncthreads();
  const int cur = fft_stockham(r0, i0, r1, i1, twr, twi, N);
  const double* zr = cur ? r1 : r0;
  const double* zi = cur ? i1 : i0;
  float* fbuf = reinterpret_cast<float*>(cur ? r0 : r1);  // the idle pair: 2N doubles = 4N floats
  float *gr = fbuf, *gi = fbuf + N, *orr = fbuf + 2 * N, *oi = fbuf + 3 * N;
  const float S0 = sums[0], S1 = sums[1];
  const float c_sc = (S0 > 0.f && S1 > 0.f) ? w_sc * rsqrtf(S0) * rsqrtf(S1) : 0.f;
  const float c_mag = w_mag * inv_numel;
  const int half = N >> 1;
  for (int k = threadIdx.x; k <= half; k += blockDim.x) {
    double xr, xi, yr, yi;
    split_bin(zr, zi, k, N, xr, xi, yr, yi);
    const double px = xr * xr + xi * xi;
    const float xm = (float)sqrt(fmax(px, (double)eps));
    const float ym = (float)sqrt(fmax(yr * yr + yi * yi, (double)eps));
    // d/dxm [ w_sc * sqrt(S0)/sqrt(S1) + w_mag/numel * |ln ym - ln xm| ]
    const float dl = logf(ym) - logf(xm);
    float gk = c_sc * (xm - ym) - c_mag * (dl > 0.f ? 1.f : (dl < 0.f ? -1.f : 0.f)) / xm;
    if (!(px >= (double)eps)) gk = 0.f;  // clamp(min=eps) passes gradient only where px >= eps
    const float sc = gk / xm;
    gr[k] = sc * (float)xr;
    gi[k] = -sc * (float)xi;  // conj
    if (k > 0 && k < half) {
      gr[N - k] = 0.f;
      gi[N - k] = 0.f;
    }
  }
  __syncthreads();
  adjoint_to_dx(gr, gi, orr, oi, twr, twi, g, window, f, dx + (int64_t)b * g.T);
}

// ---- mel ---------------------------------------------------------------------------
// Mel projection of the two amplitude spectra held in (ax, ay)[0..N/2] -> extra[0..2*n_mels):
// mel_x at extra[m], mel_y at extra[n_mels + m].  Uses extra[2*n_mels ...] as scratch.
// rng (optional): the filters are narrow triangles — rng[2m], rng[2m+1] = first / one-past-last bin with a non-zero
// weight in filter m; rng[2*n_mels + 2k], [.. + 1] = the filters touching bin k.
__device__ __forceinline__ void mel_project(const float* ax, const float* ay, const float* __restrict__ melmat,
                                            int n_bins, int n_mels, float* extra, const int* __restrict__ rng = nullptr) {
  const int ngroups = max(1, (int)blockDim.x / n_mels);
  float* part = extra + 2 * n_mels;  // [ngroups][2][n_mels]
  const int gi = threadIdx.x / n_mels, m = threadIdx.x % n_mels;
  if (gi < ngroups) {
    float sx = 0.f, sy = 0.f;
    const int k_lo = rng != nullptr ? __ldg(rng + 2 * m) : 0, k_hi = rng != nullptr ? __ldg(rng + 2 * m + 1) : n_bins;
    for (int k = k_lo + gi; k < k_hi; k += ngroups) {
      const float w = __ldg(melmat + (int64_t)k * n_mels + m);
      sx = fmaf(ax[k], w, sx);
      sy = fmaf(ay[k], w, sy);
    }
    part[(gi * 2 + 0) * n_mels + m] = sx;
    part[(gi * 2 + 1) * n_mels + m] = sy;
  }
  __syncthreads();
  if (threadIdx.x < n_mels) {
    float sx = 0.f, sy = 0.f;
    for (int i = 0; i < ngroups; ++i) {
      sx += part[(i * 2 + 0) * n_mels + threadIdx.x];
      sy += part[(i * 2 + 1) * n_mels + threadIdx.x];
    }
    extra[threadIdx.x] = sx;
    extra[n_mels + threadIdx.x] = sy;
  }
  __syncthreads();
}

template <bool BWD>
__global__ void __launch_bounds__(512) mel_loss_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                       FrameGeom g, const float* __restrict__ window,
                                                       const float* __restrict__ melmat, int n_mels, float eps,
                                                       float log_scale, float scale, float* __restrict__ slot,
                                                       float* __restrict__ dx, const int* __restrict__ rng = nullptr,
                                                       float loss_scale = 0.f) {
  // BWD: dx += scale * d(sum |..|)/dx and, when slot != nullptr, slot[0] += loss_scale * sum |..| (the backward
  // recomputes the whole forward, so the separate forward launch is redundant in a train step)
  extern __shared__ __align__(16) float smem_f[];
  Smem sm(smem_f, g.N);
  const int f = blockIdx.x, b = blockIdx.y;
  const int half = g.N >> 1, n_bins = half + 1;
  make_twiddles(sm.twr, sm.twi, g.N);
  load_frame(x + (int64_t)b * g.T, y + (int64_t)b * g.T, window, g, f, sm.r0, sm.i0);
  __syncthreads();
  const int cur = fft_stockham(sm.r0, sm.i0, sm.r1, sm.i1, sm.twr, sm.twi, g.N);
  const float* zr = cur ? sm.r1 : sm.r0;
  const float* zi = cur ? sm.i1 : sm.i0;
  float* ax = cur ? sm.r0 : sm.r1;  // other buffer: amplitudes
  float* ay = cur ? sm.i0 : sm.i1;
  for (int k = threadIdx.x; k <= half; k += blockDim.x) {
    float xr, xi, yr, yi;
    split_bin(zr, zi, k, g.N, xr, xi, yr, yi);
    ax[k] = sqrtf(fmaxf(xr * xr + xi * xi, eps));
    ay[k] = sqrtf(fmaxf(yr * yr + yi * yi, eps));
  }
  __syncthreads();
  mel_project(ax, ay, melmat, n_bins, n_mels, sm.extra, rng);
  if (!BWD) {
    float s = 0.f;
    if (threadIdx.x < n_mels) {
      const float lx = logf(fmaxf(sm.extra[threadIdx.x], eps)) * log_scale;
      const float ly = logf(fmaxf(sm.extra[n_mels + threadIdx.x], eps)) * log_scale;
      s = fabsf(lx - ly);
    }
    s = block_sum(s, sm.red);
    if (threadIdx.x == 0) atomicAdd(slot, s * scale);
    return;
  } else {
    // d/dmel_x of scale * |log(clamp(mel_x)) - log(clamp(mel_y))| * log_scale
    float* dm = sm.extra + 2 * n_mels;  // overwrite the (now dead) partial-sum scratch
    float labs = 0.f;
    if (threadIdx.x < n_mels) {
      const float mx = sm.extra[threadIdx.x], my = sm.extra[n_mels + threadIdx.x];
      const float lx = logf(fmaxf(mx, eps)) * log_scale;
      const float ly = logf(fmaxf(my, eps)) * log_scale;
      const float d = lx - ly;
      labs = fabsf(d);
      float gm = (d > 0.f ? scale : (d < 0.f ? -scale : 0.f)) * log_scale / fmaxf(mx, eps);
      if (!(mx >= eps)) gm = 0.f;
      dm[threadIdx.x] = gm;
    }
    if (slot != nullptr) {      // uniform branch
      labs = block_sum(labs, sm.red);
      if (threadIdx.x == 0) atomicAdd(slot, labs * loss_scale);
    }
    __syncthreads();
    // da[k] = sum_m dm[m] * melmat[k][m].  One WARP per bin (lanes along the mel index: coalesced rows, a
    // shuffle reduction) when the scratch area can hold da[]; a thread per bin walking its own row touches
    // 32 different cache lines per load instruction and made this loop the bulk of the kernel.
    const int ngroups = max(1, (int)blockDim.x / n_mels);
    const bool warp_rows = rng == nullptr && 2 * ngroups * n_mels >= n_mels + n_bins;
    float* da_s = dm + n_mels;
    if (warp_rows) {
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
      for (int k = warp; k <= half; k += nwarp) {
        const float* mrow = melmat + (int64_t)k * n_mels;
        float acc = 0.f;
        for (int m = lane; m < n_mels; m += 32) acc = fmaf(dm[m], __ldg(mrow + m), acc);
        acc = warp_sum(acc);
        if (lane == 0) da_s[k] = acc;
      }
      __syncthreads();
    }
    for (int k = threadIdx.x; k <= half; k += blockDim.x) {
      float xr, xi, yr, yi;
      split_bin(zr, zi, k, g.N, xr, xi, yr, yi);
      const float px = xr * xr + xi * xi;
      float da = 0.f;
      if (warp_rows) {
        da = da_s[k];
      } else if (rng != nullptr) {
        const float* mrow = melmat + (int64_t)k * n_mels;
        const int m_hi = __ldg(rng + 2 * n_mels + 2 * k + 1);
        for (int m = __ldg(rng + 2 * n_mels + 2 * k); m < m_hi; ++m) da = fmaf(dm[m], __ldg(mrow + m), da);
      } else {
        const float* mrow = melmat + (int64_t)k * n_mels;
        for (int m = 0; m < n_mels; ++m) da = fmaf(dm[m], __ldg(mrow + m), da);
      }
      if (!(px >= eps)) da = 0.f;
      const float sc = da / ax[k];
      ax[k] = sc * xr;
      ay[k] = -sc * xi;
      if (k > 0 && k < half) {
        ax[g.N - k] = 0.f;
        ay[g.N - k] = 0.f;
      }
    }
    __syncthreads();
    adjoint_to_dx(ax, ay, const_cast<float*>(zr), const_cast<float*>(zi), sm.twr, sm.twi, g, window, f, dx + (int64_t)b * g.T);
  }
}

static int check_geom(int B, int T, int n_fft, int hop, int win) {
  if (B < 0 || T < 1 || hop < 1 || win < 1 || win > n_fft) return 0;
  if (n_fft < 64 || n_fft > 4096 || (n_fft & (n_fft - 1))) return 0;
  if (T <= n_fft / 2) return 0;  // reflect padding needs pad < T (same rule as torch.stft)
  return 1;
}

static size_t smem_bytes(int N, int n_mels) {
  const int ngroups = n_mels > 0 ? (512 / n_mels > 0 ? 512 / n_mels : 1) : 0;
  return sizeof(float) * ((size_t)5 * N + 32 + 2 * n_mels + (size_t)2 * ngroups * n_mels + 16);
}

template <typename K>
static int ensure_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(%zu B smem): %s", bytes, cudaGetErrorString(e));
      return ARTIC_ECUDA;
    }
  }
  return ARTIC_OK;
}

static int fft_threads(int N) { return (N / 2) < 512 ? (N / 2) : 512; }

}  // namespace artic

using namespace artic;

extern "C" int artic_stft_loss_fwd(const float* x, const float* y, int32_t B, int32_t T, int32_t n_fft, int32_t hop,
                                   int32_t win_length, const float* window, float eps, float* sums, void* stream) {
  ARTIC_CHECK_ARG(x && y && window && sums, "null pointer");
  ARTIC_CHECK_ARG(check_geom(B, T, n_fft, hop, win_length), "unsupported STFT geometry");
  if (B == 0) return ARTIC_OK;
  FrameGeom g{T, n_fft, hop, win_length, (n_fft - win_length) / 2};
  const size_t sb = smem_bytes(n_fft, 0);
  int rc = ensure_smem(stft_loss_fwd_kernel, sb);
  if (rc) return rc;
  dim3 grid(1 + T / hop, B);
  stft_loss_fwd_kernel<<<grid, fft_threads(n_fft), sb, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, g, window, eps, sums);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_stft_loss_bwd(const float* x, const float* y, int32_t B, int32_t T, int32_t n_fft, int32_t hop,
                                   int32_t win_length, const float* window, float eps, const float* sums, float w_sc,
                                   float w_mag, float* dx, void* stream) {
  ARTIC_CHECK_ARG(x && y && window && sums && dx, "null pointer");
  ARTIC_CHECK_ARG(check_geom(B, T, n_fft, hop, win_length), "unsupported STFT geometry");
  if (B == 0) return ARTIC_OK;
  FrameGeom g{T, n_fft, hop, win_length, (n_fft - win_length) / 2};
  const size_t sb = sizeof(double) * (size_t)5 * n_fft;
  int rc = ensure_smem(stft_loss_bwd_kernel, sb);
  if (rc) return rc;
  const int frames = 1 + T / hop;
  const float inv_numel = 1.0f / ((float)B * (float)frames * (float)(n_fft / 2 + 1));
  dim3 grid(frames, B);
  stft_loss_bwd_kernel<<<grid, fft_threads(n_fft), sb, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, y, g, window, eps, sums, w_sc, w_mag, inv_numel, dx);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_mel_loss_fwd(const float* x, const float* y, int32_t B, int32_t T, int32_t n_fft, int32_t hop,
                                  int32_t win_length, const float* window, const float* melmat, int32_t n_mels,
                                  float eps, float log_scale, float scale, float* slot, void* stream) {
  ARTIC_CHECK_ARG(x && y && window && melmat && slot, "null pointer");
  ARTIC_CHECK_ARG(check_geom(B, T, n_fft, hop, win_length), "unsupported STFT geometry");
  ARTIC_CHECK_ARG(n_mels >= 1 && n_mels <= fft_threads(n_fft), "n_mels out of range");
  if (B == 0) return ARTIC_OK;
  FrameGeom g{T, n_fft, hop, win_length, (n_fft - win_length) / 2};
  const size_t sb = smem_bytes(n_fft, n_mels);
  int rc = ensure_smem(mel_loss_kernel<false>, sb);
  if (rc) return rc;
  dim3 grid(1 + T / hop, B);
  mel_loss_kernel<false><<<grid, fft_threads(n_fft), sb, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, y, g, window, melmat, n_mels, eps, log_scale, scale, slot, nullptr);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_mel_loss_fwd_bwd(const float* x, const float* y, int32_t B, int32_t T, int32_t n_fft, int32_t hop,
                                      int32_t win_length, const float* window, const float* melmat,
                                      const int32_t* mel_ranges, int32_t n_mels, float eps, float log_scale,
                                      float loss_scale, float* slot, float grad_scale, float* dx, void* stream) {
  ARTIC_CHECK_ARG(x && y && window && melmat && dx, "null pointer");
  ARTIC_CHECK_ARG(check_geom(B, T, n_fft, hop, win_length), "unsupported STFT geometry");
  ARTIC_CHECK_ARG(n_mels >= 1 && n_mels <= fft_threads(n_fft), "n_mels out of range");
  if (B == 0) return ARTIC_OK;
  FrameGeom g{T, n_fft, hop, win_length, (n_fft - win_length) / 2};
  const size_t sb = smem_bytes(n_fft, n_mels);
  int rc = ensure_smem(mel_loss_kernel<true>, sb);
  if (rc) return rc;
  dim3 grid(1 + T / hop, B);
  mel_loss_kernel<true><<<grid, fft_threads(n_fft), sb, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, y, g, window, melmat, n_mels, eps, log_scale, grad_scale, slot, dx, mel_ranges, loss_scale);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}

extern "C" int artic_mel_loss_bwd(const float* x, const float* y, int32_t B, int32_t T, int32_t n_fft, int32_t hop,
                                  int32_t win_length, const float* window, const float* melmat, int32_t n_mels,
                                  float eps, float log_scale, float scale, float* dx, void* stream) {
  ARTIC_CHECK_ARG(x && y && window && melmat && dx, "null pointer");
  ARTIC_CHECK_ARG(check_geom(B, T, n_fft, hop, win_length), "unsupported STFT geometry");
  ARTIC_CHECK_ARG(n_mels >= 1 && n_mels <= fft_threads(n_fft), "n_mels out of range");
  if (B == 0) return ARTIC_OK;
  FrameGeom g{T, n_fft, hop, win_length, (n_fft - win_length) / 2};
  const size_t sb = smem_bytes(n_fft, n_mels);
  int rc = ensure_smem(mel_loss_kernel<true>, sb);
  if (rc) return rc;
  dim3 grid(1 + T / hop, B);
  mel_loss_kernel<true><<<grid, fft_threads(n_fft), sb, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, y, g, window, melmat, n_mels, eps, log_scale, scale, nullptr, dx);
  ARTIC_LAUNCH_CHECK();
  return ARTIC_OK;
}
