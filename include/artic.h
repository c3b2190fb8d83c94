/*
 * artic.h — C ABI of libartic_sm100.so, the B200 (sm_100a) kernels behind the
 * HiFi-GAN / HiFi-CAR hot path of articulatory/articulatory.
 *
 * The reference is 100 % Python: every entry point below replaces a PyTorch library
 * call site (cuDNN / cuBLAS / cuFFT / ATen) on the hot path; the call site is cited on
 * each declaration (paths relative to /root/reference/articulatory).  The reference-side
 * binding (a ctypes stub) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named h_*;
 *   - `stream` is a cudaStream_t passed as void*; nothing here allocates, synchronises
 *     or touches the host: every call is CUDA-graph capturable and re-entrant per stream;
 *   - return value: 0 = ok, otherwise a negative ARTIC_E* code; artic_last_error() returns
 *     a thread-local message;
 *   - activations are CHANNELS-LAST: element (n, l, c) of a sequence batch lives at
 *     base + (n / n_inner) * s_outer + (n % n_inner) * s_inner + l * s_row + c
 *     (plain batch: n_inner = 1).  dtype codes: ARTIC_F32 / ARTIC_BF16 (storage type of
 *     activations and prepared weights; accumulation is always fp32).
 */
#ifndef ARTIC_H_
#define ARTIC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARTIC_F32 0
#define ARTIC_BF16 1

#define ARTIC_OK 0
#define ARTIC_EINVAL (-1)  /* bad argument */
#define ARTIC_ECUDA (-2)   /* CUDA runtime / launch error */
#define ARTIC_ENOSUP (-3)  /* shape not supported by this kernel */

#define ARTIC_ACT_NONE 0
#define ARTIC_ACT_LRELU 1
#define ARTIC_ACT_TANH 2

#define ARTIC_MAX_TAPS 48

int artic_version(void);          /* 100 * major + minor */
const char* artic_arch(void);     /* "sm_100a" */
const char* artic_last_error(void);
/* Debug / tuning knobs of the kernels' planners (keys 0..31); not part of the reference surface. */
int artic_debug_set(int key, int value);
int artic_debug_get(int key);     /* current value of a knob (0 for an unknown key) */
/* Debug: device buffer (>= 4001 int64, zeroed) into which CTA 0 of the tensor-core conv kernel records a
 * (tag, clock64) timeline; NULL disables. */
int artic_debug_buffer(void* dev_buf);
/* Debug: SM-occupancy trace.  Every CTA of the tensor-core kernels appends one record of four int64
 * {launch id << 32 | kind << 28 | blockIdx.x, smid, globaltimer ns at start, at exit} after slot 0
 * (= record count, zero it before the traced region); kind 0..7 = problem index of a conv launch,
 * 8 = weight gradient.  dev_buf holds 1 + 4 * capacity_records int64; NULL disables.  The pointer is
 * baked into launches when they are enqueued (CUDA graphs: set it before capture). */
int artic_trace_buffer(void* dev_buf, long long capacity_records);

/* Addressing of one channels-last sequence batch (see header comment). */
typedef struct {
  int64_t s_outer, s_inner, s_row; /* element strides */
  int32_t n_inner;                 /* >= 1 */
  int32_t len;                     /* rows per sequence */
} artic_seq_t;

/*
 * Tap-gather GEMM: the one contraction behind Conv1d / ConvTranspose1d / Conv2d(k,1) /
 * Linear, forward and data-gradient (torch.nn.Conv1d, ConvTranspose1d, Conv2d, Linear
 * at models/hifigan.py:108-172,365-388,549-615; layers/residual_block.py:172-205;
 * layers/pytorch_layers.py:437-449; and their cuDNN dgrad).
 *
 *   for n < N, q in [q0, q0+nq), g < G, co < Cog, row = q*so + ro (skipped unless 0 <= row < y.len):
 *     acc = sum_{t < ntaps} sum_{ci < Cig} X[n, q*si + off[t], g*Cig+ci] * W[widx[t]][g][ci][co]
 *           (X rows outside [0, x.len) read as zero)
 *     v = alpha * acc + bias[g*Cog+co]
 *     v += res_pre[n,row,c];  v *= (mask[n,row,c] > 0 ? 1 : mask_slope);  v += res[..] + res2[..]
 *     Y[n,row,c] = v;   Y2[n,row,c] = act(v)        (any of bias/res_pre/mask/res/res2/Y/Y2 may be NULL)
 *
 * W is a PREPARED weight [K][G][Cig][Cog] (artic_weight_prep).  Wt (optional) is the SAME
 * weight in the transposed prepared layout [Wt_taps][G][Cog][Cig]; when it is given and the
 * shape is eligible (bf16 in/out, si == 1, Cig % 16 == 0, Cog % 32 == 0) the contraction
 * runs on the tcgen05 tensor-core kernel, otherwise on the CUDA-core kernel (same results
 * up to fp32 summation order).  res_pre/mask/res/res2/Y2 use Y's addressing.  Conv forward: si=stride, so=1, off[t]=t*dil-pad.  Conv dgrad /
 * ConvTranspose forward: one call per output phase r with so=stride, si=1.
 */
typedef struct {
  const void* X; const void* W; const float* bias;
  const void* res_pre; const void* mask; const void* res; const void* res2;
  void* Y; void* Y2;
  artic_seq_t x, y;
  int32_t N, G, Cig, Cog;
  int32_t q0, nq, si, so, ro;
  int32_t ntaps;
  int32_t off[ARTIC_MAX_TAPS];
  int32_t widx[ARTIC_MAX_TAPS];
  float alpha, mask_slope, act_slope;
  int32_t act;       /* ARTIC_ACT_* applied to Y2 */
  int32_t dtype;     /* storage type of X and W */
  int32_t out_dtype; /* storage type of Y, Y2, res_pre, mask, res, res2 */
  const void* Wt;    /* optional transposed prepared weight [Wt_taps][G][Cog][Cig] (same dtype as W) */
  int32_t Wt_taps;   /* number of taps K stored in Wt */
  int32_t reserved_;
  /* bf16x3 mode (fp32 storage, error-compensated tensor-core contraction): with dtype = out_dtype = ARTIC_F32,
   * X_sp = a SPLIT COPY of X (artic_split: bf16 `hi` plane with X's element strides, `lo` plane x_plane elements
   * further) and Wt_sp = the split copy of Wt (lo plane w_plane elements further), an eligible shape runs on the
   * tcgen05 kernel as x_hi*w_hi + x_hi*w_lo + x_lo*w_hi (fp32 accumulate, ~2^-16 relative operand error);
   * otherwise the fp32 CUDA-core kernel reads X / W.  Y_sp / Y2_sp (optional): split copies of Y / Y2 (planes
   * y_plane elements apart, Y's element strides), ALWAYS written when given, whichever kernel ran. */
  const void* X_sp; const void* Wt_sp;
  void* Y_sp; void* Y2_sp;
  int64_t x_plane, w_plane, y_plane;
} artic_tapconv_t;

int artic_tapconv(const artic_tapconv_t* p, void* stream);
/* n (<= 64) INDEPENDENT problems (no output of one is an input of another) in as few launches as possible:
 * tensor-core eligible problems share grids of per-problem persistent CTA ranges (the phases of a strided
 * data gradient or transposed conv, the three MRF blocks of a generator stage, the sub-discriminators at
 * one depth), the others run one by one. */
int artic_tapconv_multi(const artic_tapconv_t* ps, int32_t n, void* stream);

/*
 * Fused residual unit of an MRF block, forward (layers/residual_block.py:207-222, one (convs1[i], convs2[i]) pair):
 *     at  = lrelu(conv1(ax) + b1)      conv1: C -> C, kernel k, dilation dil, "same" padding
 *     xn  = conv2(at) + b2 + xres      conv2: C -> C, kernel k, dilation 1
 *     axn = lrelu(xn)
 * in ONE tcgen05 launch for the narrow stages (bf16, C = 32 or 64, odd k <= 11, both weights resident in shared memory):
 * the intermediate goes TMEM -> registers -> shared memory and feeds the second GEMM without an HBM round trip.
 * All tensors are plain (N, L, C) channels-last bf16 batches.  AX = lrelu(x) (the previous layer's activated output),
 * XRES = x.  AT (optional) receives `at` (the backward needs it), Y (optional) xn, Y2 (optional) axn.
 * W1t / W2t: the layers' transposed prepared weights [k][1][C_out][C_in] (artic_weights_prep `out_b`).
 * Returns ARTIC_ENOSUP when the shape is not covered (the caller then issues two artic_tapconv calls).
 *
 * mode = 1: the DATA GRADIENT of the same unit, also one launch (the two transposed convolutions, conv2's first):
 *     dt = conv2^T(gx) * lrelu'(at)          gn = conv1^T(dt) * lrelu'(ax) + gx
 * with AX = XRES = gx (gradient wrt xn), M1 = at, M2 = ax (the saved activations whose signs are the LeakyReLU masks),
 * W1t = conv2's weight and W2t = conv1's weight in the FORWARD prepared layout [k][1][C_in][C_out] (`out_f`), AT
 * receives dt (conv1's weight gradient needs it), Y receives gn; b1 / b2 / Y2 are unused.
 */
typedef struct {
  const void* AX; const void* XRES; const void* W1t; const void* W2t;
  const float* b1; const float* b2;
  void* AT; void* Y; void* Y2;
  int32_t N, L, C, k, dil;
  float slope;
  int32_t mode, reserved_;
  const void* M1; const void* M2;
} artic_resunit_t;

int artic_resunit_fwd(const artic_resunit_t* p, void* stream);

/*
 * Weight gradient of the same contraction (cuDNN wgrad):
 *   dW[widx[t]][g][ci][co] += sum_n sum_q X[n, q*si+off[t], g*Cig+ci] * dY[n, q*so+yoff[t], g*Cog+co]
 * dW is fp32 [K][G][Cig][Cog] and is ACCUMULATED into (zero it first).
 */
typedef struct {
  const void* X; const void* dY; float* dW;
  artic_seq_t x, y;
  int32_t N, G, Cig, Cog;
  int32_t q0, nq, si, so;
  int32_t ntaps;
  int32_t off[ARTIC_MAX_TAPS];
  int32_t yoff[ARTIC_MAX_TAPS];
  int32_t widx[ARTIC_MAX_TAPS];
  int32_t dtype;   /* storage type of X */
  int32_t y_dtype; /* storage type of dY */
  /* bf16x3 mode (see artic_tapconv_t): split copies of X and dY; used when both are given with fp32 X / dY */
  const void* X_sp; const void* dY_sp;
  int64_t x_plane, y_plane;
  /* optional bias gradient: dbias[g*Cog + co] += sum over all (n, row) of dY (fp32, G*Cog entries).  The tcgen05
   * kernel gets it from one extra accumulator (a ones-matrix times the dY tile it has staged anyway) — no second
   * pass over dY; otherwise a column-sum kernel is launched (artic_colsum).  Requires q0 + yoff = 0 and nq = y.len,
   * i.e. the launch covers every row of dY (true for every convolution's weight gradient). */
  float* dbias;
} artic_tapwgrad_t;

int artic_tapconv_wgrad(const artic_tapwgrad_t* p, void* stream);

/* Bias gradient: out[c] += sum over all (n, row) of a channels-last batch (C channels). */
int artic_colsum(const void* dY, const artic_seq_t* y, int32_t N, int32_t C, int32_t dtype,
                 float* out, void* stream);

/*
 * Weight preparation (replaces the per-forward torch._weight_norm hook,
 * models/hifigan.py:268-278,430-438 — w = g * v / ||v||, norm over all dims but 0 —
 * plus the relayout the kernels want), batched over all layers of a network in ONE launch
 * per pass.  One descriptor per layer, the table lives in DEVICE memory:
 *
 *   v      torch weight, fp32, viewed as [rows][row_len] for the norm and through the element
 *          strides (sk, sg, sa, sb) of its logical dims (K taps, G groups, A in-ch/group,
 *          B out-ch/group) for the relayout;  g = weight_g (rows floats) or NULL (plain weight)
 *   scale  2*rows floats (NULL iff g == NULL): g/||v|| in [0,rows), ||v|| in [rows,2*rows)
 *   out_f  'fwd' layout [K][G/m][a_pad][b_pad] in dtype_f          (NULL = skip)
 *   out_b  'bwd' (transposed) layout [K][G/m][b_pad][a_pad] in dtype_b  (NULL = skip)
 *          m = merge > 1 packs m narrow groups into one BLOCK-DIAGONAL super-group (a_pad >= m*A,
 *          b_pad >= m*B) so that grouped convs with < 32 channels per group still fill
 *          tensor-core tiles; a_pad / b_pad > m*A / m*B zero-pads odd channel counts (141 -> 144).
 *          Only the live entries are written: zero the buffers once.
 *   dWp    fp32 gradient in the 'fwd' layout (written by artic_tapconv_wgrad), or in the 'bwd' layout
 *          when dw_swapped != 0 (transposed-conv layers compute dW^T as a strided-conv weight gradient
 *          with the roles of X and dY exchanged)
 *   dv,dg  gradients of v and g in torch layout, OVERWRITTEN by artic_weights_unprep
 */
typedef struct {
  const float* v; const float* g; float* scale;
  void* out_f; void* out_b;
  const float* dWp; float* dv; float* dg;
  int64_t row_len, sk, sg, sa, sb;
  int64_t tile_begin;   /* exclusive prefix sum of artic_wperm_tiles() over the layers the generic kernels take */
  int64_t tile2_begin;  /* exclusive prefix sum of artic_wrow_tiles() (row-run kernels, the default) */
  int32_t rows, K, G, A, B, merge, a_pad, b_pad, dtype_f, dtype_b, dw_swapped;
  int32_t row_chunk;    /* row-run kernels: inner tiles per work unit (1..8), the value given to artic_wrow_tiles() */
} artic_wdesc_t;

/* Host helpers that size the two relayout tile spaces.  A layer is taken by the ROW-RUN kernels (whole torch rows of
 * 32 x TI x K tiles, 16-byte accesses on both sides) when artic_wrow_tiles() > 0 — every conv / linear / transposed
 * conv weight whose taps are the innermost torch index and K <= 352 — and then counts 0 generic tiles; otherwise it
 * counts artic_wperm_tiles() generic tiles (32 x 32 x <= 8 taps, any strides). */
int64_t artic_wperm_tiles(int32_t K, int32_t G, int32_t A, int32_t B);
int64_t artic_wrow_tiles(int32_t K, int32_t G, int32_t A, int32_t B, int64_t sk, int64_t sa, int64_t sb,
                         int32_t row_chunk);
/* scale + out_f + out_b of every descriptor (any_norm = 0 skips the norm pass); total_tiles / total_tiles2 = the sums
 * behind tile_begin / tile2_begin. */
int artic_weights_prep(const artic_wdesc_t* descs, int32_t n, int32_t any_norm, int64_t total_tiles,
                       int64_t total_tiles2, void* stream);
/* Backward of artic_weights_prep: dWp -> dv (and dg, through the weight-norm Jacobian). */
int artic_weights_unprep(const artic_wdesc_t* descs, int32_t n, int32_t any_norm, int64_t total_tiles,
                         int64_t total_tiles2, void* stream);

/* ---- small fused elementwise ops on the path ---------------------------------------- */

/* Generator input assembly (models/hifigan.py:209-211): out[b,t,:] = cat(c[b,:,t], ar_feats[b,:])
 * c is (B, Cc, T) channel-FIRST fp32 (the plugin boundary), ar_feats (B, Ca) in `dtype`
 * (may be NULL with Ca = 0); out is channels-last (B, T, Cpad) in `dtype`, Cpad >= Cc+Ca, the
 * extra channels written as zeros (pads 141 channels to a tensor-core friendly 160). */
int artic_gen_input(const float* c, const void* ar_feats, void* out, int32_t B, int32_t Cc,
                    int32_t Ca, int32_t Cpad, int32_t T, int32_t dtype, void* stream);
/* its backward wrt ar_feats: d_ar[b,a] = sum_t dX[b,t,Cc+a]  (fp32 out, overwritten) */
int artic_gen_input_bwd(const void* dX, float* d_ar, int32_t B, int32_t Cc, int32_t Ca, int32_t Cpad,
                        int32_t T, int32_t dtype, void* stream);

/* PastFCEncoder forward (layers/pytorch_layers.py:426-460: Linear, then [LeakyReLU, Linear] x (n_layers - 1)) in one
 * launch, one batch item per thread block.  in: (B, dims[0]) fp32; W[l]: prepared weight [dims[l]][dims[l+1]] in `dtype`
 * (artic_weights_prep `out_f` of a linear layer); bias[l]: fp32 or NULL; act0 (optional): the input cast to `dtype`;
 * outs[l] (optional): the layer's output (B, dims[l+1]) in `dtype` — activated for all but the last layer.  Widths
 * <= 1024; output widths 8 x a divisor of 1024 (else ARTIC_ENOSUP: the caller runs the layers one by one). */
typedef struct {
  const float* in;
  const void* W[8];
  const float* bias[8];
  void* act0;
  void* outs[8];
  int32_t dims[9];
  int32_t B, n_layers, dtype;
  float slope;
} artic_mlp_t;
int artic_mlp_fwd(const artic_mlp_t* p, void* stream);

/* Discriminator input preamble in one launch (one thread block per batch row):
 *   signal   x (N, T) given, or assembled: row n = cat(ar[n % B] (La samples), y[n / B][n % B] (T - La samples))
 *            (models/hifigan.py:809-811 fed by bin/train.py:345-346; up to two waveform batches stacked along N);
 *            x_out (optional) receives it
 *   pool[l]  AvgPool1d(k, stride, pad, count_include_pad = True) of level l - 1 (level -1 = the signal),
 *            pool_len[l] = (len + 2 pad - k) / stride + 1 (hifigan.py:733-736)
 *   xp[j]    the signal right reflect-padded to xp_len[j] samples (hifigan.py:413-416)
 * Same arithmetic as artic_concat_time / artic_avgpool1d / artic_reflect_pad_right.  ARTIC_ENOSUP when a row with its
 * pyramid exceeds 200 KB of shared memory (the caller then uses those). */
typedef struct {
  const float* x;
  const float* ar;
  const float* y[2];
  float* x_out;
  float* pool[4];
  float* xp[8];
  int32_t pool_len[4];
  int32_t xp_len[8];
  int32_t N, B, T, La, n_pool, n_xp, k, stride, pad, reserved_;
} artic_disc_prep_t;
int artic_disc_prep(const artic_disc_prep_t* p, void* stream);

/* MRF average + activation (models/hifigan.py:226-230 and the LeakyReLU of the next
 * layer): out_act = lrelu((a+b+c)/3, slope); n elements; inputs in `dtype`, output in `out_dtype`. */
int artic_mean3_act(const void* a, const void* b, const void* c, void* out_act, int64_t n,
                    float slope, int32_t dtype, int32_t out_dtype, void* stream);

/* out = a + b + c (n elements of `dtype`): joins the input gradients of the three MRF blocks, which
 * run on concurrent streams (models/hifigan.py:226-228 backward). */
int artic_sum3(const void* a, const void* b, const void* c, void* out, int64_t n, int32_t dtype, void* stream);

/* dpre = dy * (1 - y*y)  (torch.nn.Tanh backward, models/hifigan.py:158); fp32 dy/y in, `dtype` out. */
int artic_tanh_bwd(const float* dy, const float* y, void* dpre, int64_t n, int32_t dtype, void* stream);

/* bf16x3 operand split: hi[i] = bf16(src[i]), lo[i] = bf16(src[i] - hi[i]); lo lives `plane` elements after hi. */
int artic_split(const float* src, void* hi, int64_t plane, int64_t n, void* stream);

/* Host-side counters of which kernel family took each contraction since the last reset (tests assert that every
 * eligible layer runs on the tensor cores): out[0..11] = conv {tcgen05 bf16, tcgen05 bf16x3, CUDA-core generic,
 * channel-1 kernels}, wgrad {tcgen05 bf16, tcgen05 bf16x3, CUDA-core generic, channel-1 kernels}, tcgen05 weight
 * gradients that also produced the bias gradient, tcgen05 conv launches with cluster weight multicast, fused residual units (artic_resunit_fwd: two convs each), reserved.
 * Counted when a launch is ENQUEUED (graph replays do not count).  reset != 0 clears them after the read. */
int artic_path_counts(int64_t* h_out, int32_t reset);

/* dtype conversion helpers (n elements). */
int artic_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream);

/* Training windows cut ON THE DEVICE from a dataset resident in HBM (SpeechCollater.__call__, bin/train.py:965-1098,
 * random_window branch :1009-1027 and AR slice :1082-1097; same integer indexing, bit exact):
 *   audio      all utterances' samples back to back (fp32), audio_off[u] = first sample of utterance u
 *   art        all utterances' features back to back, (T'_u, C) row-major each, art_off[u] = first element
 *   pick       B x 2 int32 (device): (utterance, start frame) per batch item, drawn by the host with the collater's RNG
 *   x (B, C, frames + 2*aux)   = art[start-aux : start+frames+aux].T
 *   y (B, frames*hop)          = audio[start*hop : start*hop + frames*hop]
 *   ar (B, ar_len) (optional)  = audio[start*hop - ar_len : start*hop], zero where the index is negative */
int artic_cut_windows(const float* audio, const int64_t* audio_off, const float* art, const int64_t* art_off,
                      const int32_t* pick, int32_t B, int32_t C, int32_t frames, int32_t aux, int32_t hop,
                      int32_t ar_len, float* x, float* y, float* ar, void* stream);

/* D input assembly (bin/train.py:345-346): out[b] = cat(ar[b] (La), y[b] (Ly)) rows, fp32 in,
 * `dtype` out with row pitch `out_pitch` elements (>= La+Ly, rest untouched). */
int artic_concat_time(const float* ar, const float* y, void* out, int32_t B, int32_t La, int32_t Ly,
                      int64_t out_pitch, int32_t dtype, void* stream);

/* dst[r, 0:cols] += src[r, 0:cols] for r < rows (fp32; row pitches in elements): adds the
 * waveform part of the discriminator-input gradient (bin/train.py:345-346 slices) to dL/dy_. */
int artic_add_rows(const float* src, int64_t src_pitch, float* dst, int64_t dst_pitch, int32_t rows,
                   int32_t cols, void* stream);

/* AvgPool1d(kernel, stride, padding, count_include_pad=True) on (B, L) signals
 * (models/hifigan.py:719-721,733-736). Lout = (L + 2*pad - k)/stride + 1. */
int artic_avgpool1d(const void* x, void* y, int32_t B, int32_t L, int32_t Lout, int32_t k,
                    int32_t stride, int32_t pad, int32_t dtype, void* stream);
/* backward: dx[b,l] (+)= ... ; dx is overwritten unless accumulate != 0 */
int artic_avgpool1d_bwd(const void* dy, void* dx, int32_t B, int32_t L, int32_t Lout, int32_t k,
                        int32_t stride, int32_t pad, int32_t accumulate, int32_t dtype, void* stream);

/* Right reflect pad of (B, L) -> (B, Lp) (F.pad(x,(0,n_pad),"reflect"), models/hifigan.py:413-416);
 * bit-exact index map: out[l] = x[l] for l < L else x[2*(L-1) - l]. */
int artic_reflect_pad_right(const void* x, void* y, int32_t B, int32_t L, int32_t Lp, int32_t dtype,
                            void* stream);
/* backward: dx[b,l] (+)= dy[b,l] + dy[b, 2(L-1)-l] (when that index is in [L, Lp)) */
int artic_reflect_pad_right_bwd(const void* dy, void* dx, int32_t B, int32_t L, int32_t Lp,
                                int32_t accumulate, int32_t dtype, void* stream);

/* ---- loss reductions (losses/adversarial_loss.py:54-55,113-117; feat_match_loss.py:41-53) ---- */

/* slot[0] += scale * sum_i (x_i - target)^2   (F.mse_loss against a constant; scale = w/numel) */
int artic_sqerr_sum(const void* x, int64_t n, float target, float scale, float* slot, int32_t dtype,
                    void* stream);
/* dx_i (=|+=) 2 * scale * (x_i - target) */
int artic_sqerr_bwd(const void* x, int64_t n, float target, float scale, void* dx, int32_t accumulate,
                    int32_t dtype, void* stream);
/* slot[0] += scale * sum_i |a_i - b_i|        (F.l1_loss; scale = w/numel) */
int artic_l1_sum(const void* a, const void* b, int64_t n, float scale, float* slot, int32_t dtype,
                 void* stream);
/* da_i (=|+=) scale * sign(a_i - b_i) */
int artic_l1_bwd(const void* a, const void* b, int64_t n, float scale, void* da, int32_t accumulate,
                 int32_t dtype, void* stream);

/* Fused feature-matching term (bf16): slot[0] += sum_scale * sum_i |a_i - b_i| and da_i = grad_scale *
 * sign(a_i - b_i) in one pass (losses/feat_match_loss.py:41-53 and its backward). */
int artic_l1_sum_bwd(const void* a, const void* b, int64_t n, float sum_scale, float* slot, float grad_scale,
                     void* da, int32_t dtype, void* stream);

/*
 * Device-side assembly of the scalars the reference logs every step (bin/train.py:289-369,
 * 415-421) from the raw accumulators, without a host sync:
 *   slots = [mel, adv, fm, real, fake] (already normalised means / sums of means),
 *   stft_sums = R x [sum (Y-X)^2, sum Y^2, sum |ln Y - ln X|], stft_numel = R element counts.
 * vals[0..8] = [sc, mag, mel, adv, fm, gen, real, fake, dis] of this step, running[i] += vals[i].
 * gen = lambda_aux*(sc+mag+mel) + lambda_adv*(adv + lambda_fm*fm);  R may be 0.
 */
int artic_train_log(const float* slots, const float* stft_sums, const float* stft_numel, int32_t R,
                    float lambda_aux, float lambda_adv, float lambda_fm, float* vals, float* running,
                    void* stream);

/* ---- spectral losses ------------------------------------------------------------------ */

/*
 * One resolution of MultiResolutionSTFTLoss (losses/stft_loss.py:16-40,50-61,71-82,101-118:
 * torch.stft(center, reflect, hann(win) zero-padded to n_fft) -> sqrt(clamp(re^2+im^2,1e-7))
 * -> Frobenius / log-L1 sums).  x = predicted, y = target, both (B, T) fp32.
 * Forward accumulates into sums[0] += sum (Y-X)^2, sums[1] += sum Y^2, sums[2] += sum |ln Y - ln X|.
 * window: win_length fp32 taps.  n_fft must be a power of two in [64, 4096].
 */
int artic_stft_loss_fwd(const float* x, const float* y, int32_t B, int32_t T, int32_t n_fft,
                        int32_t hop, int32_t win_length, const float* window, float eps,
                        float* sums, void* stream);
/*
 * Backward wrt x of  w_sc * sqrt(S0)/sqrt(S1) + w_mag * S2 / numel  with S* read from
 * `sums` ON THE DEVICE (no host sync).  dx (B, T) fp32 is ACCUMULATED into.
 * The spectrogram is recomputed (never stored in HBM).
 */
int artic_stft_loss_bwd(const float* x, const float* y, int32_t B, int32_t T, int32_t n_fft,
                        int32_t hop, int32_t win_length, const float* window, float eps,
                        const float* sums, float w_sc, float w_mag, float* dx, void* stream);

/*
 * MelSpectrogramLoss (losses/mel_loss.py:82-111,151-166): STFT -> sqrt(clamp(power, eps))
 * -> melmat (n_bins x n_mels, row-major fp32) -> clamp(eps) -> log (log_scale = 1 for ln,
 * 1/ln2, 1/ln10) -> L1.  Forward: slot[0] += scale * sum |mel(x) - mel(y)|  (scale = w/numel).
 */
int artic_mel_loss_fwd(const float* x, const float* y, int32_t B, int32_t T, int32_t n_fft, int32_t hop,
                       int32_t win_length, const float* window, const float* melmat, int32_t n_mels,
                       float eps, float log_scale, float scale, float* slot, void* stream);
/* dx (B,T) fp32 += d(scale * sum |mel(x) - mel(y)|)/dx */
int artic_mel_loss_bwd(const float* x, const float* y, int32_t B, int32_t T, int32_t n_fft, int32_t hop,
                       int32_t win_length, const float* window, const float* melmat, int32_t n_mels,
                       float eps, float log_scale, float scale, float* dx, void* stream);

/* Loss and gradient in ONE launch (the backward recomputes the forward): slot[0] += loss_scale * sum |..| when
 * slot != NULL, dx += grad_scale * d(sum |..|)/dx.  mel_ranges (optional, int32 [2 * n_mels + 2 * n_bins], device):
 * [2m], [2m+1] = first / one-past-last bin with a non-zero weight in filter m, then per bin k the first /
 * one-past-last filter touching it — the triangular filters make melmat ~97 % zeros. */
int artic_mel_loss_fwd_bwd(const float* x, const float* y, int32_t B, int32_t T, int32_t n_fft, int32_t hop,
                           int32_t win_length, const float* window, const float* melmat,
                           const int32_t* mel_ranges, int32_t n_mels, float eps, float log_scale,
                           float loss_scale, float* slot, float grad_scale, float* dx, void* stream);

/* ---- speech-to-EMA inversion encoder (models/pytorch_models.py:22-77) ------------------------------ */

/*
 * One bidirectional GRU layer, inference forward (torch.nn.GRU(batch_first=True, bidirectional=True),
 * models/pytorch_models.py:27,30,63-66), given the input projections of all steps:
 *   gi   (N, T, 2*3H) fp32 = [W_ih x + b_ih | W_ih_reverse x + b_ih_reverse], gates [r|z|n] per direction
 *        (a plain GEMM: artic_tapconv with one tap)
 *   w_hh (2, 3H, H) fp32 = weight_hh_l0, weight_hh_l0_reverse;  b_hh (2, 3H) likewise
 *   out  (N, T, 2H) fp32 = [forward h_t | reverse h_t]
 * All N sequences have T steps; H <= 256, H % 4 == 0.  One persistent cluster kernel: W_hh stays in shared memory for the
 * whole sequence, h is exchanged through distributed shared memory once per step.
 */
int artic_bigru_layer(const float* gi, const float* w_hh, const float* b_hh, float* out, int32_t N, int32_t T,
                      int32_t H, void* stream);

/* ---- optimiser (torch.optim.Adam + MultiStepLR, bin/train.py:372-383,424-435,1750-1789) ---- */

/* Device-resident hyper-parameters so a captured graph can be replayed across steps. */
typedef struct {
  float lr0, beta1, beta2, eps, gamma;
  int32_t step;           /* number of optimiser steps taken so far */
  int32_t n_milestones;
  int32_t milestones[8];
} artic_adam_hyper_t;

/* p, m, v updated in place from g over n fp32 elements; reads *hyper (device) for lr/step. */
int artic_adam_step(float* p, const float* g, float* m, float* v, int64_t n,
                    const artic_adam_hyper_t* hyper, void* stream);
/* The same update with the gradient read in `g_dtype` (ARTIC_F32 / ARTIC_BF16): the data-parallel step hands the bf16
 * wire buffer of the gradient exchange straight to the optimiser. */
int artic_adam_step_wire(float* p, const void* g, int32_t g_dtype, float* m, float* v, int64_t n,
                         const artic_adam_hyper_t* hyper, void* stream);
/* hyper->step += 1 (device side; call once after all artic_adam_step of one optimiser step) */
int artic_adam_tick(artic_adam_hyper_t* hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARTIC_H_ */
