// tcgen05 weight-gradient kernel of the tap-gather contraction (artic_tapconv_wgrad) for the
// bf16 convolutions (unit or strided input, groups, period layout):
//
//   dW[t][ci][co] += sum_{n,q} X[n, q + off[t], ci] * dY[n, q + yoff, co]
//
// GEMM view: M = input channels (TMEM lanes), N = output channels, K = positions.  Both
// operands are staged by TMA exactly as they sit in HBM (channels-last rows), i.e. "MN-major"
// for the tensor core: the X tile [positions + halo][ci] is loaded ONCE per position chunk and
// every tap reads it through a row-shifted shared-memory descriptor; each tap owns its own
// fp32 accumulator in TMEM (n_acc * BN <= 512 columns).
// Narrow layers (Cig = 64 / 32) put 2 / 4 TAPS side by side in the 128 MMA rows: the leading
// byte offset of the A descriptor is the byte distance between consecutive taps (dil rows),
// so the second/third/fourth 64/32-row block of A is the same tile shifted by one more tap.
// Positions are split over CTAs (split-K); partial sums are reduced into the fp32 dW with
// vector red.global.add, coalesced through a per-warp shared-memory transpose that overlays
// the (by then idle) operand stages.
// Input stride si > 1: X is staged as si phase panels (row r of phase ph = input row r * si + ph);
// the tap groups are cut per phase when that moves fewer operand bytes, so a CTA loads only the
// ONE panel its accumulators read and the co tile widens (up to 256).
//
// bf16x3 mode (template X3; fp32 X / dY with split copies, see tapconv_tc.cu): every position chunk is issued
// three times — (x_hi, dy_hi), (x_hi, dy_lo), (x_lo, dy_hi) — into the same accumulators.
//
// Bias gradient (artic_tapwgrad_t.dbias): the column sums of dY come from ONE MORE accumulator, D_bias = 1 x dY, i.e.
// an A operand of all ones (a small constant tile in shared memory; with every element equal, swizzle and the
// leading-dimension offset do not matter) against the dY tile that is staged anyway — every row of D_bias holds the
// column sums of the chunk.  Only the CTAs of ci block 0 / tap group 0 do it, so each (co tile, split) counts once.
// This replaces a separate column-sum launch per layer that re-read every dY from HBM.
//
// Zero padding / sequence boundaries: rows outside [0, len) are zero-filled by TMA; short
// sequences are packed back to back with their halos (pitch = L + span), the padding rows of
// dY being zero so that they contribute nothing.  The staging area is zeroed once so that rows
// no TMA ever writes are exact zeros rather than stale bits.
#include "tc_common.cuh"

namespace artic {
namespace tc {

constexpr int WG_THREADS = 192;
constexpr int WG_MAX_STAGES = 8;
constexpr int WG_EPI_BYTES = 0;   // the epilogue's transpose stages overlay the operand stages

struct WPlan {
  int32_t xrb, yrb;            // row bytes (= swizzle span) of the X / dY panels
  int32_t x_layout, y_layout;  // UMMA swizzle codes
  int32_t mci;                 // input channels per CTA (128, 64 or 32)
  int32_t slots;               // taps side by side in the 128 MMA rows (1, 2 or 4)
  int32_t nxp, nyp;            // panels per stage for X / dY
  int32_t x_rows, kp;          // staged X rows per panel, positions per chunk
  int32_t x_panel_bytes, y_panel_bytes, stage_bytes, n_stages;
  int32_t bn, n_nt, n_mb, n_tg, tpc;  // co tile, #co tiles, #ci blocks, #tap groups, taps per group
  int32_t n_acc;               // accumulators per CTA (= ceil(tpc / slots))
  int32_t tmem_cols;
  int32_t packed, seg_per_chunk, seg_pitch, chunks_per_seq, n_chunks, chunks_per_split, n_splits;
  int32_t boxr, nxb;           // plain: rows per X box, X boxes per panel
  int32_t min_off;             // first tap offset
  int32_t a_lbo;               // leading byte offset of the A descriptor
  int32_t n_ph;                // input-stride phases (= si); X panels per stage = n_ph * nxp
  int32_t tap_stride;          // tap-index distance between side-by-side slots
  int32_t n_acc_total, apc;    // accumulators over all CTAs of a (ci, co) tile; (max) accumulators per CTA
  int16_t tg_begin[ARTIC_MAX_TAPS + 1];   // tap group g owns accumulators [tg_begin[g], tg_begin[g + 1]); with an input
                                          // stride a group never mixes phases, so its CTAs load ONE phase panel of X
  long long* trace;            // SM-occupancy trace buffer or nullptr
  long long trace_cap;
  int32_t launch_id;
  int32_t epi_transposed;      // 1: coalesced reductions through a shared-memory transpose (debug key 18 = 1 turns it off)
  int32_t dbg_flags;           // debug key 13: 1 = skip the reductions, 2 = skip the staging-area zeroing (timing experiments)
  int32_t bias_acc;            // 1: CTAs with mb == 0 && tg == 0 also accumulate the bias gradient (ones x dY)
  int32_t ones_off;            // byte offset (from the staging base) of the all-ones A tile
  long long* dbg;              // artic_debug_buffer: CTA 0 writes clock64 marks to dbg[200 + tag] (tools/wgrad_timeline.py)
  // per accumulator: phase panel, row shift of slot 0, tap index of slot 0, number of valid slots
  int8_t acc_panel[ARTIC_MAX_TAPS];
  int16_t acc_shift[ARTIC_MAX_TAPS];
  int8_t acc_tap0[ARTIC_MAX_TAPS];
  int8_t acc_cnt[ARTIC_MAX_TAPS];
};

// MN-major shared-memory matrix descriptor: 64-element (SW128) / 32-element (SW64) channel blocks
// LBO bytes apart, 8-position groups SBO = 8 * row_bytes apart.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t row_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)(((8u * row_bytes) >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct WMaps {
  CUtensorMap x, y, x_lo, y_lo;
};

template <bool X3>
__global__ void __launch_bounds__(WG_THREADS, 1)
tapwgrad_tc_kernel(const __grid_constant__ artic_tapwgrad_t p, const __grid_constant__ WPlan pl,
                   const __grid_constant__ WMaps maps) {
  const CUtensorMap& map_x = maps.x;
  const CUtensorMap& map_y = maps.y;
  constexpr int R = X3 ? 3 : 1;     // passes per position chunk
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[WG_MAX_STAGES], empty[WG_MAX_STAGES];
  __shared__ __align__(8) uint64_t acc_full;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  auto mark = [&](int tag) { if (pl.dbg != nullptr && blockIdx.x == 0) pl.dbg[200 + tag] = clock64(); };
  if (threadIdx.x == 0) mark(1);
  const long long t_trace = (pl.trace != nullptr && threadIdx.x == 0) ? global_timer() : 0;
  if (pl.dbg_flags & 4) return;   // what-if timing (debug key 13 = 4): every CTA exits at once

  // ---- work decode: blockIdx.x -> (split z, co tile nt, ci block mb, tap group tg, group g)
  int w = blockIdx.x;
  const int z = w % pl.n_splits; w /= pl.n_splits;
  const int nt = w % pl.n_nt; w /= pl.n_nt;
  const int mb = w % pl.n_mb; w /= pl.n_mb;
  const int tg = w % pl.n_tg;
  const int g = w / pl.n_tg;
  const int c_begin = z * pl.chunks_per_split;
  const int c_end = min(pl.n_chunks, c_begin + pl.chunks_per_split);
  const int acc0 = pl.tg_begin[tg];
  const int n_acc = pl.tg_begin[tg + 1] - acc0;
  const bool do_bias = pl.bias_acc && mb == 0 && tg == 0;
  uint32_t ph_mask = 0;                     // phase panels this CTA's accumulators read
  for (int a = 0; a < n_acc; ++a) ph_mask |= 1u << pl.acc_panel[acc0 + a];
  const int n_ph_used = __popc(ph_mask);

  // ---- one-time setup: zero the staging area, barriers, TMEM
  if (!(pl.dbg_flags & 2)) {
    uint4* z4 = reinterpret_cast<uint4*>(smem_raw + (smem0 - smem_u32(smem_raw)));
    const int n16 = pl.n_stages * pl.stage_bytes / 16;
    for (int i = threadIdx.x; i < n16; i += WG_THREADS) z4[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (do_bias) {   // all-ones A tile: kp positions x (<= 128-byte rows) of bf16 1.0
    uint4* o4 = reinterpret_cast<uint4*>(smem_raw + (smem0 + (uint32_t)pl.ones_off - smem_u32(smem_raw)));
    for (int i = threadIdx.x; i < pl.kp * pl.xrb / 16; i += WG_THREADS) o4[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (threadIdx.x == 0) {
    prefetch_tmap(&map_x);
    prefetch_tmap(&map_y);
    if (X3) { prefetch_tmap(&maps.x_lo); prefetch_tmap(&maps.y_lo); }
    for (int i = 0; i < pl.n_stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)pl.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) mark(2);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      PipeState ps(pl.n_stages);
      const int cx0 = g * p.Cig + mb * pl.mci;
      const int cy0 = g * p.Cog + nt * pl.bn;
      const int xch = pl.xrb / 2, ych = pl.yrb / 2;
      for (int cc = c_begin * R; cc < c_end * R; ++cc) {
        const int c = cc / R, part = cc % R;
        const CUtensorMap* mx = (X3 && part == 2) ? &maps.x_lo : &map_x;
        const CUtensorMap* my = (X3 && part == 1) ? &maps.y_lo : &map_y;
        mbar_wait(&empty[ps.stage], ps.phase ^ 1);
        const uint32_t xs = smem0 + (uint32_t)ps.stage * pl.stage_bytes;
        const uint32_t ys = xs + (uint32_t)(pl.n_ph * pl.nxp) * pl.x_panel_bytes;
        if (!pl.packed) {
          const int n = c / pl.chunks_per_seq;
          const int qc = p.q0 + (c % pl.chunks_per_seq) * pl.kp;
          mbar_expect_tx(&full[ps.stage],
                         (uint32_t)(n_ph_used * pl.nxp * pl.nxb * pl.boxr * pl.xrb + pl.nyp * pl.kp * pl.yrb));
          for (int ph = 0; ph < pl.n_ph; ++ph)
            for (int pn = 0; pn < pl.nxp && ((ph_mask >> ph) & 1u); ++pn)
              for (int b = 0; b < pl.nxb; ++b)
                tma_load_4d(xs + (uint32_t)(ph * pl.nxp + pn) * pl.x_panel_bytes + (uint32_t)b * pl.boxr * pl.xrb, mx,
                            &full[ps.stage], cx0 + pn * xch, n % p.x.n_inner, (qc + b * pl.boxr) * p.si + pl.min_off + ph,
                            n / p.x.n_inner);
          for (int pn = 0; pn < pl.nyp; ++pn)
            tma_load_4d(ys + (uint32_t)pn * pl.y_panel_bytes, my, &full[ps.stage], cy0 + pn * ych, n % p.y.n_inner,
                        qc + p.yoff[0], n / p.y.n_inner);
        } else {
          const int n0 = c * pl.seg_per_chunk;
          mbar_expect_tx(&full[ps.stage],
                         (uint32_t)(pl.seg_per_chunk * pl.seg_pitch * (n_ph_used * pl.nxp * pl.xrb + pl.nyp * pl.yrb)));
          for (int j = 0; j < pl.seg_per_chunk; ++j) {
            const int n = n0 + j;  // n >= N: fully out of bounds -> zero fill
            for (int ph = 0; ph < pl.n_ph; ++ph)
              for (int pn = 0; pn < pl.nxp && ((ph_mask >> ph) & 1u); ++pn)
                tma_load_4d(xs + (uint32_t)(ph * pl.nxp + pn) * pl.x_panel_bytes + (uint32_t)j * pl.seg_pitch * pl.xrb,
                            mx, &full[ps.stage], cx0 + pn * xch, n % p.x.n_inner, p.q0 * p.si + pl.min_off + ph,
                            n / p.x.n_inner);
            for (int pn = 0; pn < pl.nyp; ++pn)
              tma_load_4d(ys + (uint32_t)pn * pl.y_panel_bytes + (uint32_t)j * pl.seg_pitch * pl.yrb, my,
                          &full[ps.stage], cy0 + pn * ych, n % p.y.n_inner, p.q0 + p.yoff[0], n / p.y.n_inner);
          }
        }
        ps.next();
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    {   // whole warp runs the loop (uniform registers); one elected lane issues the tcgen05 instructions
      const bool leader = elect_one();
      PipeState ps(pl.n_stages);
      // D fp32, A/B bf16, both MN-major, N = bn, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(pl.bn >> 3) << 17) | ((128u >> 4) << 24);
      const int ksteps = pl.kp / 16;
      // descriptors: constant high words, low word = 16-byte address | LBO << 16 (one add per k-step)
      const uint32_t a_hi = (((8u * (uint32_t)pl.xrb) >> 4) & 0x3fffu) | (1u << 14) | ((uint32_t)(pl.x_layout & 7) << 29);
      const uint32_t b_hi = (((8u * (uint32_t)pl.yrb) >> 4) & 0x3fffu) | (1u << 14) | ((uint32_t)(pl.y_layout & 7) << 29);
      const uint32_t a_lo0 = (((uint32_t)pl.a_lbo >> 4) & 0x3fffu) << 16;
      const uint32_t b_lo0 = (((uint32_t)pl.y_panel_bytes >> 4) & 0x3fffu) << 16;
      uint32_t accum = 0, accum_b = 0;
      const uint32_t ones16 = ((smem0 + (uint32_t)pl.ones_off) >> 4) & 0x3fffu;     // LBO = 0: every 64-row block reads the same ones
      for (int cc = c_begin * R; cc < c_end * R; ++cc) {
        mbar_wait(&full[ps.stage], ps.phase);
        tc_fence_after();
        if (leader && cc == c_begin * R) mark(20);
        const uint32_t xs = smem0 + (uint32_t)ps.stage * pl.stage_bytes;
        const uint32_t ys = xs + (uint32_t)(pl.n_ph * pl.nxp) * pl.x_panel_bytes;
        const uint32_t b16 = b_lo0 | ((ys >> 4) & 0x3fffu);
        for (int a = 0; a < n_acc; ++a) {
          const uint32_t xa = xs + (uint32_t)(pl.acc_panel[acc0 + a] * pl.nxp) * pl.x_panel_bytes +
                              (uint32_t)pl.acc_shift[acc0 + a] * pl.xrb;
          uint32_t ad = a_lo0 | ((xa >> 4) & 0x3fffu);
          uint32_t bd = b16;
          const uint32_t dcol = tmem_base + (uint32_t)a * pl.bn;
          uint32_t acc_k = accum;
#pragma unroll 4
          for (int k = 0; k < ksteps; ++k) {
            if (leader) umma_bf16(dcol, ((uint64_t)a_hi << 32) | ad, ((uint64_t)b_hi << 32) | bd, idesc, acc_k);
            acc_k = 1;
            ad += (uint32_t)pl.xrb;    // 16 positions * xrb bytes / 16
            bd += (uint32_t)pl.yrb;
          }
        }
        if (do_bias && (!X3 || cc % R != 2)) {      // bf16x3: dY = dy_hi + dy_lo, i.e. passes 0 and 1
          uint32_t ad = ones16, bd = b16;
          const uint32_t dcol = tmem_base + (uint32_t)n_acc * pl.bn;
#pragma unroll 4
          for (int k = 0; k < ksteps; ++k) {
            if (leader) umma_bf16(dcol, ((uint64_t)a_hi << 32) | ad, ((uint64_t)b_hi << 32) | bd, idesc, accum_b);
            accum_b = 1;
            ad += (uint32_t)pl.xrb;
            bd += (uint32_t)pl.yrb;
          }
        }
        accum = 1;
        if (leader) umma_commit(&empty[ps.stage]);
        ps.next();
      }
      if (leader) { umma_commit(&acc_full); mark(22); }
    }
  } else {
    // =============================== epilogue ===================================
    // Split-K partial sums are reduced into dW with red.global.add.v4.f32, one dW row (= TMEM lane) per lane.
    const int ew = warp & 3;
    const int row = ew * 32 + lane;           // MMA row = TMEM lane
    const int slot = row / pl.mci;            // which tap of the side-by-side group (0 when mci == 128)
    const int ci = mb * pl.mci + (row - slot * pl.mci);
    float4* stage = reinterpret_cast<float4*>(smem_raw + (smem0 - smem_u32(smem_raw))) + ew * (32 * 8);   // [32 rows][8 units]
    float** rowdst = reinterpret_cast<float**>(smem_raw + (smem0 + 4 * 32 * 8 * 16 - smem_u32(smem_raw))) + ew * 32;
    const int g8 = lane & 7, rsub = lane >> 3;
    if (c_end > c_begin) {
      mbar_wait(&acc_full, 0);
      tc_fence_after();
      if (threadIdx.x == 64) mark(30);
      for (int a = 0; a < n_acc; ++a) {
        const bool valid = slot < pl.acc_cnt[acc0 + a] && ci < p.Cig;
        float* dst = nullptr;
        if (valid) {
          const int tap = pl.acc_tap0[acc0 + a] + slot * pl.tap_stride;
          dst = p.dW + (((int64_t)p.widx[tap] * p.G + g) * p.Cig + ci) * p.Cog + nt * pl.bn;
        }
        if (pl.epi_transposed) rowdst[lane] = dst;
        for (int c0 = 0; c0 < pl.bn; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)a * pl.bn + c0, r);
          tmem_ld_wait();
          if (pl.epi_transposed) {
            // the slab goes through a swizzled per-warp stage (the operand staging area is free once acc_full has
            // fired) so that 8 lanes cover one dW row's 32 channels: whole 128-byte segments per reduction
            // instruction, half the L2 atomic operations of the row-per-lane form
#pragma unroll
            for (int u = 0; u < 8; ++u)
              stage[lane * 8 + (u ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * u]), __uint_as_float(r[4 * u + 1]),
                                                               __uint_as_float(r[4 * u + 2]), __uint_as_float(r[4 * u + 3]));
            __syncwarp();
            if (!(pl.dbg_flags & 1)) {
#pragma unroll
              for (int it = 0; it < 32; it += 4) {
                const int rr = it + rsub;
                float* d = rowdst[rr];
                const float4 f = stage[rr * 8 + (g8 ^ (rr & 7))];
                if (d != nullptr) red_add_v4(d + c0 + g8 * 4, f.x, f.y, f.z, f.w);
              }
            }
            __syncwarp();
          } else if (valid && !(pl.dbg_flags & 1)) {
            // posted reductions straight from the TMEM row-per-lane layout (one dW row per lane)
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              red_add_v4(dst + c0 + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]), __uint_as_float(r[i + 2]),
                         __uint_as_float(r[i + 3]));
          }
        }
      }
      if (do_bias && ew == 0) {
        // every row of the bias accumulator holds the column sums: lane i of the first lane quarter adds column i
        for (int c0 = 0; c0 < pl.bn; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + (uint32_t)n_acc * pl.bn + c0, r);
          tmem_ld_wait();
          float v = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) v = (lane == i) ? __uint_as_float(r[i]) : v;
          if (!(pl.dbg_flags & 1)) atomicAdd(p.dbias + (int64_t)g * p.Cog + nt * pl.bn + c0 + lane, v);
        }
      }
    }
  }

  if (threadIdx.x == 64) mark(31);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)pl.tmem_cols);
  }
  if (threadIdx.x == 0) mark(32);
  if (pl.trace != nullptr && threadIdx.x == 0) trace_cta(pl.trace, pl.trace_cap, pl.launch_id, 8, t_trace);
}

static int g_wg_smem = 0;
static int wg_max_smem() {
  if (g_wg_smem == 0) {
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (optin <= 0) optin = 227 * 1024;
    cudaFuncAttributes fa;
    int stat = 2048;
    if (cudaFuncGetAttributes(&fa, tapwgrad_tc_kernel<false>) == cudaSuccess) stat = (int)fa.sharedSizeBytes;
    if (cudaFuncGetAttributes(&fa, tapwgrad_tc_kernel<true>) == cudaSuccess && (int)fa.sharedSizeBytes > stat) stat = (int)fa.sharedSizeBytes;
    int dyn = optin - stat;
    if (cudaFuncSetAttribute(tapwgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess ||
        cudaFuncSetAttribute(tapwgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess) {
      cudaGetLastError();
      dyn = 48 * 1024;
    }
    g_wg_smem = dyn;
  }
  return g_wg_smem;
}

static CUresult encode_seq_map(EncodeTiledFn enc, CUtensorMap* map, const void* base, const artic_seq_t& s, int N,
                               int channels, int box_ch, int box_rows, int row_bytes, int row_stride) {
  const int ni = s.n_inner;
  cuuint64_t dims[4] = {(cuuint64_t)channels, (cuuint64_t)ni, (cuuint64_t)s.len, (cuuint64_t)((N + ni - 1) / ni)};
  cuuint64_t strides[3] = {(cuuint64_t)(ni > 1 ? s.s_inner : s.s_row) * 2, (cuuint64_t)s.s_row * 2, (cuuint64_t)s.s_outer * 2};
  if (dims[3] == 1 && strides[2] == 0) strides[2] = strides[1] * dims[2];
  // rows traversed with stride row_stride: a box extent of rows*stride loads `rows` rows
  cuuint32_t box[4] = {(cuuint32_t)box_ch, 1, (cuuint32_t)(box_rows * row_stride), 1};
  cuuint32_t es[4] = {1, 1, (cuuint32_t)row_stride, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, es,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

}  // namespace tc
}  // namespace artic

using namespace artic;

// returns 1 if the launch was taken, 0 if the shape is not eligible, <0 on error.
int artic_tapwgrad_tc_try(const artic_tapwgrad_t* pp, cudaStream_t st, int* bias_fused) {
  const artic_tapwgrad_t& p = *pp;
  *bias_fused = 0;
  if (tc::g_debug[1] || tc::g_debug[5]) return 0;
  const bool x3 = p.dtype == ARTIC_F32 && p.y_dtype == ARTIC_F32 && p.X_sp != nullptr && p.dY_sp != nullptr;
  if (!x3 && (p.dtype != ARTIC_BF16 || p.y_dtype != ARTIC_BF16)) return 0;
  if (x3 && (tc::g_debug[24] == 1 || (p.x_plane % 8) || (p.y_plane % 8))) return 0;
  const void* Xb = x3 ? p.X_sp : p.X;
  const void* Yb = x3 ? p.dY_sp : p.dY;
  if (p.si < 1 || p.si > 8 || p.so != 1) return 0;
  if (!(p.Cig == 32 || p.Cig == 64 || p.Cig % 128 == 0) || p.Cog % 32 != 0) return 0;
  if ((p.x.s_row % 8) || (p.x.s_outer % 8) || (p.x.n_inner > 1 && (p.x.s_inner % 8))) return 0;
  if ((p.y.s_row % 8) || (p.y.s_outer % 8) || (p.y.n_inner > 1 && (p.y.s_inner % 8))) return 0;
  if ((reinterpret_cast<uintptr_t>(Xb) & 15) || (reinterpret_cast<uintptr_t>(Yb) & 15) ||
      (reinterpret_cast<uintptr_t>(p.dW) & 15))
    return 0;
  for (int t = 1; t < p.ntaps; ++t)
    if (p.yoff[t] != p.yoff[0]) return 0;
  // positions past nq (chunk overrun) must fall outside dY so that TMA zero-fills them
  if (p.q0 + p.yoff[0] + p.nq < p.y.len) return 0;
  // taps must be uniformly spaced and ascending (conv: off[t] = t*dil - pad)
  const int step = p.ntaps > 1 ? p.off[1] - p.off[0] : 1;
  if (step < 1) return 0;
  for (int t = 1; t < p.ntaps; ++t)
    if (p.off[t] - p.off[t - 1] != step) return 0;
  if (p.ntaps > 127) return 0;
  tc::EncodeTiledFn enc = tc::encode_fn();
  if (enc == nullptr) return 0;

  tc::WPlan pl;
  memset(&pl, 0, sizeof(pl));
  const int si = p.si;
  pl.min_off = p.off[0];
  pl.n_ph = si;
  pl.mci = p.Cig >= 128 ? 128 : p.Cig;
  pl.slots = 128 / pl.mci;
  pl.n_mb = p.Cig / pl.mci;
  pl.xrb = p.Cig >= 64 ? 128 : 64;
  pl.x_layout = pl.xrb == 128 ? 2 : 4;
  pl.nxp = pl.mci == 128 ? 2 : 1;
  // taps t and t + u read the same phase panel, slot_rows rows apart (u = si / gcd(step, si))
  int gcd = step, tmp = si;
  while (tmp) { const int r = gcd % tmp; gcd = tmp; tmp = r; }
  const int u = si / gcd, slot_rows = step / gcd;
  pl.tap_stride = u;
  // accumulators: for every residue class r < u, the taps r, r+u, r+2u, ... in groups of `slots`
  int n_acc_total = 0, span_rows = 0, ext = 0;
  int class_begin[ARTIC_MAX_TAPS + 1], n_class = 0;      // accumulators of one residue class (= one phase panel) are contiguous
  for (int r = 0; r < u && r < p.ntaps; ++r) {
    const int cnt = (p.ntaps - r + u - 1) / u;
    class_begin[n_class++] = n_acc_total;
    for (int j = 0; j * pl.slots < cnt; ++j) {
      const int t0 = r + u * (j * pl.slots);
      const int a = n_acc_total++;
      if (a >= ARTIC_MAX_TAPS) return 0;
      pl.acc_panel[a] = (int8_t)((t0 * step) % si);
      pl.acc_shift[a] = (int16_t)((t0 * step) / si);
      pl.acc_tap0[a] = (int8_t)t0;
      pl.acc_cnt[a] = (int8_t)min(pl.slots, cnt - j * pl.slots);
      span_rows = max(span_rows, (int)pl.acc_shift[a] + (pl.acc_cnt[a] - 1) * slot_rows);
      ext = max(ext, (int)pl.acc_shift[a] + (pl.slots - 1) * slot_rows);
    }
  }
  pl.n_acc_total = n_acc_total;
  class_begin[n_class] = n_acc_total;
  const int class_max = n_class > 0 ? class_begin[1] - class_begin[0] : 1;   // class 0 is the largest
  // co tile: accumulators per CTA * bn <= 512 TMEM columns
  if (p.Cog % 64 != 0) pl.bn = 32;
  else if (p.Cog % 128 == 0 && n_acc_total <= 4) pl.bn = 128;
  else pl.bn = 64;
  bool per_phase = false;
  if (u > 1 && tc::g_debug[19] != 1) {
    // Input stride: cutting the tap groups per phase (below) lets a CTA load ONE phase panel of X instead of all
    // `si`, and the widest co tile that still holds one phase's accumulators in TMEM re-reads X the fewest times —
    // but it multiplies the CTAs that stream dY.  Operand bytes per position decide.
    int bn_pp = 0;
    for (int bn = 256; bn >= 64; bn >>= 1)
      if (p.Cog % bn == 0 && class_max * bn <= 512) { bn_pp = bn; break; }
    if (bn_pp > 0) {
      const int x_phase = (p.Cig >= 128 ? 2 : 1) * (p.Cig >= 64 ? 128 : 64);          // bytes per position of one phase panel set
      const int tg_old = (n_acc_total + 512 / pl.bn - 1) / (512 / pl.bn);
      int tg_pp = 0;
      for (int c = 0; c < n_class; ++c) tg_pp += (class_begin[c + 1] - class_begin[c] + 512 / bn_pp - 1) / (512 / bn_pp);
      const long long cost_old = (long long)(p.Cog / pl.bn) * tg_old * (si * x_phase + 2 * pl.bn);
      const long long cost_pp = (long long)(p.Cog / bn_pp) * tg_pp * (x_phase + 2 * bn_pp);
      if (cost_pp < cost_old) { per_phase = true; pl.bn = bn_pp; }
    }
  }
  if (tc::g_debug[6] > 0 && p.Cog % tc::g_debug[6] == 0) pl.bn = tc::g_debug[6];
  pl.yrb = pl.bn >= 64 ? 128 : 64;
  pl.y_layout = pl.yrb == 128 ? 2 : 4;
  pl.nyp = pl.bn >= 64 ? pl.bn / 64 : 1;
  pl.n_nt = p.Cog / pl.bn;
  const int max_acc = 512 / pl.bn;
  pl.n_tg = 0;
  pl.apc = 0;
  if (per_phase && !(tc::g_debug[6] > 0)) {
    for (int c = 0; c < n_class; ++c) {               // every class in equal chunks of <= max_acc accumulators
      const int cnt = class_begin[c + 1] - class_begin[c];
      const int parts = (cnt + max_acc - 1) / max_acc, per = (cnt + parts - 1) / parts;
      for (int a = class_begin[c]; a < class_begin[c + 1]; a += per) {
        pl.tg_begin[pl.n_tg++] = (int16_t)a;
        pl.apc = max(pl.apc, min(per, class_begin[c + 1] - a));
      }
    }
  } else {
    const int n_tg = (n_acc_total + max_acc - 1) / max_acc, per = (n_acc_total + n_tg - 1) / n_tg;
    for (int a = 0; a < n_acc_total; a += per) pl.tg_begin[pl.n_tg++] = (int16_t)a;
    pl.apc = per;
  }
  pl.tg_begin[pl.n_tg] = (int16_t)n_acc_total;
  pl.n_acc = pl.apc;
  // bias gradient: one more accumulator on the CTAs of ci block 0 / tap group 0, when TMEM has room for it
  const int tg0_acc = pl.tg_begin[1] - pl.tg_begin[0];
  pl.bias_acc = (p.dbias != nullptr && (tg0_acc + 1) * pl.bn <= 512 && tc::g_debug[21] != 1) ? 1 : 0;
  const int cols = (pl.bias_acc && tg0_acc + 1 > pl.apc ? tg0_acc + 1 : pl.apc) * pl.bn;
  pl.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  const int align = 2;  // keeps every TMA destination 128-byte aligned for 64-byte rows
  const int lpad = ((p.nq + span_rows + align - 1) / align) * align;
  pl.packed = (p.N >= 2 && lpad <= 64 && lpad * si <= 256) ? 1 : 0;
  if (ext > 1024) return 0;
  const int ones_bytes = pl.bias_acc ? 16 * 1024 : 0;      // kp <= 128 positions x 128-byte rows
  const int budget = tc::wg_max_smem() - 1024 - tc::WG_EPI_BYTES - ones_bytes;
  // positions per chunk: shrink until at least 3 stages fit in shared memory
  int kp_cap = 128;
  for (;;) {
    if (pl.packed) {
      pl.seg_pitch = lpad;
      pl.seg_per_chunk = max(1, kp_cap / lpad);
      if (pl.seg_per_chunk > p.N) pl.seg_per_chunk = p.N;
      pl.kp = ((pl.seg_per_chunk * lpad + 15) / 16) * 16;
      pl.n_chunks = (p.N + pl.seg_per_chunk - 1) / pl.seg_per_chunk;
      pl.x_rows = pl.kp + ext + 8;
    } else {
      // Narrow layers (64-byte / 128-byte rows, one panel): the main loop is bound by the ROWS the TMA unit requests
      // (~7 clocks per row whatever its width: tools/wgrad_timeline.py: 1.4 k clocks per 64-position chunk of a
      // 32-channel layer for 8 or for 16 MMAs), so the X tile is ONE box of exactly the rows the taps read
      // (kp + span) instead of 64-row boxes that over-fetch up to 2x, and a chunk covers 128 positions.
      const bool narrow = si == 1 && p.Cig <= 64 && tc::g_debug[29] != 1;
      pl.kp = min(kp_cap, narrow ? 128 : 64);
      pl.chunks_per_seq = (p.nq + pl.kp - 1) / pl.kp;
      pl.n_chunks = p.N * pl.chunks_per_seq;
      if (si == 1 && pl.kp + span_rows <= 256 && tc::g_debug[29] != 1 && (narrow || tc::g_debug[29] != 2)) {   // wide layers too (key 29 = 2: narrow only)
        pl.boxr = ((pl.kp + span_rows + 7) / 8) * 8;
        pl.nxb = 1;
      } else {
        pl.boxr = si == 1 ? min(64, pl.kp) : 32;
        pl.nxb = (pl.kp + span_rows + pl.boxr - 1) / pl.boxr;
      }
      pl.x_rows = max(pl.nxb * pl.boxr, pl.kp + ext + 8);
    }
    pl.x_panel_bytes = ((pl.x_rows * pl.xrb + 1023) / 1024) * 1024;
    pl.y_panel_bytes = ((pl.kp * pl.yrb + 1023) / 1024) * 1024;
    pl.stage_bytes = pl.n_ph * pl.nxp * pl.x_panel_bytes + pl.nyp * pl.y_panel_bytes;
    pl.n_stages = budget / pl.stage_bytes;
    if (pl.n_stages >= 3 || kp_cap <= 32) break;
    kp_cap /= 2;
  }
  pl.a_lbo = pl.mci == 128 ? pl.x_panel_bytes : slot_rows * pl.xrb;
  if ((pl.a_lbo >> 4) > 0x3fff || (pl.y_panel_bytes >> 4) > 0x3fff) return 0;
  if (pl.n_stages < 2) return 0;
  if (pl.n_stages > tc::WG_MAX_STAGES) pl.n_stages = tc::WG_MAX_STAGES;
  const int64_t base = (int64_t)pl.n_nt * pl.n_mb * pl.n_tg * p.G;
  int64_t splits = num_sms() / base;
  if (splits > pl.n_chunks) splits = pl.n_chunks;
  {
    // Every CTA pays ~6 us of fixed cost (launch, staging, first-load latency, the split-K reduction of its
    // whole accumulator) while it holds an SM that the concurrent streams could use: give each split a main
    // loop of at least `min_clk` tensor-core clocks (debug key 14, in units of 1000 clocks; round 1: 6, 12.70 -> 12.55 ms;
    // re-swept on the round-2 build: 3 -> 11.36, 6 -> 11.27, 10 -> 11.23, 12 -> 11.19, 16 -> 11.26 ms: default 12) rather
    // than spreading a small layer over all SMs.
    const double per_mma = pl.bn / 2.0 > 32.0 + pl.bn / 4.0 ? pl.bn / 2.0 : 32.0 + pl.bn / 4.0;
    const double chunk_clk = (double)pl.apc * (pl.kp / 16) * per_mma * (x3 ? 3 : 1);
    const double min_clk = 1000.0 * (tc::g_debug[14] > 0 ? tc::g_debug[14] : tc::g_debug[14] < 0 ? 0 : 12);
    int64_t min_chunks = (int64_t)(min_clk / chunk_clk + 0.999);
    if (min_chunks < 1) min_chunks = 1;
    const int64_t cap = (pl.n_chunks + min_chunks - 1) / min_chunks;
    if (splits > cap) splits = cap;
  }
  if (splits < 1) splits = 1;
  pl.chunks_per_split = (int)((pl.n_chunks + splits - 1) / splits);
  pl.n_splits = (pl.n_chunks + pl.chunks_per_split - 1) / pl.chunks_per_split;
  if (pl.n_stages > (x3 ? 3 : 1) * pl.chunks_per_split + 1) pl.n_stages = (x3 ? 3 : 1) * pl.chunks_per_split + 1;
  if (pl.n_stages < 2) pl.n_stages = 2;
  const int64_t grid = base * pl.n_splits;
  if (grid > (1 << 30)) return 0;

  static thread_local tc::WMaps maps;
  CUresult rc = tc::encode_seq_map(enc, &maps.x, Xb, p.x, p.N, p.G * p.Cig, pl.xrb / 2,
                                   pl.packed ? pl.seg_pitch : pl.boxr, pl.xrb, si);
  if (rc == CUDA_SUCCESS && x3)
    rc = tc::encode_seq_map(enc, &maps.x_lo, reinterpret_cast<const __nv_bfloat16*>(Xb) + p.x_plane, p.x, p.N, p.G * p.Cig,
                            pl.xrb / 2, pl.packed ? pl.seg_pitch : pl.boxr, pl.xrb, si);
  if (rc != CUDA_SUCCESS) { set_error("artic_tapconv_wgrad: cuTensorMapEncodeTiled(X) failed (%d)", (int)rc); return ARTIC_ECUDA; }
  rc = tc::encode_seq_map(enc, &maps.y, Yb, p.y, p.N, p.G * p.Cog, pl.yrb / 2, pl.packed ? pl.seg_pitch : pl.kp, pl.yrb, 1);
  if (rc == CUDA_SUCCESS && x3)
    rc = tc::encode_seq_map(enc, &maps.y_lo, reinterpret_cast<const __nv_bfloat16*>(Yb) + p.y_plane, p.y, p.N, p.G * p.Cog,
                            pl.yrb / 2, pl.packed ? pl.seg_pitch : pl.kp, pl.yrb, 1);
  if (rc != CUDA_SUCCESS) { set_error("artic_tapconv_wgrad: cuTensorMapEncodeTiled(dY) failed (%d)", (int)rc); return ARTIC_ECUDA; }
  if (!x3) { maps.x_lo = maps.x; maps.y_lo = maps.y; }
  pl.dbg_flags = tc::g_debug[13];
  pl.dbg = tc::g_dbg_buf;
  pl.epi_transposed = tc::g_debug[18] == 1 ? 0 : 1;
  pl.trace = tc::g_trace_buf;
  pl.trace_cap = tc::g_trace_cap;
  pl.launch_id = tc::g_trace_buf != nullptr ? tc::g_trace_launch++ : 0;
  const int epi_need = 4 * 32 * 8 * 16 + 4 * 32 * 8;    // transpose stages + row pointers (overlaid on the operand stages)
  pl.ones_off = pl.n_stages * pl.stage_bytes > epi_need ? pl.n_stages * pl.stage_bytes : epi_need;
  const int smem_bytes = pl.ones_off + ones_bytes + 1024 + tc::WG_EPI_BYTES;
  if (x3) tc::tapwgrad_tc_kernel<true><<<(unsigned)grid, tc::WG_THREADS, smem_bytes, st>>>(p, pl, maps);
  else tc::tapwgrad_tc_kernel<false><<<(unsigned)grid, tc::WG_THREADS, smem_bytes, st>>>(p, pl, maps);
  ++g_path_counts[x3 ? PATH_WGRAD_TC_X3 : PATH_WGRAD_TC];
  *bias_fused = pl.bias_acc;
  if (pl.bias_acc) ++g_path_counts[8];
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) {
    set_error("artic_tapconv_wgrad(tc): launch failed: %s (grid %lld, smem %d of %d, bn %d acc %d stages %d packed %d)",
              cudaGetErrorString(le), (long long)grid, smem_bytes, tc::wg_max_smem(), pl.bn, pl.n_acc, pl.n_stages, pl.packed);
    return ARTIC_ECUDA;
  }
  return 1;
}
