// tcgen05 implicit-GEMM path of artic_tapconv for the dense bf16 contractions with unit
// input stride (every generator conv / transposed-conv phase, every data-gradient phase,
// the stride-1 discriminator convs).  sm_100a only.
//
//   D[m, co] (TMEM, fp32)  +=  A_t[m, ci] (smem, bf16, K-major)  x  B_t[co, ci] (smem, bf16, K-major)
//
// * M = 128-row sub-tiles of output positions, N = BN output channels, K = (tap, ci-chunk).
// * The activation tile is staged ONCE per ci-chunk with its halo (rows q+min_off ..
//   q+max_off) by TMA; every tap re-uses it through a row-shifted shared-memory matrix
//   descriptor (im2col-free).  Out-of-range rows are zero-filled by TMA, which IS the
//   convolution's zero padding (the row index is its own tensor-map dimension, so a tile
//   never bleeds into the neighbouring sequence).
// * Short sequences (discriminator tails, L <= 53) are packed: several zero-padded sequences
//   are laid end to end in the staged tile, so a 128-row MMA still does useful work.
// * Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
//   warps 2..9 = epilogue (TMEM -> registers -> fused bias / residual / LeakyReLU-mask /
//   activation -> global; two warps per TMEM lane quarter on alternating 32-channel chunks).
//   Persistent over tiles; the accumulator is double-buffered in TMEM when it fits, so the
//   epilogue of tile i overlaps the main loop of tile i+1.
// * Weight multicast (Plan.cs = 2 / 4): the CTAs of a thread-block cluster work on consecutive row tiles of the SAME
//   channel tile, so they stream identical weight stages.  Each CTA fetches 1/cs of every stage (a slice of the bn
//   output-channel rows) and TMA-multicasts it into all cs shared memories; a stage is refilled once all cs MMA warps
//   have released it (tcgen05.commit multicast onto every CTA's `w_empty`).  The deep layers (C >= 256) re-stream
//   their whole weight once per 128..512-row tile and are bound by that L2 -> SM traffic, which this divides by cs.
// * bf16x3 mode (template X3): fp32 activations and weights live in HBM next to SPLIT COPIES
//   (bf16 hi = rn(x), lo = rn(x - hi); artic_split / this kernel's epilogue).  Every K chunk is
//   issued three times — (x_hi, w_hi), (x_hi, w_lo), (x_lo, w_hi) — into the same fp32 TMEM
//   accumulator, i.e. the classic error-compensated product with ~2^-16 relative operand error
//   (the dropped x_lo * w_lo term is 2^-16 of the result as well).  The epilogue reads fp32
//   residual / mask operands and writes the fp32 result plus, on request, its split copy for
//   the next tensor-core consumer.
#include <mutex>

#include "tc_common.cuh"

namespace artic {
namespace tc {

constexpr int NTHREADS = 320;   // max: warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (4 or 8 of them)
constexpr int MAX_WS = 10;  // weight stages (streaming mode)
constexpr int MAX_AS = 8;   // activation stages
#ifndef ARTIC_TC_TRACE
#define ARTIC_TC_TRACE 0
#endif
constexpr int MAX_MT = 8;   // 128-row sub-tiles per CTA tile (narrow layers: bn <= 64 leaves TMEM room for 8)
constexpr int EPI_WARP_BYTES = MAX_MT * 32 * 8 + 8 * 32 * 4;   // per epilogue warp: row offsets of MAX_MT sub-tiles + bias of its (<= 8) channel chunks

struct Plan {
  int32_t w_early;    // 1: the first weight stages may be loaded before the grid-dependency wait
  int32_t kch;        // channels per K chunk (64 / 32 / 16)
  int32_t row_bytes;  // kch * 2 = swizzle span
  int32_t n_kc;       // ci chunks
  int32_t bn;         // output channels per tile
  int32_t n_nt;       // channel tiles per group
  int32_t mt;         // 128-row sub-tiles per CTA tile
  int32_t packed;     // 1: several short sequences per tile
  int32_t seg_per_tile, seg_pitch, seg_rows;  // packed: sequences per tile, smem row pitch, box rows
  int32_t tiles_per_seq;                      // plain: tiles per sequence
  int32_t boxr, nbox;                         // plain: box rows and boxes per activation stage
  int32_t n_mt;       // row tiles in total
  int32_t total_tiles;
  int32_t a_stage_bytes, w_stage_bytes, n_as, n_ws;
  int32_t w_tile_bytes;  // one tap's [bn x kch] weight tile (1024-byte multiple)
  int32_t tps;           // taps per weight stage: one barrier hand-shake feeds tps * mt * kch/16 MMAs
  int32_t w_resident;   // 1: the whole weight [n_kc][ntaps][bn x kch] stays in shared memory for the CTA's lifetime
  int32_t acc_stages, tmem_cols;
  int32_t min_off;
  int32_t n_ph, panel_bytes;      // input-stride phases (= si) and bytes of one phase panel
  int32_t shift[ARTIC_MAX_TAPS];  // (off[t] - min_off) / si : row shift inside the tap's phase panel
  int32_t phase[ARTIC_MAX_TAPS];  // (off[t] - min_off) % si : which phase panel the tap reads
  int32_t a_off16[ARTIC_MAX_TAPS];  // (phase * panel_bytes + shift * row_bytes) / 16 : descriptor offset of the tap
  int32_t layout_type;            // UMMA smem descriptor swizzle code
  int32_t n_kcl;                  // logical ci chunks (n_kc = 3 * n_kcl in the bf16x3 mode)
  int32_t ctas;                   // CTAs the planner wants for this problem (<= total_tiles: several tiles per CTA)
  int32_t cs;                     // cluster size (weight multicast); 1 = no cluster
  int32_t n_mg;                   // row-tile groups of cs consecutive tiles (n_mt / cs)
  long long* dbg;                 // debug timeline buffer (artic_debug_buffer) or nullptr
  int32_t dbg_flags;              // what-if timing (debug key 9): 1 = no epilogue loads / stores, 2 = no MMA issued
};

// Several independent problems (the phases of a strided data gradient / transposed conv, the three MRF
// blocks of a generator stage, the sub-discriminators at one depth) share ONE launch: the grid is the
// concatenation of per-problem persistent CTA ranges, each CTA works on its own problem only.
struct Prob {
  artic_tapconv_t p;
  Plan pl;
  CUtensorMap map_x, map_w;
  CUtensorMap map_x_lo, map_w_lo;   // bf16x3 mode: the `lo` planes
};
constexpr int MAXP = 8;
struct Multi {
  long long* trace;       // SM-occupancy trace buffer or nullptr
  long long trace_cap;
  int32_t launch_id;
  int32_t n;
  int32_t cta_begin[MAXP + 1];
  Prob prob[MAXP];
};

static_assert(sizeof(Multi) <= 32000, "kernel parameter space");

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout, sm_100):
// [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [49,52) base offset,
// [61,64) swizzle code.  Rows are `row_bytes` apart, 8-row groups SBO = 8*row_bytes apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t row_bytes, uint32_t layout_type, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;                                // LBO: unused for swizzled K-major layouts
  d |= (uint64_t)(((8u * row_bytes) >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)(layout_type & 7) << 61;
  return d;
}

// Debug timeline (artic_debug_buffer): CTA 0 records clock64 of the FIRST occurrence of each tag in a
// shared-memory slot (no atomics, ~20 clocks) and of the LAST occurrence in a second slot; dumped at exit.
__shared__ long long g_ev_first[40], g_ev_last[40];
__device__ __forceinline__ void dbg_mark(long long* dbg, int tag) {
  if (dbg != nullptr && blockIdx.x == 0) {
    const long long t = clock64();
    if (g_ev_first[tag] == 0) g_ev_first[tag] = t;
    g_ev_last[tag] = t;
  }
}

// ------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------
template <bool X3>
__global__ void __launch_bounds__(NTHREADS, 1)
tapconv_tc_kernel(const __grid_constant__ Multi mp) {
  int prob_j = 0;
  while (prob_j + 1 < mp.n && (int)blockIdx.x >= mp.cta_begin[prob_j + 1]) ++prob_j;
  const artic_tapconv_t& p = mp.prob[prob_j].p;
  const Plan& pl = mp.prob[prob_j].pl;
  const CUtensorMap& map_x = mp.prob[prob_j].map_x;
  const CUtensorMap& map_w = mp.prob[prob_j].map_w;
  const CUtensorMap& map_x_lo = mp.prob[prob_j].map_x_lo;
  const CUtensorMap& map_w_lo = mp.prob[prob_j].map_w_lo;
  // bf16x3: K chunk kc = 3 * kcl + part; part 0 = (x_hi, w_hi), 1 = (x_hi, w_lo), 2 = (x_lo, w_hi)
  const int n_wres = X3 ? 2 * pl.n_kcl : pl.n_kc;      // resident weight chunks: [w_hi chunks | w_lo chunks]
  const int cs = pl.cs;                                                   // cluster size (cs > 1: single-problem launch)
  const int crank = cs > 1 ? (int)cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
  // work units are "super tiles": (group, row-tile group of cs tiles, channel tile); CTA `crank` of a cluster takes row
  // tile mg * cs + crank of it.  cs == 1: a super tile is a tile, cta / ncta are this CTA's index / count in its problem.
  const int cta = ((int)blockIdx.x - mp.cta_begin[prob_j]) / cs;
  const int ncta = (mp.cta_begin[prob_j + 1] - mp.cta_begin[prob_j]) / cs;
  const int n_super = pl.total_tiles / cs;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[MAX_AS], a_empty[MAX_AS], w_full[MAX_WS], w_empty[MAX_WS];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2], w_res_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ long long tr_issue[64], tr_seen[64], tr_done[64];   // debug: per-W-load clocks (CTA 0)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1024-byte aligned operand staging area
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem0;
  const uint32_t w_base = smem0 + (uint32_t)pl.n_as * pl.a_stage_bytes;
  const uint32_t epi_base = w_base + (pl.w_resident ? (uint32_t)(n_wres * p.ntaps) * pl.w_tile_bytes : (uint32_t)pl.n_ws * pl.w_stage_bytes);   // 4 x 8 KB transpose stages + row offsets

  const long long t_start = clock64();
  const long long t_trace = (mp.trace != nullptr && threadIdx.x == 0) ? global_timer() : 0;
  if (pl.dbg_flags & 4) return;   // what-if timing (debug key 9 = 4): every CTA exits at once — the cost of the launches themselves
  pdl_launch_dependents();   // the next conv of the stream may start its prologue (and weight prefetch) under this one
  // Setup rendezvous on named barrier 1: the producer warp initialises the mbarriers, ARRIVES and goes
  // straight to its first TMA loads; the other warps (TMEM allocation in warp 1) SYNC on it.
  uint32_t tmem_base = 0;
  if (warp == 0) {
    if (pl.dbg != nullptr) {
      g_ev_first[lane] = g_ev_last[lane] = 0;
      if (lane < 8) g_ev_first[32 + lane] = g_ev_last[32 + lane] = 0;
      tr_issue[lane] = tr_issue[lane + 32] = tr_seen[lane] = tr_seen[lane + 32] = tr_done[lane] = tr_done[lane + 32] = 0;
    }
    if (lane == 0 && pl.dbg != nullptr && blockIdx.x == 0) g_ev_first[1] = g_ev_last[1] = t_start;
    if (lane == 0) {
      prefetch_tmap(&map_x);
      prefetch_tmap(&map_w);
      if (X3) { prefetch_tmap(&map_x_lo); prefetch_tmap(&map_w_lo); }
      for (int i = 0; i < pl.n_as; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
      for (int i = 0; i < pl.n_ws; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], (uint32_t)cs); }
      for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], blockDim.x - 64); }
      mbar_init(&w_res_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("bar.arrive 1, %0;" ::"r"(blockDim.x) : "memory");
  } else {
    if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)pl.tmem_cols);
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"r"(blockDim.x) : "memory");
    tc_fence_after();
    tmem_base = tmem_base_s;
  }
  if (cs > 1) {
    // no CTA may multicast into a peer (or arrive on its barriers) before that peer has initialised them.  (Warp 0 took
    // the early exit from the setup rendezvous: re-converge everything on the cluster barrier.)
    cluster_sync_all();
  }
  if (threadIdx.x == 32) dbg_mark(pl.dbg, 2);

  const int ntaps = p.ntaps;
  const int acc_cols = pl.mt * pl.bn;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      PipeState as(pl.n_as), ws(pl.n_ws);
#if ARTIC_TC_TRACE
      int n_wl = 0;
#endif
      const uint32_t w_bytes = (uint32_t)pl.bn * pl.row_bytes;
      int w_pre = 0;   // weight stages of the first tile issued ahead of the grid-dependency wait
      if (!pl.w_early) pdl_wait();   // the predecessor may have written the weights
      if (pl.w_resident) {
        // small layers (n_nt == 1, G == 1): every tile uses the same weights; load them once
        mbar_expect_tx(&w_res_full, (uint32_t)(n_wres * ntaps) * w_bytes);
        for (int kc = 0; kc < n_wres; ++kc) {
          const bool lo = X3 && kc >= pl.n_kcl;
          for (int t = 0; t < ntaps; ++t)
            tma_load_2d(w_base + (uint32_t)(kc * ntaps + t) * pl.w_tile_bytes, lo ? &map_w_lo : &map_w, &w_res_full,
                        (lo ? kc - pl.n_kcl : kc) * pl.kch, p.widx[t] * p.Cog);
        }
      } else if (cta < n_super && cs == 1) {
        // the weights are not written by the predecessor kernel (tc::note_weights_written): fill the
        // weight pipeline with the first tile's stages while the predecessor is still finishing
        const int nt = cta % pl.n_nt;
        const int g = (cta / pl.n_nt) / pl.n_mt;
        const int spk = (ntaps + pl.tps - 1) / pl.tps;
        const int npre = min(pl.n_ws, pl.n_kc * spk);
        for (; w_pre < npre; ++w_pre) {
          const int kc = w_pre / spk, t0 = (w_pre % spk) * pl.tps;
          const int kcl = X3 ? kc / 3 : kc;
          const CUtensorMap* mw = (X3 && kc % 3 == 1) ? &map_w_lo : &map_w;
          const int nt_g = min(pl.tps, ntaps - t0);
          mbar_expect_tx(&w_full[ws.stage], (uint32_t)nt_g * w_bytes);
          for (int j = 0; j < nt_g; ++j)
            tma_load_2d(w_base + (uint32_t)ws.stage * pl.w_stage_bytes + (uint32_t)j * pl.w_tile_bytes, mw,
                        &w_full[ws.stage], kcl * pl.kch, (p.widx[t0 + j] * p.G + g) * p.Cog + nt * pl.bn);
          ws.next();
        }
      }
      pdl_wait();   // activations (and everything the epilogue touches) come from the predecessor
      const int wsl_rows = pl.bn / cs;                                   // weight rows this CTA fetches per tap (multicast slice)
      const uint32_t wsl_off = (uint32_t)(crank * wsl_rows) * pl.row_bytes;
      for (int tile = cta; tile < n_super; tile += ncta) {
        const int nt = tile % pl.n_nt;
        const int r = tile / pl.n_nt;
        const int mtile = (r % pl.n_mg) * cs + crank;
        const int g = r / pl.n_mg;
        for (int kc = 0; kc < pl.n_kc; ++kc) {
          const int kcl = X3 ? kc / 3 : kc;
          const CUtensorMap* mx = (X3 && kc % 3 == 2) ? &map_x_lo : &map_x;
          const CUtensorMap* mw = (X3 && kc % 3 == 1) ? &map_w_lo : &map_w;
          const int c0 = g * p.Cig + kcl * pl.kch;
          // bf16x3: pass 1 (x_hi, w_lo) re-uses the x_hi tile staged for pass 0 — no second load, the stage is
          // released after pass 1
          const bool a_reuse = X3 && kc % 3 == 1;
          if (!a_reuse) {
          mbar_wait(&a_empty[as.stage], as.phase ^ 1);
          const uint32_t a_dst = a_base + (uint32_t)as.stage * pl.a_stage_bytes;
          if (!pl.packed) {
            const int n = mtile / pl.tiles_per_seq;
            const int qt = mtile % pl.tiles_per_seq;
            const int qrow0 = p.q0 + qt * pl.mt * 128;
            mbar_expect_tx(&a_full[as.stage], (uint32_t)pl.n_ph * pl.nbox * pl.boxr * pl.row_bytes);
            for (int ph = 0; ph < pl.n_ph; ++ph)
              for (int b = 0; b < pl.nbox; ++b)
                tma_load_4d(a_dst + (uint32_t)ph * pl.panel_bytes + (uint32_t)b * pl.boxr * pl.row_bytes, mx,
                            &a_full[as.stage], c0, n % p.x.n_inner, (qrow0 + b * pl.boxr) * p.si + pl.min_off + ph,
                            n / p.x.n_inner);
          } else {
            const int n0 = mtile * pl.seg_per_tile;
            const int nseg = min(pl.seg_per_tile, p.N - n0);
            mbar_expect_tx(&a_full[as.stage], (uint32_t)nseg * pl.n_ph * pl.seg_rows * pl.row_bytes);
            for (int j = 0; j < nseg; ++j) {
              const int n = n0 + j;
              for (int ph = 0; ph < pl.n_ph; ++ph)
                tma_load_4d(a_dst + (uint32_t)ph * pl.panel_bytes + (uint32_t)j * pl.seg_pitch * pl.row_bytes, mx,
                            &a_full[as.stage], c0, n % p.x.n_inner, p.q0 * p.si + pl.min_off + ph, n / p.x.n_inner);
            }
          }
          as.next();
          }
          if (pl.w_resident) continue;
          for (int t0 = 0; t0 < ntaps; t0 += pl.tps) {
            if (w_pre > 0) { --w_pre; continue; }
            const int nt_g = min(pl.tps, ntaps - t0);
            mbar_wait(&w_empty[ws.stage], ws.phase ^ 1);
#if ARTIC_TC_TRACE
            if (pl.dbg != nullptr && n_wl < 64) tr_issue[n_wl++] = clock64();
#endif
            mbar_expect_tx(&w_full[ws.stage], (uint32_t)nt_g * w_bytes);
            if (cs == 1) {
              for (int j = 0; j < nt_g; ++j)
                tma_load_2d(w_base + (uint32_t)ws.stage * pl.w_stage_bytes + (uint32_t)j * pl.w_tile_bytes, mw,
                            &w_full[ws.stage], kcl * pl.kch, (p.widx[t0 + j] * p.G + g) * p.Cog + nt * pl.bn);
            } else {
              for (int j = 0; j < nt_g; ++j)     // this CTA's row slice of every tap tile, into all cs CTAs
                tma_load_2d_mc(w_base + (uint32_t)ws.stage * pl.w_stage_bytes + (uint32_t)j * pl.w_tile_bytes + wsl_off, mw,
                               &w_full[ws.stage], kcl * pl.kch,
                               (p.widx[t0 + j] * p.G + g) * p.Cog + nt * pl.bn + crank * wsl_rows, cmask);
            }
            ws.next();
          }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer =================================
    // One thread issues every MMA, so the instructions between two tcgen05.mma are the kernel's
    // critical path for narrow tiles: the shared-memory descriptors are reduced to one 32-bit add
    // per operand (constant high word; the low word is the 16-byte address, to which the tap's
    // precomputed row shift, the sub-tile offset and the k-step are added).
    {
      const bool leader = elect_one();
      const bool skip_mma = (pl.dbg_flags & 2) != 0;      // what-if timing (debug key 9): the pipeline runs, no MMA is issued
      PipeState as(pl.n_as), ws(pl.n_ws), acc(pl.acc_stages);
#if ARTIC_TC_TRACE
      int n_ws_seen = 0;
#endif
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = bn, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(pl.bn >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t desc_hi = (((8u * (uint32_t)pl.row_bytes) >> 4) & 0x3fffu) | (1u << 14) | ((uint32_t)(pl.layout_type & 7) << 29);
      const uint32_t desc_lo = 1u << 16;                      // LBO field (unused for swizzled K-major)
      const uint32_t m_step16 = (128u * (uint32_t)pl.row_bytes) >> 4;
      const int ksteps = pl.kch / 16;
      const int mt = pl.mt, tps = pl.tps;
      const bool w_res = pl.w_resident != 0;
      const uint32_t bn = (uint32_t)pl.bn;
      const uint32_t w_base16 = w_base >> 4, wt16 = (uint32_t)pl.w_tile_bytes >> 4, ws16 = (uint32_t)pl.w_stage_bytes >> 4;
      if (pl.w_resident) mbar_wait(&w_res_full, 0);
      for (int tile = cta; tile < n_super; tile += ncta) {
        mbar_wait(&acc_empty[acc.stage], acc.phase ^ 1);
        tc_fence_after();
        const uint32_t d_base = tmem_base + (uint32_t)acc.stage * acc_cols;
        uint32_t accum = 0;
        bool first_tap = true;
        for (int kc = 0; kc < pl.n_kc; ++kc) {
          if (!(X3 && kc % 3 == 1)) mbar_wait(&a_full[as.stage], as.phase);     // bf16x3 pass 1: same x_hi stage as pass 0
          if (lane == 0) dbg_mark(pl.dbg, 20);
          const uint32_t a16 = desc_lo | (((a_base + (uint32_t)as.stage * pl.a_stage_bytes) >> 4) & 0x3fffu);
          for (int t0 = 0; t0 < ntaps; t0 += tps) {
            const int nt_g = min(tps, ntaps - t0);
            uint32_t w16;
            if (w_res) {
              const int wkc = X3 ? (kc % 3 == 1 ? pl.n_kcl + kc / 3 : kc / 3) : kc;
              w16 = w_base16 + (uint32_t)(wkc * ntaps + t0) * wt16;
            } else {
              mbar_wait(&w_full[ws.stage], ws.phase);
              w16 = w_base16 + (uint32_t)ws.stage * ws16;
            }
#if ARTIC_TC_TRACE
            if (lane == 0 && pl.dbg != nullptr && n_ws_seen < 64) tr_seen[n_ws_seen] = clock64();
#endif
            tc_fence_after();
            for (int j = 0; j < nt_g; ++j) {
              const uint32_t b16 = desc_lo | ((w16 + (uint32_t)j * wt16) & 0x3fffu);
              uint32_t at16 = a16 + (uint32_t)pl.a_off16[t0 + j];
              uint32_t dcol = d_base;
              for (int m = 0; m < mt; ++m) {
                if (ksteps == 4) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    if (leader && !skip_mma) umma_bf16(dcol, ((uint64_t)desc_hi << 32) | (at16 + 2 * k), ((uint64_t)desc_hi << 32) | (b16 + 2 * k), idesc, accum);
                    accum = 1;
                  }
                } else if (ksteps == 2) {
#pragma unroll
                  for (int k = 0; k < 2; ++k) {
                    if (leader && !skip_mma) umma_bf16(dcol, ((uint64_t)desc_hi << 32) | (at16 + 2 * k), ((uint64_t)desc_hi << 32) | (b16 + 2 * k), idesc, accum);
                    accum = 1;
                  }
                } else {
                  if (leader && !skip_mma) umma_bf16(dcol, ((uint64_t)desc_hi << 32) | at16, ((uint64_t)desc_hi << 32) | b16, idesc, accum);
                  accum = 1;
                }
                // the accumulate flag must stay 0 for the first k-step of EVERY sub-tile of the tile
                if (first_tap && m + 1 < mt) accum = 0;
                at16 += m_step16;
                dcol += bn;
              }
              first_tap = false;
            }
#if ARTIC_TC_TRACE
            if (pl.dbg != nullptr && n_ws_seen < 64) { if (lane == 0) tr_done[n_ws_seen] = clock64(); ++n_ws_seen; }
#endif
            if (!w_res) {
              if (leader) {
                if (cs == 1) umma_commit(&w_empty[ws.stage]);
                else umma_commit_mc(&w_empty[ws.stage], cmask);      // the stage is shared: release it in every CTA
              }
              ws.next();
            }
          }
          if (!(X3 && kc % 3 == 0)) {       // bf16x3: the x_hi stage stays for pass 1
            if (leader) umma_commit(&a_empty[as.stage]);
            as.next();
          }
        }
        if (leader) umma_commit(&acc_full[acc.stage]);
        if (lane == 0) dbg_mark(pl.dbg, 22);
        acc.next();
      }
    }
  } else {
    // =============================== epilogue ===================================
    // Eight warps work on a tile: two per TMEM lane quarter, on alternating 32-channel chunks.  Everything
    // that does not depend on the accumulator — output row offsets, the bias, the residual / mask operands
    // of the chunk — is fetched BEFORE the accumulator is waited for / read.
    pdl_wait();
    const int ew = warp & 3;            // TMEM lane quarter this warp may access
    const int ewarp = warp - 2;         // 0..7
    const int eh = ewarp >> 2;          // channel-chunk parity
    PipeState acc(pl.acc_stages);
    using TO = __nv_bfloat16;
    const TO* __restrict__ res_pre = reinterpret_cast<const TO*>(p.res_pre);
    const TO* __restrict__ mask = reinterpret_cast<const TO*>(p.mask);
    const TO* __restrict__ res = reinterpret_cast<const TO*>(p.res);
    TO* __restrict__ Y = reinterpret_cast<TO*>(p.Y);
    TO* __restrict__ Y2 = reinterpret_cast<TO*>(p.Y2);
    uint8_t* epi = smem_raw + (epi_base - smem_u32(smem_raw));
    const int n_ew = (int)(blockDim.x >> 5) - 2;            // 4 or 8 epilogue warps
    long long* rowoff = reinterpret_cast<long long*>(epi + ewarp * EPI_WARP_BYTES);          // [sub-tile][32 rows]
    float* bias_s = reinterpret_cast<float*>(epi + ewarp * EPI_WARP_BYTES + MAX_MT * 32 * 8);  // [chunk][32]
    const float neg_slope = p.act == ARTIC_ACT_LRELU ? p.act_slope : 1.f;
    for (int tile = cta; tile < n_super; tile += ncta) {
      const int nt = tile % pl.n_nt;
      const int r = tile / pl.n_nt;
      const int mtile = (r % pl.n_mg) * cs + crank;
      const int g = r / pl.n_mg;
      const int cbase = g * p.Cog + nt * pl.bn;
      for (int m = 0; m < pl.mt; ++m) {   // output offset of this lane's row in every sub-tile (-1: not stored)
        const int mrow = m * 128 + ew * 32 + lane;
        int n, q;
        bool valid;
        if (!pl.packed) {
          n = mtile / pl.tiles_per_seq;
          q = (mtile % pl.tiles_per_seq) * pl.mt * 128 + mrow;
          valid = q < p.nq;
        } else {
          const int j = mrow / pl.seg_pitch;
          q = mrow - j * pl.seg_pitch;
          n = mtile * pl.seg_per_tile + j;
          valid = j < pl.seg_per_tile && q < p.nq && n < p.N;
        }
        long long o = -1;
        if (valid) {
          const int row = (p.q0 + q) * p.so + p.ro;
          if (row >= 0 && row < p.y.len) o = seq_base(p.y, n) + (int64_t)row * p.y.s_row + cbase;
        }
        rowoff[m * 32 + lane] = (pl.dbg_flags & 1) ? -1 : o;      // what-if timing (debug key 9 = 1): no epilogue loads / stores
      }
      // bn == 32 (one channel chunk): the two warp quartets take ALTERNATE sub-tiles instead of alternate chunks (the
      // second quartet would idle otherwise)
      const bool alt_m = pl.bn == 32 && n_ew == 8;
      const int c_first = alt_m ? 0 : eh * 32;
      {  // bias of this warp's channel chunks (same for every sub-tile): loaded before the accumulator is ready
        int j = 0;
        for (int c0 = c_first; c0 < pl.bn; c0 += 8 * n_ew, ++j)
          bias_s[j * 32 + lane] = p.bias != nullptr ? __ldg(p.bias + cbase + c0 + lane) : 0.f;
      }
      __syncwarp();
      bool waited = false;
      for (int m = 0; m < pl.mt; ++m) {
        if (alt_m && (m & 1) != eh) continue;
        const uint32_t t_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)acc.stage * acc_cols + (uint32_t)m * pl.bn;
        int cj = 0;
        for (int c0 = c_first; c0 < pl.bn; c0 += 8 * n_ew, ++cj) {
          // Fused epilogue in the TMEM row-per-lane layout: lane = output row, 32 channels per chunk as
          // four 16-byte pieces.  (A shared-memory transposed variant with lanes along the channels
          // coalesces better but costs ~5x the instructions; the epilogue of these small tiles is
          // issue-latency bound, not bandwidth bound: measured 5.6k vs 11.7k clocks per 256x128 tile.)
          const long long o = rowoff[m * 32 + lane];
          if constexpr (X3) {
            // bf16x3 mode: fp32 epilogue operands / outputs, optional split copies of the outputs
            const float* __restrict__ res_pre32 = reinterpret_cast<const float*>(p.res_pre);
            const float* __restrict__ mask32 = reinterpret_cast<const float*>(p.mask);
            const float* __restrict__ res32 = reinterpret_cast<const float*>(p.res);
            float* __restrict__ Y32 = reinterpret_cast<float*>(p.Y);
            float* __restrict__ Y2_32 = reinterpret_cast<float*>(p.Y2);
            TO* __restrict__ Ysp = reinterpret_cast<TO*>(p.Y_sp);
            TO* __restrict__ Y2sp = reinterpret_cast<TO*>(p.Y2_sp);
            if (!waited) {
              mbar_wait(&acc_full[acc.stage], acc.phase);
              tc_fence_after();
              waited = true;
            }
            uint32_t acc_r[32];
            tmem_ld32(t_row + c0, acc_r);
            tmem_ld_wait();
            if (o >= 0) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float4 b0 = reinterpret_cast<const float4*>(bias_s + cj * 32)[2 * u], b1 = reinterpret_cast<const float4*>(bias_s + cj * 32)[2 * u + 1];
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                const long long oc = o + c0 + 8 * u;
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaf(p.alpha, __uint_as_float(acc_r[8 * u + i]), bb[i]);
                if (res_pre32) {
                  const float4 t0 = __ldg(reinterpret_cast<const float4*>(res_pre32 + oc)), t1 = __ldg(reinterpret_cast<const float4*>(res_pre32 + oc + 4));
                  v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w; v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
                }
                if (mask32) {
                  const float4 t0 = __ldg(reinterpret_cast<const float4*>(mask32 + oc)), t1 = __ldg(reinterpret_cast<const float4*>(mask32 + oc + 4));
                  const float mk[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] *= (mk[i] > 0.f ? 1.f : p.mask_slope);
                }
                if (res32) {
                  const float4 t0 = __ldg(reinterpret_cast<const float4*>(res32 + oc)), t1 = __ldg(reinterpret_cast<const float4*>(res32 + oc + 4));
                  v[0] += t0.x; v[1] += t0.y; v[2] += t0.z; v[3] += t0.w; v[4] += t1.x; v[5] += t1.y; v[6] += t1.z; v[7] += t1.w;
                }
                if (Y32) {
                  *reinterpret_cast<float4*>(Y32 + oc) = make_float4(v[0], v[1], v[2], v[3]);
                  *reinterpret_cast<float4*>(Y32 + oc + 4) = make_float4(v[4], v[5], v[6], v[7]);
                }
                if (Ysp) {
                  const uint4 h = pack8(v);
                  float hv[8], lv[8];
                  unpack8<TO>(h, hv);
#pragma unroll
                  for (int i = 0; i < 8; ++i) lv[i] = v[i] - hv[i];
                  *reinterpret_cast<uint4*>(Ysp + oc) = h;
                  *reinterpret_cast<uint4*>(Ysp + p.y_plane + oc) = pack8(lv);
                }
                if (Y2_32 || Y2sp) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = v[i] > 0.f ? v[i] : neg_slope * v[i];
                  if (Y2_32) {
                    *reinterpret_cast<float4*>(Y2_32 + oc) = make_float4(v[0], v[1], v[2], v[3]);
                    *reinterpret_cast<float4*>(Y2_32 + oc + 4) = make_float4(v[4], v[5], v[6], v[7]);
                  }
                  if (Y2sp) {
                    const uint4 h = pack8(v);
                    float hv[8], lv[8];
                    unpack8<TO>(h, hv);
#pragma unroll
                    for (int i = 0; i < 8; ++i) lv[i] = v[i] - hv[i];
                    *reinterpret_cast<uint4*>(Y2sp + oc) = h;
                    *reinterpret_cast<uint4*>(Y2sp + p.y_plane + oc) = pack8(lv);
                  }
                }
              }
            }
          } else {
            // ---- (1) everything independent of the accumulator: operands of this lane's row, bias of the chunk
            uint4 q_rp[4], q_mk[4], q_rs[4];
  #pragma unroll
            for (int u = 0; u < 4; ++u) {
              q_rp[u] = q_mk[u] = q_rs[u] = make_uint4(0, 0, 0, 0);
              if (o >= 0) {
                if (res_pre) q_rp[u] = __ldg(reinterpret_cast<const uint4*>(res_pre + o + c0 + 8 * u));
                if (mask) q_mk[u] = __ldg(reinterpret_cast<const uint4*>(mask + o + c0 + 8 * u));
                if (res) q_rs[u] = __ldg(reinterpret_cast<const uint4*>(res + o + c0 + 8 * u));
              }
            }
            if (!waited) {
              mbar_wait(&acc_full[acc.stage], acc.phase);
              if (threadIdx.x == 64) dbg_mark(pl.dbg, 30);
              tc_fence_after();
              waited = true;
            }
            // ---- (2) accumulator chunk
            uint32_t acc_r[32];
            tmem_ld32(t_row + c0, acc_r);
            tmem_ld_wait();
            // ---- (3) bias / residual / mask / activation, stores
            if (o >= 0) {
  #pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float4 b0 = reinterpret_cast<const float4*>(bias_s + cj * 32)[2 * u], b1 = reinterpret_cast<const float4*>(bias_s + cj * 32)[2 * u + 1];
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float v[8], tmp[8];
  #pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = fmaf(p.alpha, __uint_as_float(acc_r[8 * u + i]), bb[i]);
                if (res_pre) {
                  unpack8<TO>(q_rp[u], tmp);
  #pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] += tmp[i];
                }
                if (mask) {
                  unpack8<TO>(q_mk[u], tmp);
  #pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] *= (tmp[i] > 0.f ? 1.f : p.mask_slope);
                }
                if (res) {
                  unpack8<TO>(q_rs[u], tmp);
  #pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] += tmp[i];
                }
                if (Y) *reinterpret_cast<uint4*>(Y + o + c0 + 8 * u) = pack8(v);
                if (Y2) {   // second output: LeakyReLU (or identity: neg_slope = 1); tanh layers never reach this kernel
  #pragma unroll
                  for (int i = 0; i < 8; ++i) v[i] = v[i] > 0.f ? v[i] : neg_slope * v[i];
                  *reinterpret_cast<uint4*>(Y2 + o + c0 + 8 * u) = pack8(v);
                }
              }
            }
          }
        }
      }
      if (!waited) {   // bn == 32: the odd-chunk warps have no channels, but still own a share of the barrier
        mbar_wait(&acc_full[acc.stage], acc.phase);
        tc_fence_after();
      }
      tc_fence_before();
      if (threadIdx.x == 64) dbg_mark(pl.dbg, 31);
      mbar_arrive(&acc_empty[acc.stage]);
      acc.next();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();     // peers may still multicast into / arrive on this CTA's shared memory
  if (pl.dbg != nullptr && blockIdx.x == 0 && threadIdx.x < 40) {
    pl.dbg[1 + 2 * threadIdx.x] = g_ev_first[threadIdx.x];
    pl.dbg[2 + 2 * threadIdx.x] = g_ev_last[threadIdx.x];
    if (threadIdx.x == 0) { pl.dbg[0] = 40; pl.dbg[90] = clock64(); }
  }
  if (pl.dbg != nullptr && blockIdx.x == 0 && threadIdx.x < 64) {   // debug trace: slots [3000 + 3*i ..]
    pl.dbg[3000 + 3 * threadIdx.x] = tr_issue[threadIdx.x];
    pl.dbg[3001 + 3 * threadIdx.x] = tr_seen[threadIdx.x];
    pl.dbg[3002 + 3 * threadIdx.x] = tr_done[threadIdx.x];
  }
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)pl.tmem_cols);
  }
  if (mp.trace != nullptr && threadIdx.x == 0) trace_cta(mp.trace, mp.trace_cap, mp.launch_id, prob_j & 7, t_trace);
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

int g_debug[32] = {0};
long long* g_dbg_buf = nullptr;
long long* g_trace_buf = nullptr;
long long g_trace_cap = 0;
int g_trace_launch = 0;
static int g_smem_optin = 0;

// Largest dynamic shared-memory size the kernel may be launched with (opt-in limit minus the
// kernel's static shared memory); sets the function attribute once.
static int max_smem() {
  if (g_smem_optin == 0) {
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (optin <= 0) optin = 227 * 1024;
    cudaFuncAttributes fa;
    int stat = 2048;
    if (cudaFuncGetAttributes(&fa, tapconv_tc_kernel<false>) == cudaSuccess) stat = (int)fa.sharedSizeBytes;
    if (cudaFuncGetAttributes(&fa, tapconv_tc_kernel<true>) == cudaSuccess && (int)fa.sharedSizeBytes > stat) stat = (int)fa.sharedSizeBytes;
    int dyn = optin - stat;
    if (cudaFuncSetAttribute(tapconv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess ||
        cudaFuncSetAttribute(tapconv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn) != cudaSuccess) {
      cudaGetLastError();
      dyn = 48 * 1024;
    }
    g_smem_optin = dyn;
  }
  return g_smem_optin;
}

// Streams whose prepared weights were rewritten since their last tensor-core conv launch.
static std::mutex g_wfence_mu;
static cudaStream_t g_wfence[256];
static int g_wfence_n = 0;
static bool g_wfence_overflow = false;   // list full once: early weight prefetch stays off for good
void note_weights_written(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_wfence_mu);
  for (int i = 0; i < g_wfence_n; ++i)
    if (g_wfence[i] == st) return;
  if (g_wfence_n < 256) g_wfence[g_wfence_n++] = st;
  else g_wfence_overflow = true;
}
static bool consume_weight_fence(cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_wfence_mu);
  if (g_wfence_overflow) return true;
  for (int i = 0; i < g_wfence_n; ++i)
    if (g_wfence[i] == st) { g_wfence[i] = g_wfence[--g_wfence_n]; return true; }
  return false;
}

}  // namespace tc
}  // namespace artic

using namespace artic;

extern "C" int artic_debug_buffer(void* dev_buf) {
  tc::g_dbg_buf = reinterpret_cast<long long*>(dev_buf);
  return ARTIC_OK;
}

extern "C" int artic_trace_buffer(void* dev_buf, long long capacity_records) {
  tc::g_trace_buf = reinterpret_cast<long long*>(dev_buf);
  tc::g_trace_cap = dev_buf != nullptr ? capacity_records : 0;
  tc::g_trace_launch = 0;
  return ARTIC_OK;
}

extern "C" int artic_debug_set(int key, int value) {
  if (key < 0 || key >= 32) return ARTIC_EINVAL;
  tc::g_debug[key] = value;
  return ARTIC_OK;
}

extern "C" int artic_debug_get(int key) { return (key < 0 || key >= 32) ? 0 : tc::g_debug[key]; }

static inline int w_all_bytes(const tc::Plan& pl, int ntaps, bool x3) { return (x3 ? 2 * pl.n_kcl : pl.n_kc) * ntaps * pl.w_tile_bytes; }

// Plans one problem for the tensor-core kernel: returns 1 (pr filled: parameters, plan, tensor maps,
// pr_smem / pr_cost set), 0 if the shape is not eligible, <0 on error.
// n_share = number of problems that will share the grid (each gets ~1/n_share of the SMs).
static int tc_plan_problem(const artic_tapconv_t* pp, tc::Prob& pr, int& pr_smem, double& pr_cost, int n_share,
                           bool allow_cluster) {
  const artic_tapconv_t& p = *pp;
  if (tc::g_debug[1]) return 0;                       // debug: force the generic kernel
  const bool x3 = p.dtype == ARTIC_F32 && p.out_dtype == ARTIC_F32 && p.X_sp != nullptr && p.Wt_sp != nullptr;
  if (!x3 && (p.Wt == nullptr || p.dtype != ARTIC_BF16 || p.out_dtype != ARTIC_BF16)) return 0;
  if (x3 && tc::g_debug[24] == 1) return 0;            // debug key 24: bf16x3 problems on the CUDA-core kernel
  if (p.act == ARTIC_ACT_TANH || p.res2 != nullptr) return 0;
  if (p.si < 1 || p.si > 8) return 0;
  if (p.Cig % 16 != 0 || p.Cog % 32 != 0) return 0;
  if ((p.x.s_row % 8) || (p.x.s_outer % 8) || (p.x.n_inner > 1 && (p.x.s_inner % 8))) return 0;
  if ((p.y.s_row % 8) || (p.y.s_outer % 8) || (p.y.n_inner > 1 && (p.y.s_inner % 8))) return 0;
  const void* ptrs[] = {x3 ? p.X_sp : p.X, x3 ? p.Wt_sp : p.Wt, p.res_pre, p.mask, p.res, p.res2, p.Y, p.Y2, x3 ? p.Y_sp : nullptr,
                        x3 ? p.Y2_sp : nullptr};
  for (const void* q : ptrs)
    if (q != nullptr && (reinterpret_cast<uintptr_t>(q) & 15)) return 0;
  if (x3 && ((p.x_plane % 8) || (p.w_plane % 8) || (p.y_plane % 8))) return 0;
  tc::EncodeTiledFn enc = tc::encode_fn();
  if (enc == nullptr) return 0;

  // "compact" CTAs (debug key 8 = shared-memory cap in KB, e.g. 110): <= half of the SM's shared memory,
  // <= 256 TMEM columns and 4 epilogue warps, so that TWO CTAs (usually of different kernels running on
  // concurrent streams) share an SM and one's prologue / epilogue tail overlaps the other's main loop.
  const int smem_cap = tc::g_debug[8] > 0 ? tc::g_debug[8] * 1024 : tc::max_smem();
  const bool compact = smem_cap < tc::max_smem();
  const int n_ew = compact ? 4 : 8;
  const int tmem_cap = compact ? 256 : 512;
  const int epi_bytes = n_ew * tc::EPI_WARP_BYTES;
  const int budget = (smem_cap < tc::max_smem() ? smem_cap : tc::max_smem()) - 1024 /*alignment slack*/ - epi_bytes;
  auto make_plan = [&](tc::Plan& pl, int bn_req, int mt_req) -> bool {
    memset(&pl, 0, sizeof(pl));
  int min_off = p.off[0], max_off = p.off[0];
  for (int t = 1; t < p.ntaps; ++t) { min_off = min(min_off, p.off[t]); max_off = max(max_off, p.off[t]); }
  if (max_off - min_off > 160 * p.si || min_off < -(1 << 20)) return false;   // also rejects the "no tap on this phase" marker
  pl.min_off = min_off;
  pl.n_ph = p.si;
  int span = 0;                                       // largest row shift inside a phase panel
  for (int t = 0; t < p.ntaps; ++t) {
    pl.shift[t] = (p.off[t] - min_off) / p.si;
    pl.phase[t] = (p.off[t] - min_off) % p.si;
    span = max(span, pl.shift[t]);
  }
  pl.kch = (p.Cig % 64 == 0) ? 64 : (p.Cig % 32 == 0) ? 32 : 16;
  pl.row_bytes = pl.kch * 2;
  pl.n_kcl = p.Cig / pl.kch;
  pl.n_kc = x3 ? 3 * pl.n_kcl : pl.n_kcl;
  pl.layout_type = pl.row_bytes == 128 ? 2 : pl.row_bytes == 64 ? 4 : 6;
  pl.bn = bn_req;
  pl.n_nt = p.Cog / pl.bn;
  int mt = tmem_cap / pl.bn;   // up to the whole TMEM budget (single-buffered accumulator when > half of it)
  if (mt < 1) mt = 1;
  // up to 4 sub-tiles; narrow layers (bn <= 64: HBM / latency bound, the CTA count is what they cost) up to MAX_MT when
  // debug key 26 asks for it
  const int mt_cap = (pl.bn <= 64 && tc::g_debug[26] > 4) ? (tc::g_debug[26] < tc::MAX_MT ? tc::g_debug[26] : tc::MAX_MT) : 4;
  if (mt > mt_cap) mt = mt_cap;
  if (mt_req > 0 && mt_req < mt) mt = mt_req;
  pl.w_tile_bytes = ((pl.bn * pl.row_bytes + 1023) / 1024) * 1024;
  pl.tps = 32 * 1024 / pl.w_tile_bytes;                  // stages of <= 32 KB
  if (pl.tps > 4) pl.tps = 4;
  if (pl.tps > p.ntaps) pl.tps = p.ntaps;
  if (pl.tps < 1) pl.tps = 1;
  if (tc::g_debug[0] > 0) pl.tps = tc::g_debug[0] < p.ntaps ? tc::g_debug[0] : p.ntaps;   // debug: force taps per stage
  pl.w_stage_bytes = pl.tps * pl.w_tile_bytes;
  {  // keep two activation stages + a few weight stages within shared memory
    auto a_bytes = [&](int m) { return p.si * ((((m * 128 + span + 64) * pl.row_bytes + 1023) / 1024) * 1024); };
    while (mt > 1 && 2 * a_bytes(mt) + 3 * pl.w_stage_bytes > budget) --mt;
  }
  const int rows_align = 128 / pl.row_bytes;          // TMA shared-memory destinations are 128-byte aligned
  const int lpad = p.nq + span;
  pl.packed = (p.N >= 2 && lpad * p.si <= 256 && 2 * lpad <= mt * 128) ? 1 : 0;
  if (tc::g_debug[4] == 1) pl.packed = 0;
  int a_rows;
  if (pl.packed) {
    pl.seg_rows = lpad;
    pl.seg_pitch = ((lpad + rows_align - 1) / rows_align) * rows_align;
    pl.seg_per_tile = (mt * 128) / pl.seg_pitch;
    if (pl.seg_per_tile > p.N) {
      pl.seg_per_tile = p.N;
      mt = (pl.seg_per_tile * pl.seg_pitch + 127) / 128;
    }
    pl.n_mt = (p.N + pl.seg_per_tile - 1) / pl.seg_per_tile;
    a_rows = mt * 128 + span + 8;
  } else {
    while (mt > 1 && (mt - 1) * 128 >= p.nq) --mt;
    pl.tiles_per_seq = (p.nq + mt * 128 - 1) / (mt * 128);
    pl.n_mt = p.N * pl.tiles_per_seq;
    pl.boxr = p.si == 1 ? 64 : 32;                     // TMA box extent (rows * si) must stay <= 256
    pl.nbox = (mt * 128 + span + pl.boxr - 1) / pl.boxr;
    a_rows = pl.nbox * pl.boxr;
  }
  pl.mt = mt;
  pl.acc_stages = (2 * mt * pl.bn <= tmem_cap) ? 2 : 1;
  int cols = pl.acc_stages * mt * pl.bn;
  pl.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  if (cols > tmem_cap) return false;
  pl.panel_bytes = ((a_rows * pl.row_bytes + 1023) / 1024) * 1024;
  pl.a_stage_bytes = pl.n_ph * pl.panel_bytes;
  // Shared memory: weights either RESIDENT (single channel tile, single group, <= 96 KB: the C = 32 / 64
  // generator stages — their per-tile weight traffic would otherwise serialise the producer) or
  // streamed through up to MAX_WS stages; the rest goes to activation stages, because the bytes in
  // flight per SM (x ~2 us TMA latency) are what bounds the operand bandwidth.
  const int w_all = (x3 ? 2 * pl.n_kcl : pl.n_kc) * p.ntaps * pl.w_tile_bytes;
  pl.w_resident = (pl.n_nt == 1 && p.G == 1 && w_all <= 96 * 1024 && budget - w_all >= 2 * pl.a_stage_bytes &&
                   tc::g_debug[7] != 1) ? 1 : 0;
  if (pl.w_resident) {
    pl.n_ws = 1;   // unused
    pl.n_as = (budget - w_all) / pl.a_stage_bytes;
    if (pl.n_as > tc::MAX_AS) pl.n_as = tc::MAX_AS;
  } else {
    const int steps = pl.n_kc * ((p.ntaps + pl.tps - 1) / pl.tps);
    int want_ws = steps < tc::MAX_WS ? steps : tc::MAX_WS;
    if (want_ws < 2) want_ws = 2;
    pl.n_as = 2;
    if (budget - pl.n_as * pl.a_stage_bytes < 2 * pl.w_stage_bytes) {
      pl.n_as = 1;
      if (budget - pl.a_stage_bytes < 2 * pl.w_stage_bytes) return false;
    }
    int rem = budget - pl.n_as * pl.a_stage_bytes;
    pl.n_ws = rem / pl.w_stage_bytes;
    if (pl.n_ws > want_ws) pl.n_ws = want_ws;
    rem -= pl.n_ws * pl.w_stage_bytes;
    while (pl.n_as < tc::MAX_AS && pl.n_as < 2 * pl.n_kc + 1 && rem >= pl.a_stage_bytes) { ++pl.n_as; rem -= pl.a_stage_bytes; }
    int more = rem / pl.w_stage_bytes;     // leftover: deepen the weight pipeline
    while (more-- > 0 && pl.n_ws < tc::MAX_WS) ++pl.n_ws;
  }
  const int64_t total = (int64_t)pl.n_mt * pl.n_nt * p.G;
  if (total > (1 << 30)) return false;
  pl.total_tiles = (int)total;
  // Weight multicast across a cluster of cs CTAs on consecutive row tiles: debug key 22 = 2 / 4 turns it on.  OFF by
  // default: alone, the deep layers gain 20-30 % (256 -> 256 k7: 30.5 -> 21.8 us, 1024 -> 1024 k5: 67.8 -> 49.4 us,
  // profiles/r2_cluster_sweep.log), but inside the train step, where 3..8 chains keep every SM busy, co-scheduling CTA
  // pairs costs more than the saved L2 traffic (12.46 ms off, 12.66 all layers, 12.54 only >= 1.1 MB of weights per
  // tile; profiles/r2_cluster_step.log).  Debug key 23: minimum KB of weights streamed per tile for a layer to cluster.
  pl.cs = 1;
  const int w_kb_tile = (int)((int64_t)pl.n_kc * p.ntaps * pl.bn * pl.row_bytes / 1024);
  if (allow_cluster && !pl.w_resident && (tc::g_debug[22] == 2 || tc::g_debug[22] == 4) && w_kb_tile >= tc::g_debug[23] &&
      (tc::g_debug[28] <= 0 || total <= tc::g_debug[28])) {            // key 28: only launches of at most that many tiles
    const int want = tc::g_debug[22];
    for (int c = want; c >= 2; c >>= 1)
      if (pl.n_mt % c == 0 && (pl.bn / c) % 8 == 0 && total >= 2 * c) { pl.cs = c; break; }
  }
  pl.n_mg = pl.n_mt / pl.cs;
    return true;
  };
  // Tile shape: a small clock model per candidate (bn, mt), fitted to tools/tc_sweep.py on B200:
  //   per MMA (M=128, K=16): max(tensor time bn/2, operand smem reads 32 + bn/4) clocks;
  //   epilogue ~700 clocks per [128 x 64] chunk pair, overlapped with the next main loop when the
  //   accumulator is double-buffered; every tile pays a ~600-clock hand-over bubble;
  //   all CTAs together cannot pull operands from L2 faster than ~2500 B/clk.
  tc::Plan pl, cand;
  double best = -1.0;
  const int bns[4] = {256, 128, 64, 32};
  const int sms_avail = n_share > 1 ? (num_sms() / n_share > 0 ? num_sms() / n_share : 1) : num_sms();
  for (int bi = 0; bi < 4; ++bi) {
    const int bn = bns[bi];
    if (p.Cog % bn != 0) continue;
    if (tc::g_debug[2] > 0 && bn != tc::g_debug[2]) continue;
    for (int mt_req = tc::MAX_MT; mt_req >= 1; --mt_req) {
      if (tc::g_debug[3] > 0 && mt_req != tc::g_debug[3]) continue;
      if (!make_plan(cand, bn, mt_req)) continue;
      if (cand.mt != mt_req && mt_req != tc::MAX_MT) continue;   // already evaluated at a larger request
      const double waves = (double)((cand.total_tiles + sms_avail - 1) / sms_avail);
      const double per_mma = bn / 2.0 > 32.0 + bn / 4.0 ? bn / 2.0 : 32.0 + bn / 4.0;
      const double main_clk = (double)cand.n_kc * p.ntaps * (cand.kch / 16) * cand.mt * per_mma;
      const double epi_clk = cand.mt * ((bn + 63) / 64) * 700.0;
      // operand streaming: bytes per tile over min(bytes in flight / ~3000-clock TMA latency, fair share of L2)
      const double w_tile = cand.w_resident ? 0.0 : (double)cand.n_kc * p.ntaps * bn * cand.row_bytes / cand.cs;
      const double tile_bytes = (double)(x3 ? 2 * cand.n_kcl : cand.n_kc) * cand.a_stage_bytes + w_tile;
      const double inflight = (double)cand.n_as * cand.a_stage_bytes + (cand.w_resident ? 0.0 : (double)cand.n_ws * cand.w_stage_bytes);
      // every SM is busy with some stream's CTA: the L2 share of a CTA is 1 / #SMs (debug key 17 = 1: of this launch only)
      const double active = (n_share > 1 || tc::g_debug[17] != 1) ? num_sms() : cand.total_tiles < num_sms() ? cand.total_tiles : num_sms();
      double bw = inflight / 3000.0;
      if (bw > 6000.0 / active) bw = 6000.0 / active;
      const double mem_clk = tile_bytes / bw;
      double tile_clk = main_clk > mem_clk ? main_clk : mem_clk;
      if (cand.acc_stages == 2) tile_clk = tile_clk > epi_clk ? tile_clk : epi_clk;
      else tile_clk += epi_clk;
      (void)waves;
      cand.ctas = cand.total_tiles;
      // Tiles per CTA (debug key 25 = largest factor tried, default 1): a CTA that loops over f tiles pays the fixed
      // cost once and — with a double-buffered accumulator — hides all but the last epilogue under the next main loop.
      const int fmax = tc::g_debug[25] > 1 ? tc::g_debug[25] : 1;
      for (int f = 1; f <= fmax; ++f) {
        // Objective: the SM TIME of the launch, not its stand-alone latency.  The train step runs 3..8 independent
        // chains on concurrent streams, every CTA owns an SM (shared memory), and the SM-occupancy trace
        // (tools/sm_timeline.py) shows the step bound by the sum of CTA lifetimes: a plan with fewer, fuller CTAs
        // that is slower alone leaves SMs to the other streams.  Measured: 14.4 -> 13.5 ms per step.
        // Debug key 15 = percentage of the SM-time term (default 100; -1 = pure latency), key 16 = fixed cost per
        // CTA in units of 500 clocks (default 6: launch, barrier / TMEM setup, first TMA round trip).
        const double a = tc::g_debug[15] > 0 ? tc::g_debug[15] / 100.0 : tc::g_debug[15] < 0 ? 0.0 : 1.0;
        const double fixed = tc::g_debug[16] > 0 ? 500.0 * tc::g_debug[16] : 3000.0;
        double ctas = (double)((cand.total_tiles + f - 1) / f);
        if (ctas > sms_avail) ctas = sms_avail;
        const double waves_f = (double)(((long long)cand.total_tiles + (long long)ctas - 1) / (long long)ctas);
        if (f > 1 && (waves_f < 2.0 || cand.cs > 1)) break;      // nothing left to fold
        const double tail = cand.acc_stages == 2 ? epi_clk : 0.0;
        const double lat = waves_f * (tile_clk + 600.0) + tail;
        // (f == 1 keeps the round-1 form of the SM-time term, without the exposed last epilogue: the defaults were
        // tuned against it)
        const double sm_time = ctas * (fixed + lat - (f == 1 ? tail : 0.0)) / sms_avail;
        const double cost_f = (1.0 - a) * lat + a * sm_time;
        if (best < 0 || cost_f < best) { best = cost_f; pl = cand; pl.ctas = (int)ctas; }
      }
    }
  }
  if (best < 0) return 0;
  pr_cost = best;
  pl.dbg = tc::g_dbg_buf;
  pl.dbg_flags = tc::g_debug[9];
  for (int t = 0; t < p.ntaps; ++t) pl.a_off16[t] = (pl.phase[t] * pl.panel_bytes + pl.shift[t] * pl.row_bytes) >> 4;

  // ---- tensor maps
  CUtensorMap& map_x = pr.map_x;
  CUtensorMap& map_w = pr.map_w;
  {
    const int ni = p.x.n_inner;
    cuuint64_t dims[4] = {(cuuint64_t)p.G * p.Cig, (cuuint64_t)ni, (cuuint64_t)p.x.len, (cuuint64_t)((p.N + ni - 1) / ni)};
    cuuint64_t strides[3] = {(cuuint64_t)(ni > 1 ? p.x.s_inner : p.x.s_row) * 2, (cuuint64_t)p.x.s_row * 2,
                             (cuuint64_t)p.x.s_outer * 2};
    if (dims[3] == 1 && strides[2] == 0) strides[2] = strides[1] * dims[2];
    // rows are traversed with stride si: a box of rows*si elements loads `rows` rows (one phase)
    cuuint32_t box[4] = {(cuuint32_t)pl.kch, 1, (cuuint32_t)((pl.packed ? pl.seg_rows : pl.boxr) * p.si), 1};
    cuuint32_t es[4] = {1, 1, (cuuint32_t)p.si, 1};
    CUresult rc = enc(&map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x3 ? p.X_sp : p.X), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, tc::swizzle_of(pl.row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc == CUDA_SUCCESS && x3)
      rc = enc(&pr.map_x_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
               const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(p.X_sp) + p.x_plane), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, tc::swizzle_of(pl.row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("artic_tapconv: cuTensorMapEncodeTiled(X) failed (%d)", (int)rc); return ARTIC_ECUDA; }
  }
  {
    int wmax = 0;
    for (int t = 0; t < p.ntaps; ++t) wmax = max(wmax, p.widx[t]);
    const int rows_total = p.Wt_taps > 0 ? p.Wt_taps : (wmax + 1);
    cuuint64_t dims[2] = {(cuuint64_t)p.Cig, (cuuint64_t)rows_total * p.G * p.Cog};
    cuuint64_t strides[1] = {(cuuint64_t)p.Cig * 2};
    cuuint32_t box[2] = {(cuuint32_t)pl.kch, (cuuint32_t)(pl.bn / pl.cs)};     // multicast: every CTA fetches a row slice
    cuuint32_t es[2] = {1, 1};
    CUresult rc = enc(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x3 ? p.Wt_sp : p.Wt), dims, strides, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, tc::swizzle_of(pl.row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc == CUDA_SUCCESS && x3)
      rc = enc(&pr.map_w_lo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
               const_cast<__nv_bfloat16*>(reinterpret_cast<const __nv_bfloat16*>(p.Wt_sp) + p.w_plane), dims, strides, box, es,
               CU_TENSOR_MAP_INTERLEAVE_NONE, tc::swizzle_of(pl.row_bytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { set_error("artic_tapconv: cuTensorMapEncodeTiled(W) failed (%d)", (int)rc); return ARTIC_ECUDA; }
  }
  pr_smem = pl.n_as * pl.a_stage_bytes + (pl.w_resident ? w_all_bytes(pl, p.ntaps, x3) : pl.n_ws * pl.w_stage_bytes) + 1024 + epi_bytes;
  pr.p = p;
  pr.pl = pl;
  return 1;
}

// Launches up to MAXP planned problems as ONE grid of per-problem persistent CTA ranges.
static int tc_launch_group(tc::Multi& mp, const int* smem, const double* cost, cudaStream_t st, bool x3) {
  const int n = mp.n;
  const int n_ew = (tc::g_debug[8] > 0 && tc::g_debug[8] * 1024 < tc::max_smem()) ? 4 : 8;
  int64_t tiles = 0;
  double work = 0.0;
  for (int j = 0; j < n; ++j) { tiles += mp.prob[j].pl.total_tiles; work += cost[j]; }
  int smem_bytes = 0;
  mp.cta_begin[0] = 0;
  const int cs = (n == 1) ? mp.prob[0].pl.cs : 1;       // clusters only in single-problem launches (tc_plan_problem)
  for (int j = 0; j < n; ++j) {
    const int t = mp.prob[j].pl.total_tiles;
    int c = t;
    if (tiles > num_sms()) {   // share the SMs in proportion to the planner's time estimate of each problem
      c = (int)(num_sms() * (cost[j] / work) + 0.5);
      if (c < 1) c = 1;
      if (c > t) c = t;
    }
    if (mp.prob[j].pl.ctas > 0 && c > mp.prob[j].pl.ctas) c = mp.prob[j].pl.ctas;     // several tiles per CTA by plan
    if (cs > 1) {              // whole clusters: at most floor(SMs / cs) of them, one per super tile
      int cl = t / cs < num_sms() / cs ? t / cs : num_sms() / cs;
      c = cl * cs;
    }
    mp.cta_begin[j + 1] = mp.cta_begin[j] + c;
    if (smem[j] > smem_bytes) smem_bytes = smem[j];
  }
  const int grid = mp.cta_begin[n];
  // Programmatic dependent launch (debug key 11 = 1 disables): the kernel may start under its predecessor
  // in the stream; it prefetches weights before its grid-dependency wait unless they were just rewritten.
  const bool pdl = tc::g_debug[11] != 1;
  const bool w_early = pdl && !tc::consume_weight_fence(st);
  for (int j = 0; j < n; ++j) mp.prob[j].pl.w_early = w_early ? 1 : 0;
  mp.trace = tc::g_trace_buf;
  mp.trace_cap = tc::g_trace_cap;
  mp.launch_id = tc::g_trace_buf != nullptr ? tc::g_trace_launch++ : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)(64 + 32 * n_ew));
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (w_early) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cs > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cs;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    g_path_counts[9] += 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t le = x3 ? cudaLaunchKernelEx(&cfg, tc::tapconv_tc_kernel<true>, mp) : cudaLaunchKernelEx(&cfg, tc::tapconv_tc_kernel<false>, mp);
  g_path_counts[x3 ? PATH_CONV_TC_X3 : PATH_CONV_TC] += n;
  if (le == cudaSuccess) le = cudaGetLastError();
  if (le != cudaSuccess) {
    const tc::Plan& pl = mp.prob[0].pl;
    set_error("artic_tapconv(tc): launch failed: %s (%d problems, grid %d, smem %d of %d, bn %d mt %d as %d ws %d packed %d)",
              cudaGetErrorString(le), n, grid, smem_bytes, tc::max_smem(), pl.bn, pl.mt, pl.n_as, pl.n_ws, pl.packed);
    return ARTIC_ECUDA;
  }
  return ARTIC_OK;
}

// Multi-problem front end used by artic_tapconv / artic_tapconv_multi (tapconv.cu): tries to plan every problem
// for the tensor-core kernel; taken[i] = 1 for those launched here (in groups of <= MAXP), 0 for the rest.
int artic_tapconv_tc_multi(const artic_tapconv_t* ps, int n, int* taken, cudaStream_t st) {
  static thread_local tc::Multi mp;       // ~13 KB: keep it off the stack
  int smem[tc::MAXP];
  double cost[tc::MAXP];
  mp.n = 0;
  int n_live = 0;
  for (int i = 0; i < n; ++i) n_live += (ps[i].N != 0 && ps[i].nq != 0) ? 1 : 0;
  const int n_share = tc::g_debug[10] == 1 ? 1 : n_live < tc::MAXP ? n_live : tc::MAXP;   // debug key 10 = 1: plan as if alone
  bool group_x3 = false;
  for (int i = 0; i < n; ++i) {
    taken[i] = 0;
    if (ps[i].N == 0 || ps[i].nq == 0) continue;
    const bool x3 = ps[i].dtype == ARTIC_F32;       // the two modes are different kernel instantiations
    if (mp.n > 0 && x3 != group_x3) {
      const int lrc = tc_launch_group(mp, smem, cost, st, group_x3);
      if (lrc != ARTIC_OK) return lrc;
      mp.n = 0;
    }
    const int rc = tc_plan_problem(&ps[i], mp.prob[mp.n], smem[mp.n], cost[mp.n], n_share, n_live == 1);
    if (rc < 0) return rc;
    if (rc == 0) continue;
    taken[i] = 1;
    group_x3 = x3;
    if (++mp.n == tc::MAXP) {
      const int lrc = tc_launch_group(mp, smem, cost, st, group_x3);
      if (lrc != ARTIC_OK) return lrc;
      mp.n = 0;
    }
  }
  if (mp.n > 0) {
    const int lrc = tc_launch_group(mp, smem, cost, st, group_x3);
    if (lrc != ARTIC_OK) return lrc;
  }
  return ARTIC_OK;
}
