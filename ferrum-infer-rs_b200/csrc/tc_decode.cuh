// tc_decode.cuh -- the decode-batch (<= 32 rows per block) form of the tensor-core W4A16 kernel.  Included by
// tc_gemm.cu (same translation unit: it shares TcParams, the stream-K bookkeeping and the launch helpers).
//
// What differs from tc_body (tc_gemm.cu):
//   * RAW NIBBLES.  The int4 weights go to the tensor core UNSCALED, as fp16 subnormals: `w & 0x000f000f` is the
//     half2 (q, q') * 2^-24, `w & 0x00f000f0` is (q, q') * 2^-20 -- 5 integer instructions per packed word
//     instead of 13 (tcgen05.mma kind::f16 honours subnormal operands: measured, round 2).  The tile format keeps
//     every MMA k-step class-pure (common.cuh, tile_nibble_k); the four hi-class k-steps of a unit are issued first
//     and the first lo-class MMA carries scale-input-d = 4 (D = A.B + D * 2^-4), so one accumulator holds
//     2^-24 * sum_k q_k x_k exactly.
//   * ZERO POINT ON THE TENSOR CORE, SCALE ON THE ACCUMULATOR SIDE.  Every unit (one 128 x 128 stage) gets a FRESH
//     accumulator D[slot].  The zero point is constant per (group, column) = per unit and TMEM lane, and the term
//     it contributes, -z * sum_k x_k, is linear: eight more MMAs run the unit's activation slice against an A
//     operand of 8 TMEM columns that hold half2(-z, -z) * 2^-24 (one tcgen05.st per unit), accumulating into the
//     same D.  D = 2^-24 * sum_k (q_k - z) x_k; the dequant warps read it back (tcgen05.ld, R columns = real
//     rows) and keep the running sum in REGISTERS: acc[r] += (s * 2^24) * D[r] -- one fp32 FMA per row and unit
//     instead of three fp16 instructions per weight pair.  The weight is never rounded to fp16: results are closer
//     to the f64 reference than the prefill path's.
//   * TWO UNITS IN FLIGHT PER CTA.  The eight dequant warps form two groups of four (one warp per TMEM lane
//     quarter, all 128 k of the unit); group s owns A slot s / D[s] and takes every second unit.  A unit costs each
//     warp two barrier waits instead of the old five, and while one group waits for its MMAs the other converts
//     -- the old loop was bound by that per-unit latency chain, not by instruction issue (halving its
//     instruction count bought 13 %).
//   * The accumulators live in registers, so the MMA warp knows nothing about tiles or segments, and the epilogue
//     starts from registers: the two groups swap the rows the other one finishes through shared memory, then run
//     the whole-tile / flag-in-value / counter exchange of tc_body -- with 16+ rows and 4+ contributors the rows of a
//     split tile are finished by all contributors but the last, a share each (one polling round instead of several).
//   * REGISTERS WHERE THE WORK IS.  The producer / MMA warps (one warpgroup) hand registers to the two dequant
//     warpgroups (setmaxnreg 48 / 96); the running sums are never indexed at run time, so nothing of the unit loop
//     lives in local memory (which does not fit the L1 this kernel leaves).
// Row sources: dense rows through the tensor map (DK_DENSE); dense rows of an act-order layer gathered inside the
// kernel (DK_GATHER); moe_align_block_size blocks of a stacked expert weight (DK_ROUTED, block <= 16), optionally with
// the gate | up accumulator pair and the SwiGLU epilogue (kSilu).
#pragma once

#ifndef FB2_QUARTER_MIN_R
#define FB2_QUARTER_MIN_R 32
#endif
namespace dk {

enum { DK_DENSE = 0, DK_GATHER = 1, DK_ROUTED = 2 };

constexpr int kWarpMma2 = 11;               // second MMA issuer: one per dequant group, so the groups are not coupled through
                                            // one in-order issue loop (the 12th warp costs no registers: 11 warps already
                                            // occupy the allocation of 12)
constexpr int kDkThreads = 12 * 32;
constexpr int kXSlots = 4;                  // activation ring: slots of a dense launch ...
constexpr int kXSlotsRouted = 8;            // ... and of a routed one (its gathers are L2 round trips per unit: the down projection
                                            // of a MoE layer waited for activations a third of the time with four)
constexpr int kSmemBudget = 108 * 1024;     // two CTAs per SM
// shared memory: [X slots][group exchange][barriers + misc][weight ring: run-time depth][DK_GATHER: the x rows]
// NB = UMMA N = rows of a block: 16, or 32 for the 17..32-row dense launches
template <bool kSilu, int NB, int SX = kXSlots>
struct Lay {
  static constexpr int kNacc = kSilu ? 2 : 1;
  static constexpr int kXBytes = NB * 256;                       // [2 atoms][NB rows][128 B]
  static constexpr int kXchgBase = SX * kXBytes;
  static constexpr int kXchgBytes = kNacc * NB * kTileN * 4;     // group <-> group exchange of partial sums
  static constexpr int kBarBase = kXchgBase + kXchgBytes;
  static constexpr int kWBase = kBarBase + 1024;
};
constexpr int kMinWStages = 5;
__host__ __device__ constexpr int w_stages_for(bool silu, int xs_bytes, int nb = 16) {
  const int base = silu ? Lay<true, 16>::kWBase : (nb > 16 ? Lay<false, 32>::kWBase : Lay<false, 16>::kWBase);
  const int n = (kSmemBudget - base - xs_bytes) / kStageBytes;
  return n < kMaxWStages ? n : kMaxWStages;
}
static_assert(w_stages_for(true, 0) >= 7 && w_stages_for(false, 0, 32) >= 6, "weight ring depth");
// tensor memory columns (256 allocated): A slots, per-slot accumulators, per-slot zero-point operand
constexpr int kTmemCols = 256;
constexpr int kColA = 0;                    // 2 slots x 64
constexpr int kColD = 128;                  // 2 x NB: 2^-24 * sum_k (q_k - z) x_k; behind it 2 x 8 columns of half2(-z, -z) * 2^-24
enum { BAR_FULL_W = 0, BAR_EMPTY_W = kMaxWStages, BAR_FULL_X = 2 * kMaxWStages, BAR_EMPTY_X = BAR_FULL_X + kXSlotsRouted,
       BAR_FULL_A = BAR_EMPTY_X + kXSlotsRouted, BAR_DONE = BAR_FULL_A + 2, BAR_XS = BAR_DONE + 2, BAR_COUNT = BAR_XS + 1 };
static_assert(BAR_COUNT * 8 <= 512, "barrier area too small");
// TcParams::dk_flags
enum { DKF_META_EARLY = 1,      // routed: the routing arrays were complete before the PREVIOUS kernel started its
                                // dependents (fb2_moe_ffn_forward: the down GEMM behind the gate|up GEMM): planning and the
                                // weight stream do not wait for the grid dependency
       DKF_LATE_TRIGGER = 2 };  // launch_dependents only after this kernel's own grid dependency has resolved (what makes
                                // DKF_META_EARLY of the next kernel true)

template <int R>
__device__ __forceinline__ void tmem_ld_rows(uint32_t taddr, uint32_t (&r)[R]) {
  if constexpr (R == 1) ptx::tmem_ld_x1(taddr, r);
  else if constexpr (R == 2) ptx::tmem_ld_x2(taddr, r);
  else if constexpr (R == 4) ptx::tmem_ld_x4(taddr, r);
  else if constexpr (R == 8) ptx::tmem_ld_x8(taddr, r);
  else ptx::tmem_ld_x16(taddr, r);
}

template <int VW>
__device__ __forceinline__ void ld_slots(const uint32_t *a, uint32_t (&o)[VW]) {
  if constexpr (VW == 1) o[0] = ptx::ld_relaxed_u32(a);
  else if constexpr (VW == 2) ptx::ld_relaxed_v2(a, o);
  else ptx::ld_relaxed_v4(a, o);
}
template <int VW>
__device__ __forceinline__ void st_slots(uint32_t *a, const uint32_t (&u)[VW]) {
  if constexpr (VW == 1) ptx::st_relaxed_u32(a, u[0]);
  else if constexpr (VW == 2) ptx::st_relaxed_v2(a, u);
  else ptx::st_relaxed_v4(a, u);
}
template <int VW>
__device__ __forceinline__ void reset_slots(uint32_t *a) {     // plain stores: the kernel boundary publishes them
  if constexpr (VW == 1) *a = 0xffffffffu;
  else if constexpr (VW == 2) *reinterpret_cast<uint2 *>(a) = make_uint2(0xffffffffu, 0xffffffffu);
  else *reinterpret_cast<uint4 *>(a) = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
}

// One CTA's share of the launch: units [u0, u1) of T, among G participating CTAs.
struct Plan {
  u32 T, u0, u1;
  int G;
};
template <int kMode, int NACC>
__device__ __forceinline__ Plan make_plan_dev(const TcParams &p) {
  Plan pl;
  pl.G = (int)gridDim.x;
  const u32 cta = blockIdx.x, upt = (u32)(NACC * p.groups), nt = (u32)p.n_tiles;
  if constexpr (kMode != DK_ROUTED) {
    pl.T = p.fd_units.d;                                         // dense rows, one block: n_tiles * upt
    if (p.whole_tiles) {
      pl.u0 = fdiv(cta * nt, p.fd_grid) * upt;
      pl.u1 = fdiv((cta + 1) * nt, p.fd_grid) * upt;
    } else {
      pl.u0 = fdiv(cta * pl.T, p.fd_grid);
      pl.u1 = fdiv((cta + 1) * pl.T, p.fd_grid);
    }
  } else {
    // The block count comes from the align kernel.  The grid was sized for the worst case; with fewer units only
    // the first CTAs take part, so that every participating CTA owns a non-empty range and the contributors of a
    // tile are consecutive CTAs.
    u32 nblk = (u32)__ldcg(p.rm.num_post_padded) / (u32)p.rm.block;
    if (nblk > (u32)p.rm.max_blocks) nblk = (u32)p.rm.max_blocks;
    const u32 ntt = nblk * nt;
    pl.T = ntt * upt;
    const bool whole = whole_tile_mode((int)upt, ntt, pl.G);
    const u32 cap = whole ? ntt : pl.T / 2;                      // whole tiles | >= 2 units per CTA
    const u32 g_eff = cap > 1 ? cap : 1;
    if (g_eff < (u32)pl.G) pl.G = (int)g_eff;
    pl.u0 = pl.u1 = 0;
    if (cta < (u32)pl.G && pl.T > 0) {
      if (whole) {
        pl.u0 = tile_range_start(cta, ntt, (u32)pl.G, upt);
        pl.u1 = tile_range_start(cta + 1, ntt, (u32)pl.G, upt);
      } else {
        pl.u0 = range_start(cta, pl.T, (u32)pl.G);
        pl.u1 = range_start(cta + 1, pl.T, (u32)pl.G);
      }
    }
  }
  return pl;
}

// R = accumulator columns each dequant thread reads per unit: the real rows of a block rounded up to a power of two
// (routed: the block size).
template <int R, int kMode, bool kSilu, bool kPush>
__device__ __forceinline__ void dk_body(const TcParams &p, const CUtensorMap *xmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  static_assert(R == 1 || R == 2 || R == 4 || R == 8 || R == 16 || R == 32, "rows per thread");
  constexpr int kNB = R > 16 ? 32 : 16;             // UMMA N
  constexpr int kColZ = kColD + 2 * kNB;
  static_assert(kColZ + 16 <= kTmemCols, "tensor memory map");
  static_assert(!(kSilu && kPush) && !(kPush && kMode == DK_ROUTED) && !(kSilu && R > 8), "combination not instantiated");
  constexpr int SX = kMode == DK_ROUTED ? kXSlotsRouted : kXSlots;
  using L = Lay<kSilu, kNB, SX>;
  constexpr int kXBytes = L::kXBytes;
  constexpr int NACC = L::kNacc;
  constexpr bool kRouted = kMode == DK_ROUTED, kGather = kMode == DK_GATHER;
  constexpr int OWN = R <= 4 ? R : R / 2;           // rows a group finishes in the epilogue (group 1: none when R <= 4)
  constexpr int kEpiThreads = R <= 4 ? 128 : 256;   // threads that take part in the inter-CTA exchange
  constexpr bool kDefer = kRouted;                  // deferred segment epilogue (dense ranges have one or two segments)
  constexpr bool kQuarter = kSilu || R >= FB2_QUARTER_MIN_R;      // conversion granularity (register pressure vs tcgen05.st count)
  const int SW = p.w_stages;                                     // weight ring depth (host: w_stages_for)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FB2_TL(0);
  if (warp == kWarpW) FB2_TLW(16);
  const int cta = (int)blockIdx.x;
  const int groups = p.groups, n_tiles = p.n_tiles;
  const int upt = NACC * groups;                                 // units per output tile
  const bool meta_early = kRouted && (p.dk_flags & DKF_META_EARLY);
  Plan pl{0, 0, 0, (int)gridDim.x};
  if constexpr (!kRouted) {
    pl = make_plan_dev<kMode, NACC>(p);
    if (pl.u0 >= pl.u1) return;                                  // CTA-uniform, before any barrier
  }

  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bars = smem_base + L::kBarBase;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + L::kBarBase + 512);
  const size_t tile_stride = (size_t)groups * kStageBytes;
  // routed launches learn their range from the align kernel's output: every role plans for itself, after the grid
  // dependency (or at once when the arrays are known to be complete)
  auto plan_routed = [&]() {
    if constexpr (kRouted) {
      if (!meta_early) ptx::grid_dependency_wait();
      pl = make_plan_dev<kMode, NACC>(p);
    }
  };
  auto seg_init = [&](SegIter &si) {
    if constexpr (kRouted) si.init(pl.u0, pl.u1, upt, n_tiles);
    else si.init_fd(pl.u0, pl.u1, p.fd_upt, p.fd_tiles);
  };

  // ------------------------------------------------------------------ one-time setup
  if (warp == kWarpW) {
    // one barrier pair per lane; the CTA-wide arrive follows the first weight request (below)
    if (lane < SW) {
      ptx::mbar_init(bar(BAR_FULL_W + lane), 1);                 // the producer's arrive.expect_tx
      ptx::mbar_init(bar(BAR_EMPTY_W + lane), 4);                // read into registers by the 4 warps of one group
    }
    ptx::mbar_fence_init();
    __syncwarp();
    FB2_TLW(17);
  } else if (warp == kWarpX) {
    if (lane == 0) {
      for (int i = 0; i < SX; i++) {
        ptx::mbar_init(bar(BAR_FULL_X + i), kRouted ? 1 + 32 : 1);   // routed: + the 32 cp.async arrivals of a gathered slice
        ptx::mbar_init(bar(BAR_EMPTY_X + i), 1);                 // the MMA commit
      }
      ptx::mbar_init(bar(BAR_XS), 1);
      ptx::mbar_fence_init();
    }
    __syncwarp();
    ptx::named_bar_arrive(2, kDkThreads);
  } else if (warp == kWarpMma) {
    if (lane == 0) {
      for (int i = 0; i < 2; i++) {
        ptx::mbar_init(bar(BAR_FULL_A + i), 4);
        ptx::mbar_init(bar(BAR_DONE + i), 1);
      }
      ptx::mbar_fence_init();
    }
    __syncwarp();
    ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(s_tmem));
  }
  uint32_t tmem_base = 0;
  if (warp != kWarpW && warp != kWarpX) {      // (the second MMA warp takes part like a dequant warp)
    ptx::tc_fence_before();
    ptx::named_bar_sync(2, kDkThreads);
    ptx::tc_fence_after();
    tmem_base = *s_tmem;
  }
  FB2_TL(1);
#ifdef FB2_TRACE
  { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); FB2_TLV(10, smid); }
#endif

  // Warps 8..11 (weight / activation producers, MMA issuers) form one warpgroup and need few registers; the two dequant
  // warpgroups take what they release: 48 + 2 x 96 = 3 x 80 per thread, and the running sums, the unit being converted and
  // the routed bookkeeping stay in registers (at 80 the routed variants spilled loop-carried state to local memory, which
  // does not fit the ~24 KB of L1 this kernel's shared memory leaves: every spill load was an L2 round trip in the unit loop).
  if (warp >= kDequantWarps) {
  ptx::setmaxnreg_dec<48>();
  if (warp == kWarpW) {
    // ========================================================== weight producer (HBM stream)
    // Weights depend on no earlier kernel: under programmatic dependent launch this warp streams ahead of
    // griddepcontrol.wait (a routed launch must first learn which experts it streams).
    plan_routed();
    const uint64_t pol = ptx::policy_evict_first();
    RingPos wp{0, 0};
    u32 it = 0;
    SegIter si;
    seg_init(si);
    Seg sg;
    FB2_TLW(18);
    while (si.next(sg, upt, n_tiles)) {
      const int expert = kRouted ? __ldcg(p.rm.block_experts + sg.blk) : 0;
#pragma unroll 1
      for (int a = 0; a < NACC; a++) {
        const int g0 = max(sg.r0, a * groups) - a * groups, g1 = min(sg.r1, (a + 1) * groups) - a * groups;
        if (g0 >= g1) continue;
        const uint8_t *wsrc = p.w.tiles +
            ((size_t)expert * p.tiles_per_expert + p.w.col0 / kTileN + sg.tile + (size_t)a * n_tiles) * tile_stride +
            (size_t)g0 * kStageBytes;
#pragma unroll 1
        for (int g = g0; g < g1; g++, it++, wsrc += kStageBytes, wp.advance(SW)) {
          if (it >= (u32)SW) ptx::mbar_wait(bar(BAR_EMPTY_W + wp.st), wp.par ^ 1u);
          if (it == 0) FB2_TLW(11);
          if (ptx::elect_one()) {
            const uint32_t full = bar(BAR_FULL_W + wp.st);
            ptx::mbar_arrive_expect_tx(full, (uint32_t)kStageBytes);
            ptx::bulk_g2s_hint(smem_base + L::kWBase + wp.st * kStageBytes, wsrc, kStageBytes, full, pol);
          }
          __syncwarp();
          if (it == 0) ptx::named_bar_arrive(2, kDkThreads);       // ring barriers are initialised (setup barrier)
        }
      }
    }
    if (it == 0) ptx::named_bar_arrive(2, kDkThreads);             // routed launch with nothing for this CTA
  } else if (warp == kWarpX) {
    // ====================================================== activation producer (L2 hits)
    if (kMode == DK_DENSE && lane == 0) ptx::prefetch_tensormap(xmap);
    const int m = p.rm.m;
    const uint32_t xs_base = smem_base + L::kWBase + (uint32_t)SW * kStageBytes;   // DK_GATHER: x rows [m][K], natural order
    int4 pm = make_int4(0, 0, 0, 0);                              // DK_GATHER: perm of the next unit, 4 k' per lane
    if (kGather || kRouted) {
      // rows the gathers never write feed accumulator columns nobody reads, but they should not hold NaN patterns
      for (int i = lane; i < SX * kXBytes / 16; i += 32)
        *reinterpret_cast<uint4 *>(smem + (size_t)i * 16) = make_uint4(0, 0, 0, 0);
      ptx::fence_proxy_async_smem();
    }
    if (kGather) {
      int g_first = (int)(pl.u0 - fdiv(pl.u0, p.fd_upt) * (u32)upt);
      if (NACC == 2 && g_first >= groups) g_first -= groups;      // an up-projection unit reads the same slice as its gate unit
      pm = __ldg(reinterpret_cast<const int4 *>(p.w.perm + (size_t)g_first * kGroup) + lane);   // a weight-side constant
    }
    ptx::grid_dependency_wait();
    if (p.dk_flags & DKF_LATE_TRIGGER) ptx::launch_dependents();
    FB2_TLW(12);
    if constexpr (kRouted) pl = make_plan_dev<kMode, NACC>(p);
    if (kGather) {
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(bar(BAR_XS), (uint32_t)m * (uint32_t)p.w.k * 2u);
        for (int r = 0; r < m; r++)
          ptx::bulk_g2s(xs_base + (uint32_t)r * (uint32_t)p.w.k * 2u, p.x + (size_t)r * p.w.k, (uint32_t)p.w.k * 2u, bar(BAR_XS));
      }
      __syncwarp();
      ptx::mbar_wait(bar(BAR_XS), 0);
    }
    SegIter si;
    seg_init(si);
    Seg sg;
    RingPos xp{0, 0};
    u32 it = 0;
    while (si.next(sg, upt, n_tiles)) {
      const int row0 = sg.blk * kNB;
      int arow = -1, nchunks = 0;                                 // routed: lane i holds the activation row of slot i
      if constexpr (kRouted) {
        if (lane < p.rm.block) {
          const int id = __ldcg(p.rm.sorted_ids + (size_t)sg.blk * p.rm.block + lane);
          if (id < p.rm.num_valid) arow = id / p.rm.top_k;        // marlin contract: A row = id / top_k
        }
        // valid rows form a prefix of the block (moe_align_block_size puts real ids first)
        nchunks = __popc(__ballot_sync(0xffffffffu, arow >= 0)) * 16;
      }
#pragma unroll 1
      for (int r = sg.r0; r < sg.r1; r++, it++, xp.advance(SX)) {
        const int g = (NACC == 2 && r >= groups) ? r - groups : r;   // gate and up units read the same slice
        const uint32_t slot_addr = smem_base + xp.st * kXBytes;
        const uint32_t full = bar(BAR_FULL_X + xp.st);
        if (it >= (u32)SX) ptx::mbar_wait(bar(BAR_EMPTY_X + xp.st), xp.par ^ 1u);
        if constexpr (kMode == DK_DENSE) {
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(full, (uint32_t)kXBytes);
            // [16 rows x 64 k] boxes, hardware 128B swizzle; rows past the tensor end read as zero
            ptx::tma_load_2d(slot_addr, xmap, g * kGroup, row0, full);
            ptx::tma_load_2d(slot_addr + kNB * 128, xmap, g * kGroup + 64, row0, full);
          }
          __syncwarp();
        } else if constexpr (kRouted) {
          // gathered slice: rows [0, nvalid) x 128 k, 16 chunks of 16 B per row, straight into the swizzled layout
          // (this IS the MoE gather x[id / top_k], dispatch.rs:1763)
          if (ptx::elect_one()) ptx::mbar_arrive_expect_tx(full, 0u);
          __syncwarp();
          const __half *xg = p.x + (size_t)g * kGroup;
#ifdef FB2_EXP_NOX   // timing probe: no row gather (results are wrong)
          nchunks = 0;
#endif
          for (int ch0 = 0; ch0 < nchunks; ch0 += 32) {           // warp-uniform trip count
            const int ch = ch0 + lane;
            const int i = ch >> 4, c8 = ch & 15;
            const int ar = __shfl_sync(0xffffffffu, arow, i & 31);
            if (ch < nchunks) {
              const uint32_t dst = slot_addr + (uint32_t)(c8 >> 3) * (kNB * 128) + (uint32_t)i * 128 +
                                   (uint32_t)(((c8 & 7) ^ (i & 7)) << 4);
              ptx::cp_async_16(dst, xg + (size_t)ar * p.w.k + c8 * 8);
            }
          }
          ptx::cp_async_mbar_arrive_noinc(full);
        } else {
          // lane l builds k' = 4l .. 4l+3 of every real row: 8 bytes inside one 16-byte chunk of the swizzled row
          const int4 cur = pm;
          {
            // perm of the unit after this one (the next group of the tile, or group 0 of the next tile)
            const int gn = g + 1 < groups ? g + 1 : 0;
            pm = __ldg(reinterpret_cast<const int4 *>(p.w.perm + (size_t)gn * kGroup) + lane);
          }
          const uint32_t chunk = (uint32_t)(lane >> 1) & 7u, atom = (uint32_t)lane >> 4, sub = (uint32_t)(lane & 1) * 8u;
#pragma unroll 2
          for (int rr = 0; rr < m; rr++) {
            const uint32_t row = xs_base + (uint32_t)rr * (uint32_t)p.w.k * 2u;
            uint32_t v0, v1, v2, v3;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v0) : "r"(row + (uint32_t)cur.x * 2u));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v1) : "r"(row + (uint32_t)cur.y * 2u));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v2) : "r"(row + (uint32_t)cur.z * 2u));
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v3) : "r"(row + (uint32_t)cur.w * 2u));
            const uint32_t dst = slot_addr + atom * (kNB * 128) + (uint32_t)rr * 128u + ((chunk ^ ((uint32_t)rr & 7u)) << 4) + sub;
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(dst), "r"(v0 | (v1 << 16)), "r"(v2 | (v3 << 16)) : "memory");
          }
          ptx::fence_proxy_async_smem();                         // generic-proxy writes -> visible to the UMMA reads
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(full);
        }
      }
    }
  } else if (warp == kWarpMma || warp == kWarpMma2) {
    // ==================================================================== MMA issuers (one per group / A slot)
    plan_routed();
    const u32 n_units = pl.u1 - pl.u0;
    constexpr uint32_t idesc = ptx::umma_idesc_f16(128, kNB);
    const uint64_t b_desc_ring = ptx::umma_desc_kmajor_sw128(smem_base);
    const uint32_t s = warp == kWarpMma2 ? 1u : 0u;
    const uint32_t a = tmem_base + kColA + s * 64;
    const uint32_t d = tmem_base + kColD + s * kNB, zc = tmem_base + kColZ + s * 8;
    static_assert(SX == 4 || SX == 8, "activation ring position is computed from the unit index");
#pragma unroll 1
    for (u32 li = s; li < n_units; li += 2) {
      const uint32_t xst = li & (u32)(SX - 1), xpar = (li / (u32)SX) & 1u;
      ptx::mbar_wait(bar(BAR_FULL_X + xst), xpar);                // activation slice landed
      if (kRouted) ptx::fence_proxy_async_smem();                 // cp.async rows were generic-proxy writes
      if (li == 0) FB2_TLW(13);
      ptx::mbar_wait(bar(BAR_FULL_A + s), (li >> 1) & 1u);        // raw nibbles of the unit are in A slot s
      if (li == 0) FB2_TLW(14);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint64_t b0 = b_desc_ring + (uint64_t)((xst * kXBytes) >> 4);
        const uint64_t b1 = b0 + (uint64_t)((kNB * 128) >> 4);    // second 64-k atom
        // hi class (q * 2^-20): k-steps 2, 3 of each 64-k half; the first MMA of a unit overwrites D
        ptx::umma_f16_ts(d, a + 16, b0 + 4, idesc, 0u);
        ptx::umma_f16_ts(d, a + 24, b0 + 6, idesc, 1u);
        ptx::umma_f16_ts(d, a + 48, b1 + 4, idesc, 1u);
        ptx::umma_f16_ts(d, a + 56, b1 + 6, idesc, 1u);
        // lo class (q * 2^-24): D = A.B + D / 16 brings the hi sum to the same unit
        ptx::umma_f16_ts_scaled<4>(d, a + 0, b0 + 0, idesc);
        ptx::umma_f16_ts(d, a + 8, b0 + 2, idesc, 1u);
        ptx::umma_f16_ts(d, a + 32, b1 + 0, idesc, 1u);
        ptx::umma_f16_ts(d, a + 40, b1 + 2, idesc, 1u);
        // zero point: D -= z * sum_k x_k (the same 8 columns of -z * 2^-24 against every k-step of the slice)
#pragma unroll
        for (int kk = 0; kk < 4; kk++) ptx::umma_f16_ts(d, zc, b0 + (uint64_t)(kk * 2), idesc, 1u);
#pragma unroll
        for (int kk = 0; kk < 4; kk++) ptx::umma_f16_ts(d, zc, b1 + (uint64_t)(kk * 2), idesc, 1u);
        ptx::tc_commit(bar(BAR_DONE + s));                         // A slot s free, D[s] complete
        ptx::tc_commit(bar(BAR_EMPTY_X + xst));                    // activation slot consumed
      }
      __syncwarp();
    }
  }
  } else {
    ptx::setmaxnreg_inc<96>();
    // ======================================================== dequant + accumulate + epilogue warps
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int s = warp >> 2;              // group: A slot / accumulator, parity of the units it takes
    const int nl = q * 32 + lane;         // weight column inside the tile == TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t k_w = smem_base + L::kWBase + (uint32_t)nl * 16u, k_sz = smem_base + L::kWBase + kStageWBytes + (uint32_t)nl * 4u,
             k_ta = lane_base + kColA + s * 64, k_d = lane_base + kColD + s * kNB, k_z = lane_base + kColZ + s * 8,
             k_fa = bar(BAR_FULL_A + s), k_dn = bar(BAR_DONE + s), k_fw = bar(BAR_FULL_W), k_ew = bar(BAR_EMPTY_W);
    // loop invariants through a self-shuffle: ptxas otherwise re-derives them from %tid in every iteration
    k_w = __shfl_sync(0xffffffffu, k_w, lane); k_sz = __shfl_sync(0xffffffffu, k_sz, lane);
    k_ta = __shfl_sync(0xffffffffu, k_ta, lane); k_d = __shfl_sync(0xffffffffu, k_d, lane);
    k_z = __shfl_sync(0xffffffffu, k_z, lane); k_fa = __shfl_sync(0xffffffffu, k_fa, lane);
    k_dn = __shfl_sync(0xffffffffu, k_dn, lane); k_fw = __shfl_sync(0xffffffffu, k_fw, lane);
    k_ew = __shfl_sync(0xffffffffu, k_ew, lane);
    RingPos wp{s, 0};                     // this group's next unit is local unit s: ring stage s
    uint32_t dpar = 0;                    // parity of BAR_DONE[s]
    bool pending = false;                 // a unit of this group is in the tensor core: D[s] not accumulated yet
    bool pend_up = false;                 // ... and it is an up-projection unit (kSilu)
    float sc_prev = 0.f;                  // its scale * 2^24
    float acc[NACC][R];
#pragma unroll
    for (int a = 0; a < NACC; a++)
#pragma unroll
      for (int r = 0; r < R; r++) acc[a][r] = 0.f;
    bool dep_waited = meta_early ? false : kRouted;   // routed planning has waited already unless the arrays came early
    bool tl_first = true;
    u32 lbase = 0;                        // local index of the segment's first unit
    plan_routed();
    FB2_TLV(8, pl.u1 - pl.u0);
    const u32 T = pl.T;
    const int G = pl.G;
    SegIter si;
    seg_init(si);
    Seg sg;

    // Fold the finished unit of this group into the running sums (warp-collective).  Returning from the wait also
    // means that A slot s is free again.
    auto wait_done = [&]() {
      ptx::mbar_wait(k_dn, dpar);
      dpar ^= 1u;
      ptx::tc_fence_after();
    };
    auto fold = [&]() {
      constexpr int RC = R > 16 ? 16 : R;           // columns per tcgen05.ld (R = 32: two halves, 16 fewer live registers)
#pragma unroll
      for (int c0 = 0; c0 < R; c0 += RC) {
        uint32_t dv[RC];
        tmem_ld_rows<RC>(k_d + c0, dv);
        ptx::tmem_wait_ld();
        if (NACC == 2 && pend_up) {
#pragma unroll
          for (int r = 0; r < RC; r++) acc[NACC - 1][c0 + r] = fmaf(sc_prev, __uint_as_float(dv[r]), acc[NACC - 1][c0 + r]);
        } else {
#pragma unroll
          for (int r = 0; r < RC; r++) acc[0][c0 + r] = fmaf(sc_prev, __uint_as_float(dv[r]), acc[0][c0 + r]);
        }
      }
    };
    // one 64-k half: two LDS.128 -> 16 lo-class + 16 hi-class half2 registers
    auto convert_half = [&](const uint4 &wa, const uint4 &wb, uint32_t (&lo)[16], uint32_t (&hi)[16]) {
      const uint32_t ws[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t w = ws[j], t = w >> 8;
        lo[2 * j] = w & 0x000f000fu;
        hi[2 * j] = w & 0x00f000f0u;
        lo[2 * j + 1] = t & 0x000f000fu;
        hi[2 * j + 1] = t & 0x00f000f0u;
      }
    };
    // one LDS.128 (4 packed words = 32 k of this column): 8 lo-class + 8 hi-class half2 registers
    auto convert_quarter = [&](const uint4 &wv, uint32_t (&lo)[8], uint32_t (&hi)[8]) {
      const uint32_t ws[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t w = ws[j], t = w >> 8;
        lo[2 * j] = w & 0x000f000fu;
        hi[2 * j] = w & 0x00f000f0u;
        lo[2 * j + 1] = t & 0x000f000fu;
        hi[2 * j + 1] = t & 0x00f000f0u;
      }
    };

    // ---- the epilogue of one tile segment (this group's share), from the running sums in registers.  It runs
    // DEFERRED: after this group has folded its last unit of the segment, which happens while its first unit of the
    // next segment is already in the tensor core -- a segment boundary does not drain the pipeline (the grouped
    // down projection has 6-unit tiles: draining at every boundary cost it a third of its time).
    struct SegInfo { int blk, tile, c_first, c_last; bool whole; };
    auto epilogue = [&](const SegInfo &e) {
      FB2_TL(5);
      if (!dep_waited) {                  // before the first global write / scratch access
        ptx::grid_dependency_wait();
        dep_waited = true;
        FB2_TL(2);
      }
      const int blk = e.blk, tile = e.tile, tau = e.blk * n_tiles + e.tile;
      const bool whole = e.whole;
      const int c_first = e.c_first, c_last = e.c_last;
      // real rows of the block: a prefix (dense: m; routed: moe_align_block_size puts real ids first)
      int nv;
      int row_id = 0;                                            // routed: lane i holds the pair id of row i of the block
      if constexpr (kRouted) {
        row_id = lane < p.rm.block ? __ldcg(p.rm.sorted_ids + (size_t)blk * p.rm.block + lane) : p.rm.num_valid;
        nv = __popc(__ballot_sync(0xffffffffu, row_id < p.rm.num_valid));
      } else {
        nv = min(kNB, p.rm.m - blk * kNB);
      }
      // the two groups swap the rows the other one finishes
      float *xc = reinterpret_cast<float *>(smem + L::kXchgBase);
      ptx::named_bar_sync(1, kDequantWarps * 32);                 // everybody is done reading the previous exchange
#pragma unroll
      for (int a = 0; a < NACC; a++) {
        if constexpr (R <= 4) {
          if (s == 1) {
#pragma unroll
            for (int r = 0; r < R; r++) xc[(a * kNB + r) * kTileN + nl] = acc[a][r];
          }
        } else {
          // (selects between compile-time indices: a run-time index would move the running sums to local memory, and the
          // deferred epilogue inside the unit loop would then write them through on every unit)
#pragma unroll
          for (int jj = 0; jj < OWN; jj++)
            xc[(a * kNB + (1 - s) * OWN + jj) * kTileN + nl] = s ? acc[a][jj] : acc[a][OWN + jj];
        }
      }
      ptx::named_bar_sync(1, kDequantWarps * 32);
      const bool has_rows = R > 4 || s == 0;
      const int rb = R > 4 ? s * OWN : 0;                         // first row this group finishes
      float v[NACC][OWN];
#pragma unroll
      for (int a = 0; a < NACC; a++) {
#pragma unroll
        for (int jj = 0; jj < OWN; jj++) {
          const float mine = (R > 4 && s) ? acc[a][(R > 4 ? OWN : 0) + jj] : acc[a][jj];
          v[a][jj] = has_rows ? mine + xc[(a * kNB + rb + jj) * kTileN + nl] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < R; r++) acc[a][r] = 0.f;
      }
      if (!has_rows) return;

      constexpr int kPart = NACC * kNB * kTileN;
      const int col = tile * kTileN + nl;
      const float bias = (!kSilu && p.w.bias) ? p.w.bias[col] : 0.0f;
      const int ocol = (kRouted && p.out_map) ? __ldg(p.out_map + col) : col;
      const size_t push_off = push_world(p) ? ((size_t)(*(volatile const uint32_t *)p.push.epoch & 1u) * p.push.world +
                                              p.push.rank) * p.push.slot_elems : 0;
      const bool owner = !whole && cta == c_first;
      // flag-in-value exchange: the slot area of a CTA, [2 accumulators][8 rows][128], holds either pair of 8-row
      // accumulators or the 16 rows of one
      // 16+ rows: distributed owners (below); fewer: the first contributor finishes the tile alone -- as fast while one
      // polling round covers every contributor, and the others leave their SM slots to the next launch earlier
      constexpr bool kDist = R >= 16;
      constexpr int kAreaFloats = NACC * R * kTileN;             // kDist: one slot area, [acc][R rows][128]; a CTA has two
      static_assert(!kDist || kAreaFloats % kLlFloatsPerCta == 0, "slot areas are whole tc_body areas");
      const bool ll = !whole && (kDist ? G * 2 * (kAreaFloats / kLlFloatsPerCta) <= p.ll_ctas : G <= p.ll_ctas);
      const int nr = nv - rb;                                    // real rows among this group's: j < nr
      auto store_out = [&](int row_in_block, float v0, float v1) {
        int cr;
        if constexpr (kRouted) cr = __shfl_sync(0xffffffffu, row_id, row_in_block);   // C row = pair id (warp-uniform call)
        else cr = blk * kNB + row_in_block;
        float o;
        if constexpr (kSilu) {
          // the reference rounds gate | up to f16 before SiLU (fused_silu_mul.cu reads f16; generic_kernels.cu does the same)
          o = silu_mul(__half2float(__float2half_rn(v0)), __half2float(__float2half_rn(v1)));
        } else {
          o = v0 + bias;
        }
        const __half hv = __float2half_rn(o);
        if (push_world(p) == 0) {
          p.y[(size_t)cr * p.ld_out + ocol] = hv;
        } else {
          for (int r = 0; r < push_world(p); r++) p.push.data[r][push_off + (size_t)cr * p.ld_out + col] = hv;
        }
      };
      if (whole) {
#pragma unroll
        for (int jj = 0; jj < OWN; jj++)
          if (jj < nr) store_out(rb + jj, v[0][jj], v[NACC - 1][jj]);
      } else if (ll) {
        if constexpr (kDist) {
          // ---- flag-in-value exchange (protocol: tc_gemm.cu) with DISTRIBUTED owners.  Slot layout of this kernel:
          // [cta][acc][R / VW row vectors][column][VW].  The rows of a split tile are finished in vectors of VW rows, and
          // vector rv belongs to contributor c_first + rv % (n - 1): every contributor publishes the vectors it does not
          // own (one vector store each) and sums the ones it owns from all the others -- ONE batch of loads per owned
          // vector instead of the single owner's several (a tile of a small layer is cut 6..10 ways, and with 8 or 16 rows
          // the owner's register budget allowed four contributors per polling round).  The last contributor owns
          // nothing: the tile is the FIRST segment of its range, and waiting there would chain the CTAs one behind the
          // other; for every other contributor the tile is the last thing it does.  Rows >= nv carry whatever the
          // accumulator columns held: they are published, summed and reset like real ones (never stored to y).
          constexpr int VB = OWN < 4 ? OWN : 4;
          constexpr int VW = VB / NACC > 0 ? VB / NACC : 1;        // rows per vector (the SwiGLU pair moves two accumulators)
          constexpr int NV = OWN / VW;                             // row vectors per thread
          constexpr int UB = 32 / (NACC * VW) < 8 ? 32 / (NACC * VW) : 8;   // contributors polled together (<= 32 registers)
          const int n_own = c_last - c_first >= 3 ? c_last - c_first : 1;   // four or more contributors: all but the last own
          const int me = cta - c_first;
          // A CTA contributes to at most two split tiles -- the one its range starts in (it is that tile's last
          // contributor) and the one it ends in -- and has one slot area for each.
          auto slot_of = [&](int c, int a, int vi) {
            return p.ll + ((size_t)c * 2 + (c == c_last ? 0 : 1)) * kAreaFloats +
                   (size_t)(((a * R + rb) / VW + vi) * kTileN + nl) * VW;
          };
  #pragma unroll
          for (int vi = 0; vi < NV; vi++) {
            if ((rb / VW + vi) % n_own == me) continue;            // mine to finish (group-uniform)
  #pragma unroll
            for (int a = 0; a < NACC; a++) {
              uint32_t u[VW];
  #pragma unroll
              for (int ee = 0; ee < VW; ee++) {
                u[ee] = __float_as_uint(v[a][vi * VW + ee]);
                if (u[ee] == kLlEmpty) u[ee] = 0x7fffffffu;
              }
              st_slots<VW>(slot_of(cta, a, vi), u);
            }
          }
            if (n_own == 1) {
            // two or three contributors (the wide layers: a tile is cut once or twice): the first one takes every vector of
            // the others in ONE polling round -- vector by vector it would pay a round trip per owned vector, and the
            // others would hold their SM slots for nothing
            if (me == 0) {
              constexpr int UJ = 2;
              uint32_t t[UJ][NACC][NV][VW];
#pragma unroll
              for (int i = 0; i < UJ; i++)
#pragma unroll
                for (int a = 0; a < NACC; a++)
#pragma unroll
                  for (int vi = 0; vi < NV; vi++) {
                    if (c_first + 1 + i <= c_last) {
                      ld_slots<VW>(slot_of(c_first + 1 + i, a, vi), t[i][a][vi]);
                    } else {
#pragma unroll
                      for (int ee = 0; ee < VW; ee++) t[i][a][vi][ee] = 0u;
                    }
                  }
#ifdef FB2_WAIT_TIMEOUT
              const long long t0 = clock64();
#endif
              for (;;) {
                bool waiting = false;
#pragma unroll
                for (int i = 0; i < UJ; i++)
#pragma unroll
                  for (int a = 0; a < NACC; a++)
#pragma unroll
                    for (int vi = 0; vi < NV; vi++)
#pragma unroll
                      for (int ee = 0; ee < VW; ee++) waiting |= t[i][a][vi][ee] == kLlEmpty;
                if (!waiting) break;
#pragma unroll
                for (int i = 0; i < UJ; i++)
#pragma unroll
                  for (int a = 0; a < NACC; a++)
#pragma unroll
                    for (int vi = 0; vi < NV; vi++) {
                      bool miss = false;
#pragma unroll
                      for (int ee = 0; ee < VW; ee++) miss |= t[i][a][vi][ee] == kLlEmpty;
                      if (miss) ld_slots<VW>(slot_of(c_first + 1 + i, a, vi), t[i][a][vi]);
                    }
#ifdef FB2_WAIT_TIMEOUT
                if (clock64() - t0 > 4000000000LL) { printf("fb2: ll slot wait timed out (cta %d tile %d)\n", cta, tau); __trap(); }
#endif
              }
              FB2_TL(9);
#pragma unroll
              for (int i = 0; i < UJ; i++)
#pragma unroll
                for (int a = 0; a < NACC; a++)
#pragma unroll
                  for (int vi = 0; vi < NV; vi++)
                    if (c_first + 1 + i <= c_last) {
#pragma unroll
                      for (int ee = 0; ee < VW; ee++) v[a][vi * VW + ee] += __uint_as_float(t[i][a][vi][ee]);
                      reset_slots<VW>(slot_of(c_first + 1 + i, a, vi));
                    }
#pragma unroll
              for (int jj = 0; jj < OWN; jj++)
                if (jj < nr) store_out(rb + jj, v[0][jj], v[NACC - 1][jj]);
            }
          } else {
#pragma unroll
          for (int vi = 0; vi < NV; vi++) {
            if ((rb / VW + vi) % n_own != me) continue;
  #pragma unroll 1
            for (int c0 = c_first; c0 <= c_last; c0 += UB) {
              uint32_t t[UB][NACC][VW];
  #pragma unroll
              for (int i = 0; i < UB; i++)
  #pragma unroll
                for (int a = 0; a < NACC; a++) {
                  if (c0 + i <= c_last && c0 + i != cta) {
                    ld_slots<VW>(slot_of(c0 + i, a, vi), t[i][a]);
                  } else {
  #pragma unroll
                    for (int ee = 0; ee < VW; ee++) t[i][a][ee] = 0u;
                  }
                }
  #ifdef FB2_WAIT_TIMEOUT
              const long long t0 = clock64();
  #endif
  #ifdef FB2_TRACE
              int rounds = 0;
  #endif
              for (;;) {
                // all slots are tested first and all missing vectors re-requested together
                bool waiting = false;
  #pragma unroll
                for (int i = 0; i < UB; i++)
  #pragma unroll
                  for (int a = 0; a < NACC; a++)
  #pragma unroll
                    for (int ee = 0; ee < VW; ee++) waiting |= t[i][a][ee] == kLlEmpty;
                if (!waiting) break;
  #ifdef FB2_TRACE
                rounds++;
  #endif
  #pragma unroll
                for (int i = 0; i < UB; i++)
  #pragma unroll
                  for (int a = 0; a < NACC; a++) {
                    bool miss = false;
  #pragma unroll
                    for (int ee = 0; ee < VW; ee++) miss |= t[i][a][ee] == kLlEmpty;
                    if (miss) ld_slots<VW>(slot_of(c0 + i, a, vi), t[i][a]);   // only ever true for vectors that were requested
                  }
  #ifdef FB2_WAIT_TIMEOUT
                if (clock64() - t0 > 4000000000LL) { printf("fb2: ll slot wait timed out (cta %d tile %d)\n", cta, tau); __trap(); }
  #endif
              }
              FB2_TL(9);
  #ifdef FB2_TRACE
              FB2_TLV(15, rounds);
  #endif
  #pragma unroll
              for (int i = 0; i < UB; i++)
  #pragma unroll
                for (int a = 0; a < NACC; a++)
                  if (c0 + i <= c_last && c0 + i != cta) {
  #pragma unroll
                    for (int ee = 0; ee < VW; ee++) v[a][vi * VW + ee] += __uint_as_float(t[i][a][ee]);
                    reset_slots<VW>(slot_of(c0 + i, a, vi));         // consumed: empty again for the next launch
                  }
            }
  #pragma unroll
            for (int ee = 0; ee < VW; ee++)
              if (vi * VW + ee < nr) store_out(rb + vi * VW + ee, v[0][vi * VW + ee], v[NACC - 1][vi * VW + ee]);
          }
          }
        } else {
          // ---- flag-in-value exchange (protocol: tc_gemm.cu).  Slot layout of this kernel: [cta][acc][16 / VW][column][VW],
          // so the OWN rows of a thread are ONE vector store for the contributor and one vector load per contributor
          // for the owner.  Rows >= nv carry whatever the accumulator columns held: they are published, summed and
          // reset like real ones (never stored to y), which keeps the vectors whole.
  #ifdef FB2_LL_SCALAR
          constexpr int VW = 1;
  #else
          constexpr int VW = OWN < 4 ? OWN : 4;                    // vector width
  #endif
          constexpr int NV = OWN / VW;                             // vectors per thread, accumulator and contributor
          auto slot_of = [&](int c, int a, int vi) {
            return p.ll + (size_t)c * kLlFloatsPerCta + (size_t)(((a * R + rb) / VW + vi) * kTileN + nl) * VW;
          };
          if (!owner) {
  #pragma unroll
            for (int a = 0; a < NACC; a++)
  #pragma unroll
              for (int vi = 0; vi < NV; vi++) {
                uint32_t u[VW];
  #pragma unroll
                for (int ee = 0; ee < VW; ee++) {
                  u[ee] = __float_as_uint(v[a][vi * VW + ee]);
                  if (u[ee] == kLlEmpty) u[ee] = 0x7fffffffu;
                }
                st_slots<VW>(slot_of(cta, a, vi), u);
              }
          } else {
            constexpr int UB = 32 / (OWN * NACC) < 8 ? 32 / (OWN * NACC) : 8;   // contributors polled together (<= 32 registers: more spills)
  #pragma unroll 1
            for (int c0 = c_first + 1; c0 <= c_last; c0 += UB) {
              uint32_t t[UB][NACC][NV][VW];
  #pragma unroll
              for (int i = 0; i < UB; i++)
  #pragma unroll
                for (int a = 0; a < NACC; a++)
  #pragma unroll
                  for (int vi = 0; vi < NV; vi++) {
                    if (c0 + i <= c_last) {
                      ld_slots<VW>(slot_of(c0 + i, a, vi), t[i][a][vi]);
                    } else {
  #pragma unroll
                      for (int ee = 0; ee < VW; ee++) t[i][a][vi][ee] = 0u;
                    }
                  }
  #ifdef FB2_WAIT_TIMEOUT
              const long long t0 = clock64();
  #endif
  #ifdef FB2_TRACE
              int rounds = 0;
  #endif
              for (;;) {
                // all slots are tested first and all missing vectors re-requested together
                bool waiting = false;
  #pragma unroll
                for (int i = 0; i < UB; i++)
  #pragma unroll
                  for (int a = 0; a < NACC; a++)
  #pragma unroll
                    for (int vi = 0; vi < NV; vi++)
  #pragma unroll
                      for (int ee = 0; ee < VW; ee++) waiting |= t[i][a][vi][ee] == kLlEmpty;
                if (!waiting) break;
  #ifdef FB2_TRACE
                rounds++;
  #endif
  #pragma unroll
                for (int i = 0; i < UB; i++)
  #pragma unroll
                  for (int a = 0; a < NACC; a++)
  #pragma unroll
                    for (int vi = 0; vi < NV; vi++) {
                      bool miss = false;
  #pragma unroll
                      for (int ee = 0; ee < VW; ee++) miss |= t[i][a][vi][ee] == kLlEmpty;
                      if (miss) ld_slots<VW>(slot_of(c0 + i, a, vi), t[i][a][vi]);   // only ever true for vectors that were requested
                    }
  #ifdef FB2_WAIT_TIMEOUT
                if (clock64() - t0 > 4000000000LL) { printf("fb2: ll slot wait timed out (cta %d tile %d)\n", cta, tau); __trap(); }
  #endif
              }
              FB2_TL(9);
  #ifdef FB2_TRACE
              FB2_TLV(15, rounds);
  #endif
  #pragma unroll
              for (int i = 0; i < UB; i++)
  #pragma unroll
                for (int a = 0; a < NACC; a++)
  #pragma unroll
                  for (int vi = 0; vi < NV; vi++)
                    if (c0 + i <= c_last) {
  #pragma unroll
                      for (int ee = 0; ee < VW; ee++) v[a][vi * VW + ee] += __uint_as_float(t[i][a][vi][ee]);
                      reset_slots<VW>(slot_of(c0 + i, a, vi));         // consumed: empty again for the next launch
                    }
            }
  #pragma unroll
            for (int jj = 0; jj < OWN; jj++)
              if (jj < nr) store_out(rb + jj, v[0][jj], v[NACC - 1][jj]);
          }
        }
      } else {
        // ---- counter exchange: partial tile + self-resetting arrival counter
        if (owner) {
          if (threadIdx.x == 0) {
            const int need = c_last - c_first;
#ifdef FB2_WAIT_TIMEOUT
            const long long t0 = clock64();
#endif
            while (ptx::ld_acquire(p.counters + tau) < need) {
#ifdef FB2_WAIT_TIMEOUT
              if (clock64() - t0 > 4000000000LL) { printf("fb2: tile counter wait timed out (cta %d tile %d)\n", cta, tau); __trap(); }
#endif
            }
            p.counters[tau] = 0;                                   // self-reset for the next launch
          }
          __syncwarp();
          ptx::named_bar_sync(3, kEpiThreads);
#pragma unroll 1
          for (int c = c_first + 1; c <= c_last; c++) {
            const float *src = p.partials + (size_t)c * kPart + (size_t)rb * kTileN + nl;
            float t[NACC][OWN];
#pragma unroll
            for (int a = 0; a < NACC; a++)
#pragma unroll
              for (int jj = 0; jj < OWN; jj++) t[a][jj] = jj < nr ? __ldcg(src + (size_t)(a * kNB + jj) * kTileN) : 0.f;
#pragma unroll
            for (int a = 0; a < NACC; a++)
#pragma unroll
              for (int jj = 0; jj < OWN; jj++) v[a][jj] += t[a][jj];
          }
#pragma unroll
          for (int jj = 0; jj < OWN; jj++)
            if (jj < nr) store_out(rb + jj, v[0][jj], v[NACC - 1][jj]);
        } else {
          float *part = p.partials + (size_t)cta * kPart + (size_t)rb * kTileN + nl;
#pragma unroll
          for (int a = 0; a < NACC; a++)
#pragma unroll
            for (int jj = 0; jj < OWN; jj++)
              if (jj < nr) part[(size_t)(a * kNB + jj) * kTileN] = v[a][jj];
          ptx::named_bar_sync(3, kEpiThreads);                     // every partial store of the CTA is issued
          if (threadIdx.x == 0) ptx::red_release_add(p.counters + tau, 1);
          __syncwarp();
        }
      }
    };

    SegInfo prev{0, 0, 0, 0, true};
    bool prev_valid = false;              // the epilogue of segment `prev` is outstanding for this group
    auto drain = [&]() {
      if (pending) {
        wait_done();
        fold();
        ptx::tc_fence_before();
        pending = false;
      }
    };
    while (si.next(sg, upt, n_tiles)) {
      const int n_seg = sg.r1 - sg.r0;
      // Contributors of this tile, resolved while the first stage of the segment is still in flight.  A tile cut
      // by CTA range boundaries has contributors c_first .. c_last (consecutive CTAs); c_first owns it.
      SegInfo cur;
      cur.blk = sg.blk;
      cur.tile = sg.tile;
      cur.whole = (sg.r0 == 0 && sg.r1 == upt);                  // this CTA covers the whole tile
      cur.c_first = cur.c_last = cta;
      if (!cur.whole) {
        const u32 tile_u0 = (u32)(sg.blk * n_tiles + sg.tile) * (u32)upt;
        if constexpr (kRouted) {
          cur.c_first = cta_of_unit(tile_u0, T, (u32)G);
          cur.c_last = cta_of_unit(tile_u0 + upt - 1, T, (u32)G);
        } else {
          cur.c_first = cta_of_unit_fd(tile_u0, p.fd_units, p.fd_grid);
          cur.c_last = cta_of_unit_fd(tile_u0 + upt - 1, p.fd_units, p.fd_grid);
        }
      }
      // units of this segment that fall to this group: local indices lbase + j with parity s
      int j = (int)((lbase ^ (u32)s) & 1u);
      lbase += (u32)n_seg;
#pragma unroll 1
      for (; j < n_seg; j += 2) {
        const uint32_t stage_off = (uint32_t)wp.st * kStageBytes;
        ptx::mbar_wait(k_fw + wp.st * 8, wp.par);
        if (tl_first) { FB2_TL(3); tl_first = false; }
        const uint32_t sz = ptx::lds32(k_sz + stage_off);
        // zero point of this unit: (1024 + z) as fp16 is 0x6400 + z, so 0x8000 | z == -z * 2^-24 is one add
        const uint32_t zh = (sz >> 16) + 0x1C00u;
        const uint32_t zp = zh | (zh << 16);
        const uint32_t z8[8] = {zp, zp, zp, zp, zp, zp, zp, zp};
        if constexpr (kQuarter) {
          // 32 k at a time (8 tcgen05.st of 8 columns): 16 fewer live registers, for the variants that keep 8+ running
          // sums per accumulator
          uint4 w0 = ptx::lds128(k_w + stage_off);
          uint4 w1 = ptx::lds128(k_w + stage_off + 2048u);
          uint32_t lo[8], hi[8];
          convert_quarter(w0, lo, hi);                            // overlaps the MMAs of this group's previous unit
          w0 = ptx::lds128(k_w + stage_off + 4096u);
          if (pending) wait_done();                               // ... which are complete now: A slot s is free
          ptx::tmem_st_x8(k_ta, lo);
          ptx::tmem_st_x8(k_ta + 16, hi);
          ptx::tmem_st_x8(k_z, z8);
          if (pending) fold();                                    // D[s] is read before the next MMAs overwrite it
          convert_quarter(w1, lo, hi);
          w1 = ptx::lds128(k_w + stage_off + 6144u);
          ptx::tmem_st_x8(k_ta + 8, lo);
          ptx::tmem_st_x8(k_ta + 24, hi);
          convert_quarter(w0, lo, hi);
          ptx::tmem_st_x8(k_ta + 32, lo);
          ptx::tmem_st_x8(k_ta + 48, hi);
          convert_quarter(w1, lo, hi);
          // the packed words are in registers: release the weight stage to the HBM stream
          if (lane == 0) ptx::mbar_arrive_after(k_ew + wp.st * 8, lo[7] ^ hi[7]);
          ptx::tmem_st_x8(k_ta + 40, lo);
          ptx::tmem_st_x8(k_ta + 56, hi);
        } else {
          uint4 wa = ptx::lds128(k_w + stage_off);
          uint4 wb = ptx::lds128(k_w + stage_off + 2048u);
          uint32_t lo[16], hi[16];
          convert_half(wa, wb, lo, hi);                           // overlaps the MMAs of this group's previous unit
          wa = ptx::lds128(k_w + stage_off + 4096u);
          wb = ptx::lds128(k_w + stage_off + 6144u);
          if (pending) wait_done();                               // ... which are complete now: A slot s is free
          ptx::tmem_st_x16(k_ta, lo);
          ptx::tmem_st_x16(k_ta + 16, hi);
          ptx::tmem_st_x8(k_z, z8);
          if (pending) fold();                                    // D[s] is read before the next MMAs overwrite it
          convert_half(wa, wb, lo, hi);
          if (lane == 0) ptx::mbar_arrive_after(k_ew + wp.st * 8, lo[15] ^ hi[15]);
          ptx::tmem_st_x16(k_ta + 32, lo);
          ptx::tmem_st_x16(k_ta + 48, hi);
        }
        ptx::tmem_wait_st();
        ptx::tc_fence_before();
        if (lane == 0) ptx::mbar_arrive(k_fa);
        __syncwarp();
        sc_prev = __half2float(__ushort_as_half((unsigned short)(sz & 0xFFFF))) * 16777216.0f;
        pending = true;
        pend_up = NACC == 2 && sg.r0 + j >= groups;
        wp.st += 2;
        if (wp.st >= SW) { wp.st -= SW; wp.par ^= 1u; }
        if constexpr (kDefer) {
          if (prev_valid) {               // the fold above completed the previous segment; this unit's MMAs run meanwhile
            epilogue(prev);
            prev_valid = false;
          }
        }
      }
      FB2_TL(4);
      if constexpr (kDefer) {
        if (prev_valid) {                 // this group had no unit in this segment: finish the older one now
          drain();
          epilogue(prev);
        }
        prev = cur;
        prev_valid = true;
      } else {
        drain();
        epilogue(cur);
      }
    }
    if constexpr (kDefer) {
      if (prev_valid) {
        drain();
        epilogue(prev);
      }
    }
    FB2_TL(6);
    if (push_world(p)) {
      // every participating CTA signals every rank once: its pushed tiles are complete
      ptx::named_bar_sync(1, kDequantWarps * 32);
      if (threadIdx.x == 0) {
        const uint32_t par = *(volatile const uint32_t *)p.push.epoch & 1u;
        __threadfence_system();
        for (int r = 0; r < p.push.world; r++) atomicAdd_system(p.push.flags[r] + par * p.push.world + p.push.rank, 1u);
      }
      __syncwarp();
    }
  }

  // -------------------------------------------------------------------------- teardown
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) ptx::tmem_dealloc<kTmemCols>(tmem_base);
  FB2_TL(7);
}

template <int R, int kMode, bool kSilu, bool kPush>
__global__ void __launch_bounds__(kDkThreads, 2) dk_gemm_kernel(const TcParams p, const __grid_constant__ CUtensorMap xmap) {
  if (!(p.dk_flags & DKF_LATE_TRIGGER)) ptx::launch_dependents();
  dk_body<R, kMode, kSilu, kPush>(p, &xmap);
}

}  // namespace dk
