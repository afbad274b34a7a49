// tc_gemm.cu -- the sm_100a tensor-core W4A16 kernel (tcgen05 / TMEM / bulk-TMA), persistent
// and stream-K scheduled.
//
// Formulation ("swapped operands"): y^T[n, m] = sum_k W[n, k] * x[m, k].  The 128 weight
// columns of a tile sit on the UMMA M side, the activation rows on the UMMA N side
// (N = NB = 16 / 32 / 64), so a decode batch needs no padding to a 128-row MMA tile, and the
// GPTQ packing (8 consecutive k of ONE column per 32-bit word) is already the K-major order
// the A operand wants.
//
// Work decomposition.  A UNIT is one 128(n) x 128(k) weight stage (8704 contiguous bytes:
// packed int4 + the 128 scale/zero pairs of that group) applied to one block of <= NB
// activation rows.  Units are ordered (row block, tile, [gate|up], group); the grid is one
// persistent wave of CTAs (two per SM) and CTA c owns the contiguous unit range
// [c*T/G, (c+1)*T/G) -- every SM streams the same number of bytes whatever N, K and the
// split are (no wave quantisation, no host-chosen split-K factor).  A CTA range that covers
// a whole output tile writes the output directly; ranges that cover part of a tile write an
// fp32 partial, and the LAST contributor to arrive on that tile (self-resetting counter)
// sums the partials in contributor order -- deterministic, and there is no lock workspace
// for the caller to zero (reference: marlin.rs:540-546 has to).
//
// Per CTA:
//   warp 8      producer: per unit one cp.async.bulk (TMA engine, L2 evict-first) for the
//               weight stage, plus 16-byte cp.async row gathers of the activation slice
//               [rows x 128 k] straight into the 128B-swizzled UMMA B layout; both complete on
//               the stage's mbarrier.  The MoE row gather (x[id / top_k]) is this load.
//   warps 0-7   dequant: LDS.128 packed words -> lop3 magic-number int4->fp16, exact
//               (q - zero), one fp16 multiply by the scale [ == f16_rn of the CPU reference's
//               f32 weight ] -> tcgen05.st into a TMEM ring as the UMMA A operand.  Weights
//               never go back to shared memory.
//   warp 9      one elected thread issues tcgen05.mma (kind::f16, M=128, N=NB, K=16, A from
//               TMEM, B from the stage); tcgen05.commit frees TMEM slots and smem stages.
//   warps 0-7   epilogue at the end of each tile segment: tcgen05.ld the fp32 accumulator,
//               then fp16 store (+bias, or SiLU(gate)*up) or partial + last-arriver reduce.
// The two single-thread roles sit in the two highest warps: the issue arbiter favours the
// highest warp id of a scheduler partition.
//
// Programmatic dependent launch: the kernel prefetches its first weight stages (which depend
// on no earlier kernel) before griddepcontrol.wait, so in a chain of GEMMs the weight stream
// of GEMM i+1 starts while GEMM i drains.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"
#include "rowmap.cuh"
#include "sched.cuh"

namespace fb2 {

namespace {

constexpr int kDequantWarps = 8;
constexpr int kWarpW = kDequantWarps;               // 8: weight producer
constexpr int kWarpMma = kDequantWarps + 1;         // 9: MMA issuer
constexpr int kWarpX = kDequantWarps + 2;           // 10: activation producer
constexpr int kThreads = (kDequantWarps + 3) * 32;  // 352
constexpr int kMaxWStages = 16;
constexpr int kMaxXSlots = 4;
constexpr int kMaxASlots = 4;

// Two resource classes.  NB <= 64 ("decode"): two CTAs per SM, 256 TMEM columns each
// ([0,128) accumulators, [128,256) a 2-slot A ring).  NB >= 128 ("prefill"): one CTA per SM,
// all 512 TMEM columns ([0,256) accumulator, [256,512) a 4-slot A ring) and all shared memory:
// one dequantised weight stage then feeds a 128 x NB x 128 MMA block (NB/2 cycles per MMA), so the
// tensor pipe, not the int4->fp16 conversion, is the busy unit.
template <int NB>
struct StageCfg {
  static constexpr bool kWide = NB >= 128;
  static constexpr int kCtasPerSm = kWide ? 1 : 2;
  static constexpr int kDequantWarps = kWide ? 16 : 8;          // two teams of eight in the one-CTA-per-SM class
  static constexpr int kCtaThreads = (kDequantWarps + 3) * 32;
  // Registers are allocated per warp in groups of four warps: a 10-warp CTA occupies the register
  // space of 12, so two CTAs are co-resident only below 65536 / (2 * 12 * 32) = 85 registers per
  // thread.  Declaring 384 threads to the compiler makes it hold that budget (80); the launch still
  // uses kThreads.  (At 96 registers the occupancy query answers ONE CTA per SM.)
#ifdef FB2_REGS96
  static constexpr int kBoundThreads = kCtaThreads;
#else
  static constexpr int kBoundThreads = kWide ? kCtaThreads : 384;
#endif
  static constexpr int kTmemCols = kWide ? 512 : 256;
  static constexpr int kTmemABase = kWide ? 256 : 128;
  static constexpr int kASlots = kWide ? 4 : 2;                               // slots of one 128-k group (64 columns)
  static constexpr int kSmemBudget = kWide ? 225 * 1024 : 108 * 1024;         // per CTA
  // Two rings.  The WEIGHT ring (HBM stream, ~3 us loaded latency) is deep and a stage is released
  // as soon as the dequant warps have read it into registers; the ACTIVATION ring (L2 hits) is
  // shallow and a slot lives until the MMAs that read it have completed.  With one combined ring a
  // weight stage stayed occupied until its MMA committed, and the bytes in flight per SM -- not
  // the dequant rate -- bounded the stream at ~4 TB/s (the same kernel on L2-resident weights ran
  // 2.9x faster).
  static constexpr int kXBytes = NB * 256;                                    // [2 atoms][NB][128 B]
  static constexpr int kXSlots = kWide ? (NB >= 256 ? 2 : 3) : (NB <= 16 ? 4 : 3);
  static constexpr int kWBase = kXSlots * kXBytes;                            // 1024-aligned
  static constexpr int kWStagesRaw = (kSmemBudget - 1024 - kWBase) / kStageBytes;
  static constexpr int kWStages = kWStagesRaw > kMaxWStages ? kMaxWStages : kWStagesRaw;
  static constexpr int kBarBase = (kWBase + kWStages * kStageBytes + 127) & ~127;
  static constexpr int kSmemBytes = kBarBase + 1024;                          // + barriers / misc
};

struct TcParams {
  WeightRef w;
  RowMap rm;
  const __half *x;
  __half *y;
  float *partials;        // [grid][NACC * NB * 128]: a CTA publishes at most one partial tile
  int32_t *counters;      // [row blocks * n_tiles], zero between launches
  uint32_t *ll;           // [ll_ctas][2 acc][kLlRows][128] flag-in-value partial slots (all kLlEmpty between launches)
  int ll_ctas;            // CTAs the ll area serves (0: exchange disabled)
  int n_tiles;            // output tiles per row block (N/128, or I/128 with the SiLU epilogue)
  int groups;             // K / 128
  int tiles_per_expert;   // ld / 128
  int ld_out;             // output row stride
  int x_tma;              // activations come through the 2-D tensor map (dense / dispatch rows)
  int whole_tiles;        // dense / dispatch rows: CTA ranges are cut on tile boundaries (host decision)
  sched::FastDiv fd_grid, fd_upt, fd_tiles, fd_units;   // decode kernel: divisions by launch constants (sched.cuh)
  int dk_flags;           // decode kernel: DKF_* (tc_decode.cuh)
  const int32_t *out_map; // decode kernel, routed: output column c is stored at out_map[c] (the act-order column order of
                          // the consumer: no gather launch between the two grouped GEMMs of an FFN), or null
  int w_stages;           // decode kernel: weight ring depth
  int l2_prefetch;        // decode kernel: every CTA requests its whole range into L2 at entry
  PushTargets push;       // row-parallel linear: output tiles are pushed to every rank's receive slot
};

// barrier slots (8 bytes each)
enum { BAR_FULL_W = 0, BAR_EMPTY_W = kMaxWStages, BAR_FULL_X = 2 * kMaxWStages,
       BAR_EMPTY_X = BAR_FULL_X + kMaxXSlots, BAR_FULL_A = BAR_EMPTY_X + kMaxXSlots,
       BAR_EMPTY_A = BAR_FULL_A + 2 * kMaxASlots, BAR_ACC = BAR_EMPTY_A + 2 * kMaxASlots,
       BAR_ACC_EMPTY = BAR_ACC + 1, BAR_COUNT = BAR_ACC_EMPTY + 1 };
static_assert(BAR_COUNT * 8 <= 512, "barrier area too small");

__device__ __forceinline__ float silu_mul(float g, float u) { return (g / (1.0f + expf(-g))) * u; }

// Split tiles of blocks with at most kLlRows real rows (decode batches of 1-8, MoE blocks with a few
// pairs) exchange their partial sums through FLAG-IN-VALUE slots: a slot holds kLlEmpty until the
// contributor's plain 4-byte store lands, the tile owner polls the slots of all contributors at once,
// sums them in contributor order and writes kLlEmpty back.  No counter, no release fence, no CTA
// barrier on either side, and only the real rows travel.  kLlEmpty is a NaN bit pattern arithmetic
// never produces (hardware NaNs are 0x7fffffff); a contributor maps it to that canonical NaN anyway.
constexpr int kLlRows = 8;             // blocks with at most this many real rows use the exchange (one polling
                                       // round per warp set; two rounds -- 16 rows -- measured slower than the
                                       // counter exchange)
constexpr int kLlGroup = 4;            // rows exchanged per polling round
constexpr uint32_t kLlEmpty = 0xffffffffu;
constexpr int kLlFloatsPerCta = 2 * kLlRows * kTileN;
static_assert((size_t)kLlMaxCtas * kLlFloatsPerCta * sizeof(uint32_t) == kLlBytes, "ll area size");

// Dequantise one packed word (tile_nibble_k: even nibbles = "lo class", odd = "hi class") into 4 half2
// registers: lo[0] = nibbles (0,4), lo[1] = nibbles (2,6) -- consecutive A columns of the lo k-steps --
// and hi[0] = nibbles (1,5), hi[1] = nibbles (3,7) of the hi k-steps.
__device__ __forceinline__ void dequant_word(uint32_t q, __half2 zlo, __half2 zhi, __half2 s2,
                                             uint32_t *lo, uint32_t *hi) {
  const uint32_t kMagic = 0x64006400u;                 // half2(1024, 1024)
  const __half2 k16th = __float2half2_rn(0.0625f);
  uint32_t lo0 = ptx::lop3_and_or(q, 0x000f000fu, kMagic);   // 1024 + q
  uint32_t hi0 = ptx::lop3_and_or(q, 0x00f000f0u, kMagic);   // 1024 + 16 q
  uint32_t q2 = q >> 8;
  uint32_t lo1 = ptx::lop3_and_or(q2, 0x000f000fu, kMagic);
  uint32_t hi1 = ptx::lop3_and_or(q2, 0x00f000f0u, kMagic);
  __half2 v0 = __hsub2(*reinterpret_cast<__half2 *>(&lo0), zlo);          // q - z, exact
  __half2 v1 = __hfma2(*reinterpret_cast<__half2 *>(&hi0), k16th, zhi);   // q - z, exact
  __half2 v2 = __hsub2(*reinterpret_cast<__half2 *>(&lo1), zlo);
  __half2 v3 = __hfma2(*reinterpret_cast<__half2 *>(&hi1), k16th, zhi);
  v0 = __hmul2(v0, s2);  // one rounding: == f16_rn((q - z) * s)
  v1 = __hmul2(v1, s2);
  v2 = __hmul2(v2, s2);
  v3 = __hmul2(v3, s2);
  lo[0] = *reinterpret_cast<uint32_t *>(&v0);
  hi[0] = *reinterpret_cast<uint32_t *>(&v1);
  lo[1] = *reinterpret_cast<uint32_t *>(&v2);
  hi[1] = *reinterpret_cast<uint32_t *>(&v3);
}

// CH (8 or 16) consecutive fp32 accumulator columns of this thread's lane
template <int CH>
__device__ __forceinline__ void load_acc_chunk(uint32_t taddr, float *v) {
  static_assert(CH == 8 || CH == 16, "chunk width");
  if constexpr (CH == 8) {
    uint32_t r[8];
    ptx::tmem_ld_x8(taddr, r);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
  } else {
    uint32_t r[16];
    ptx::tmem_ld_x16(taddr, r);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
  }
}

template <int NB>
__device__ __forceinline__ void load_acc(uint32_t taddr, float *v) {
  // NB/2 consecutive fp32 columns of this thread's lane
  if constexpr (NB == 16) {
    uint32_t r[8];
    ptx::tmem_ld_x8(taddr, r);
    ptx::tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
  } else {
#pragma unroll
    for (int c = 0; c < NB / 32; c++) {
      uint32_t r[16];
      ptx::tmem_ld_x16(taddr + c * 16, r);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; i++) v[c * 16 + i] = __uint_as_float(r[i]);
    }
  }
}

// stream-K bookkeeping (range_start, whole_tile_mode, tile_range_start, cta_of_unit, Seg, SegIter): sched.cuh
using namespace sched;

// ring position: stage index + phase parity, advanced without division
struct RingPos {
  int st;
  uint32_t par;
  __device__ __forceinline__ void advance(int stages) {
    if (++st == stages) { st = 0; par ^= 1u; }
  }
};

#ifdef FB2_TRACE
__device__ long long *g_trace = nullptr;      // pinned host memory: [cta][32] progress markers (debug builds only)
#if FB2_TRACE >= 2   // per-unit progress markers (slow: every marker is a posted write to host memory)
#define FB2_MARK(v) do { if (g_trace && (threadIdx.x & 31) == 0) { \
    *((volatile long long *)g_trace + (size_t)blockIdx.x * 32 + (threadIdx.x >> 5)) = (long long)(v); } } while (0)
#else
#define FB2_MARK(v) do { } while (0)
#endif
#define FB2_MARK2(slot, v) do { if (g_trace && (threadIdx.x & 31) == 0) { \
    *((volatile long long *)g_trace + (size_t)blockIdx.x * 32 + 16 + (slot)) = (long long)(v); } } while (0)
__device__ __forceinline__ long long fb2_gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define FB2_TSTAMP(slot) do { if (g_trace && threadIdx.x == 0) { \
    *((volatile long long *)g_trace + (size_t)blockIdx.x * 32 + 16 + (slot)) = fb2_gtime(); } } while (0)
// timeline of one GEMM launch: [8 launches][296 CTAs][32] after the fused-kernel area
#define FB2_TL(slot) do { if (g_trace && threadIdx.x == 0) { \
    *((volatile long long *)g_trace + (size_t)(296 + (((p.x_tma >> 16) & 7) * 296 + blockIdx.x)) * 32 + (slot)) = fb2_gtime(); } } while (0)
#define FB2_TLV(slot, v) do { if (g_trace && threadIdx.x == 0) { \
    *((volatile long long *)g_trace + (size_t)(296 + (((p.x_tma >> 16) & 7) * 296 + blockIdx.x)) * 32 + (slot)) = (long long)(v); } } while (0)
// same, by lane 0 of whichever warp executes it
#define FB2_TLW(slot) do { if (g_trace && (threadIdx.x & 31) == 0) { \
    *((volatile long long *)g_trace + (size_t)(296 + (((p.x_tma >> 16) & 7) * 296 + blockIdx.x)) * 32 + (slot)) = fb2_gtime(); } } while (0)
#else
#define FB2_MARK(v) do { } while (0)
#define FB2_MARK2(slot, v) do { } while (0)
#define FB2_TSTAMP(slot) do { } while (0)
#define FB2_TL(slot) do { } while (0)
#define FB2_TLV(slot, v) do { } while (0)
#define FB2_TLW(slot) do { } while (0)
#endif

// kPush instantiations (row-parallel decode GEMMs with the fused all-reduce) are separate kernels: the
// push branches in the epilogue cost the prefill variant 13 % when they were a run-time test.
#define push_world(p) (kPush ? (p).push.world : 0)

// kPair (prefill row blocks only): the grid is launched in CLUSTERS OF TWO.  The two CTAs of a cluster walk the same unit
// sequence on two ADJACENT weight tiles, so they need the same activation slices at the same time: each loads half of
// every slice (NB / 2 rows) and multicasts it to both, which halves the L2 -> SM traffic of the operand that bounded
// the tensor pipe at 51 % in round 1 (64 KB of activations per 8.7 KB weight stage, re-read by every tile).
template <int NB, bool kSilu, bool kPush = false, bool kPair = false>
__device__ __forceinline__ void tc_body(const TcParams &p, const DispatchTable *tbl, const CUtensorMap *xmap) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using SC = StageCfg<NB>;
  constexpr int NACC = kSilu ? 2 : 1;
  constexpr int kTmemCols = SC::kTmemCols, kTmemABase = SC::kTmemABase, kASlots = SC::kASlots;
  // Dequant warps come in TEAMS of eight (four TMEM lane quarters x two 64-k halves).  The prefill classes run two
  // teams that take alternate units, so two units are in flight per CTA (round 1's one-team loop was latency-bound:
  // 30 % issue utilisation at 51 % tensor pipe) -- 19 warps still fit the register file of the one-CTA-per-SM class.
  constexpr int DW = SC::kDequantWarps, kTeams = DW / 8;
  constexpr int kWarpW = DW, kWarpMma = DW + 1, kWarpX = DW + 2, kThreads = (DW + 3) * 32;
  constexpr int kDequantWarps = DW;
  constexpr int HALF = NB / (DW / 4);                            // rows each set of four epilogue warps handles
  constexpr int SW = SC::kWStages, SX = SC::kXSlots;
  static_assert(SW >= 3 && SX >= 2, "smem rings too shallow");
  static_assert(NACC * NB <= kTmemABase, "accumulators overlap the A ring");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  FB2_TL(0);
  static_assert(!kPair || (NB >= 128 && !kSilu && !kPush), "paired CTAs: prefill row blocks only");
  // Scheduling view: in paired mode the stream-K bookkeeping runs over PAIRS (pair index, tile pairs); this CTA's
  // physical tile is 2 * tile + crank and its scratch slots are indexed by the physical CTA.
  const int crank = kPair ? (int)(blockIdx.x & 1u) : 0;
  int G = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;       // scheduling CTAs that take part (see below)
  const int cta = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  auto phys_cta = [&](int c) { return kPair ? 2 * c + crank : c; };
  auto phys_tile = [&](int t) { return kPair ? 2 * t + crank : t; };
  const int groups = p.groups, n_tiles = kPair ? p.n_tiles / 2 : p.n_tiles;
  const int upt = NACC * groups;                                 // units per output tile
  // Routed mode learns its block count from the align kernel (an earlier kernel): its unit range
  // is computed after the grid dependency, i.e. after the dependency-free part of the setup.
  const bool routed = p.rm.mode == ROWS_ROUTED;
  u32 T = 0, u0 = 0, u1 = 0;
  if (!routed) {
    T = (u32)rowmap_num_blocks(p.rm, NB) * (u32)n_tiles * (u32)upt;
    if (p.whole_tiles) {                                         // host decided (it sized the grid)
      u0 = tile_range_start(cta, T / upt, G, upt);
      u1 = tile_range_start(cta + 1, T / upt, G, upt);
    } else {
      u0 = range_start(cta, T, G);
      u1 = range_start(cta + 1, T, G);
    }
    if (u0 >= u1) return;                                        // CTA-uniform, before any barrier
  }

  const uint32_t smem_base = ptx::smem_u32(smem);
  const uint32_t bars = smem_base + SC::kBarBase;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  uint32_t *s_tmem = reinterpret_cast<uint32_t *>(smem + SC::kBarBase + 512);
  const size_t tile_stride = (size_t)groups * kStageBytes;

  // ------------------------------------------------------------------ one-time setup
  // Each producer owns the barriers of its ring and starts at once; TMEM allocation and the other
  // barriers proceed in parallel on the MMA warp.  Named barrier 2: the producers only ARRIVE (they
  // never wait for the rest of the CTA), the nine other warps sync.
  if (warp == kWarpW) {
    if (lane == 0) {
      for (int i = 0; i < SW; i++) {
        ptx::mbar_init(bar(BAR_FULL_W + i), 1);                  // the producer's arrive.expect_tx
        ptx::mbar_init(bar(BAR_EMPTY_W + i), 8);                 // read into registers by the 8 warps of one team
      }
      ptx::mbar_fence_init();
    }
    __syncwarp();
    ptx::named_bar_arrive(2, kThreads);
  } else if (warp == kWarpX) {
    if (lane == 0) {
      for (int i = 0; i < SX; i++) {
        // one arrive.expect_tx by the producer (+ 32 cp.async arrivals when rows are gathered)
        ptx::mbar_init(bar(BAR_FULL_X + i), (p.x_tma & 1) ? 1 : 1 + 32);
        ptx::mbar_init(bar(BAR_EMPTY_X + i), kPair ? 2 : 1);     // the MMA commit (paired: of both CTAs, which both write it)
      }
      ptx::mbar_fence_init();
    }
    __syncwarp();
    ptx::named_bar_arrive(2, kThreads);
  } else if (warp == kWarpMma) {
    if (lane == 0) {
      for (int i = 0; i < 2 * kASlots; i++) {
        ptx::mbar_init(bar(BAR_FULL_A + i), 4);
        ptx::mbar_init(bar(BAR_EMPTY_A + i), 1);
      }
      ptx::mbar_init(bar(BAR_ACC), 1);
      ptx::mbar_init(bar(BAR_ACC_EMPTY), kDequantWarps);
      ptx::mbar_fence_init();
    }
    __syncwarp();
    ptx::tmem_alloc<kTmemCols>(ptx::smem_u32(s_tmem));
  }

  if (routed) {
    ptx::grid_dependency_wait();
    T = (u32)rowmap_num_blocks(p.rm, NB) * (u32)n_tiles * (u32)upt;
    // The grid was sized for the worst-case block count; with fewer units only the first
    // max(1, T/2) CTAs take part, so that every participating CTA owns a non-empty range and the
    // contributors of a tile are exactly the consecutive CTAs c_first .. c_last.
    const bool whole_tiles = whole_tile_mode(upt, T / (u32)upt, G);
    const u32 cap = whole_tiles ? T / (u32)upt : T / 2;             // whole tiles | >= 2 units per CTA
    const u32 g_eff = cap > 1 ? cap : 1;
    if (g_eff < (u32)G) G = (int)g_eff;
    if (cta < G) {
      if (whole_tiles) {
        u0 = tile_range_start(cta, T / upt, G, upt);
        u1 = tile_range_start(cta + 1, T / upt, G, upt);
      } else {
        u0 = range_start(cta, T, G);
        u1 = range_start(cta + 1, T, G);
      }
    }
  }
  const bool idle = u0 >= u1;                                    // routed launch with fewer units than CTAs


  const uint64_t pol = ptx::policy_evict_first();
  auto weight_src = [&](int blk, int tile, int phase, int g) {
      const int expert = rowmap_block_expert(p.rm, tbl, blk);
      return p.w.tiles +
             ((size_t)expert * p.tiles_per_expert + p.w.col0 / kTileN + phys_tile(tile) + (size_t)phase * p.n_tiles) *
               tile_stride + (size_t)g * kStageBytes;
  };
  uint32_t tmem_base = 0;
  if (warp != kWarpW && warp != kWarpX) {
    ptx::tc_fence_before();
    ptx::named_bar_sync(2, kThreads);
    ptx::tc_fence_after();
    tmem_base = *s_tmem;
  }
  if constexpr (kPair) {
    // the peer multicasts into this CTA's slots and arrives on its barriers: both CTAs' barriers must be initialised
    // before either starts (and, at the end, neither may exit while the other can still reach it)
    __syncthreads();
    ptx::cluster_sync_all();
  }
  FB2_TL(1);
  FB2_TLV(8, u1 - u0);
#ifdef FB2_TRACE
  { uint32_t smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); FB2_TLV(10, smid); }
#endif

  if (idle) {
    // nothing to stream; fall through to the teardown (TMEM was allocated)
  } else if (warp == kWarpW) {
    // ========================================================== weight producer (HBM stream)
    // Weights depend on no earlier kernel: under programmatic dependent launch this warp streams
    // ahead of griddepcontrol.wait (routed launches waited above: their expert ids are an input).
    RingPos wp{0, 0};
    int it = 0;
    SegIter si;
    si.init(u0, u1, upt, n_tiles);
    Seg sg;
    while (si.next(sg, upt, n_tiles)) {
#pragma unroll 1
      for (int a = 0; a < NACC; a++) {
        const int g0 = max(sg.r0, a * groups) - a * groups, g1 = min(sg.r1, (a + 1) * groups) - a * groups;
        if (g0 >= g1) continue;
        const uint8_t *wsrc = weight_src(sg.blk, sg.tile, a, g0);
#pragma unroll 1
        for (int g = g0; g < g1; g++, it++, wsrc += kStageBytes, wp.advance(SW)) {
          if (it >= SW) ptx::mbar_wait(bar(BAR_EMPTY_W + wp.st), wp.par ^ 1u);
          if (ptx::elect_one()) {
            const uint32_t full = bar(BAR_FULL_W + wp.st);
            ptx::mbar_arrive_expect_tx(full, (uint32_t)kStageBytes);
            ptx::bulk_g2s_hint(smem_base + SC::kWBase + wp.st * kStageBytes, wsrc, kStageBytes, full, pol);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kWarpX) {
    // ====================================================== activation producer (L2 hits)
    // the descriptor is a kernel parameter: fetch it before the grid dependency resolves
    if ((p.x_tma & 1) && lane == 0) ptx::prefetch_tensormap(xmap);
    ptx::grid_dependency_wait();
    // rows of the current block: lane i holds row i (and i + 32 for NB == 64)
    int cur_blk = -1, nchunks = 0, arow0 = -1, arow1 = -1, row0 = 0;
    RingPos xp{0, 0};
    int it = 0;
    SegIter si;
    si.init(u0, u1, upt, n_tiles);
    Seg sg;
    while (si.next(sg, upt, n_tiles)) {
      if (sg.blk != cur_blk && (p.x_tma & 1)) {
        cur_blk = sg.blk;
        int ar, cr;
        rowmap_resolve(p.rm, tbl, NB, sg.blk, 0, ar, cr);        // dense / dispatch: rows are contiguous
        row0 = ar;
      } else if (sg.blk != cur_blk) {
        cur_blk = sg.blk;
        int ar, cr;
        arow0 = (lane < NB && lane < 32 && rowmap_resolve(p.rm, tbl, NB, sg.blk, lane, ar, cr)) ? ar : -1;
        // valid rows form a prefix of the block (moe_align_block_size puts real ids first)
        int nvalid = __popc(__ballot_sync(0xffffffffu, arow0 >= 0));
        if (NB > 32) {
          arow1 = (lane + 32 < NB && rowmap_resolve(p.rm, tbl, NB, sg.blk, lane + 32, ar, cr)) ? ar : -1;
          nvalid += __popc(__ballot_sync(0xffffffffu, arow1 >= 0));
        }
        nchunks = nvalid * 16;
      }
#pragma unroll 1
      for (int a = 0; a < NACC; a++) {
        const int g0 = max(sg.r0, a * groups) - a * groups, g1 = min(sg.r1, (a + 1) * groups) - a * groups;
        if (g0 >= g1) continue;
        const __half *xg = p.x + (size_t)g0 * kGroup;
#pragma unroll 1
        for (int g = g0; g < g1; g++, it++, xg += kGroup, xp.advance(SX)) {
          const uint32_t slot_addr = smem_base + xp.st * SC::kXBytes;
          const uint32_t full = bar(BAR_FULL_X + xp.st);
          if (it >= SX) ptx::mbar_wait(bar(BAR_EMPTY_X + xp.st), xp.par ^ 1u);
          if (ptx::elect_one()) {
            ptx::mbar_arrive_expect_tx(full, (p.x_tma & 1) ? (uint32_t)SC::kXBytes : 0u);
            if (p.x_tma & 1) {
              // [NB rows x 64 k] boxes, hardware 128B swizzle; rows past the tensor end read as zero
              if constexpr (kPair) {
                // this CTA's half of the rows, delivered to both CTAs of the pair (the tensor map's box is NB / 2 rows)
                const uint32_t half_off = (uint32_t)crank * (NB / 2) * 128u;
                ptx::tma_load_2d_multicast(slot_addr + half_off, xmap, g * kGroup, row0 + crank * (NB / 2), full, (uint16_t)3);
                ptx::tma_load_2d_multicast(slot_addr + NB * 128 + half_off, xmap, g * kGroup + 64, row0 + crank * (NB / 2), full,
                                           (uint16_t)3);
              } else {
                ptx::tma_load_2d(slot_addr, xmap, g * kGroup, row0, full);
                ptx::tma_load_2d(slot_addr + NB * 128, xmap, g * kGroup + 64, row0, full);
              }
            }
          }
          __syncwarp();
          if (!(p.x_tma & 1)) {
            // gathered activation slice: rows [0, nvalid) x 128 k, 16 chunks of 16 B per row
            // (this IS the MoE gather x[id / top_k], dispatch.rs:1763)
            for (int ch0 = 0; ch0 < nchunks; ch0 += 32) {         // warp-uniform trip count
              const int ch = ch0 + lane;
              const int i = ch >> 4, c8 = ch & 15;
              const int ar = __shfl_sync(0xffffffffu, (NB > 32 && i >= 32) ? arow1 : arow0, i & 31);
              if (ch < nchunks) {
                const uint32_t dst = slot_addr + (uint32_t)(c8 >> 3) * (NB * 128) + (uint32_t)i * 128 +
                                     (uint32_t)(((c8 & 7) ^ (i & 7)) << 4);
                ptx::cp_async_16(dst, xg + (size_t)ar * p.w.k + c8 * 8);
              }
            }
            ptx::cp_async_mbar_arrive_noinc(full);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == kWarpMma) {
    // ==================================================================== MMA issuer
    // Warp-uniform branch + elect.sync predicate: ptxas emits straight-line UTCHMMA with
    // uniform-register operands (no per-instruction election loop).
    constexpr uint32_t idesc = ptx::umma_idesc_f16(128, NB);
    const uint64_t b_desc_ring = ptx::umma_desc_kmajor_sw128(smem_base);
    RingPos xp{0, 0}, ap{0, 0};
    int seg = 0;
    SegIter si;
    si.init(u0, u1, upt, n_tiles);
    Seg sg;
    while (si.next(sg, upt, n_tiles)) {
      // this segment overwrites the accumulators: the previous epilogue must have read them
      if (seg > 0) ptx::mbar_wait(bar(BAR_ACC_EMPTY), (seg - 1) & 1);
#pragma unroll 1
      for (int a = 0; a < NACC; a++) {
        const int g0 = max(sg.r0, a * groups), g1 = min(sg.r1, (a + 1) * groups);
        const uint32_t d_tmem = tmem_base + a * NB;
#pragma unroll 1
        for (int g = g0; g < g1; g++, xp.advance(SX), ap.advance(kASlots)) {
          ptx::mbar_wait(bar(BAR_FULL_X + xp.st), xp.par);        // activation slice landed
          // gathered rows were written through the generic proxy (cp.async); TMA boxes already are async-proxy writes
          if (!(p.x_tma & 1)) ptx::fence_proxy_async_smem();
          const uint64_t b_desc0 = b_desc_ring + (uint64_t)((xp.st * SC::kXBytes) >> 4);
          const uint32_t acc_flag = g != g0 ? 1u : 0u;            // first unit of a phase zero-inits
#pragma unroll
          for (int h = 0; h < 2; h++) {
            ptx::mbar_wait(bar(BAR_FULL_A + ap.st * 2 + h), ap.par);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint64_t b_d = b_desc0 + (uint64_t)((h * (NB * 128)) >> 4);
              const uint32_t a_t = tmem_base + kTmemABase + ap.st * 64 + h * 32;
#pragma unroll
              for (int kk = 0; kk < 4; kk++)
                ptx::umma_f16_ts(d_tmem, a_t + kk * 8, b_d + (uint64_t)(kk * 2), idesc,
                                 (h | kk) != 0 ? 1u : acc_flag);
              ptx::tc_commit(bar(BAR_EMPTY_A + ap.st * 2 + h));
              if (h == 1) {                                          // activation slot consumed
                if constexpr (kPair) ptx::tc_commit_multicast(bar(BAR_EMPTY_X + xp.st), (uint16_t)3);
                else ptx::tc_commit(bar(BAR_EMPTY_X + xp.st));
              }
            }
            __syncwarp();
          }
        }
      }
      if (ptx::elect_one()) ptx::tc_commit(bar(BAR_ACC));
      __syncwarp();
      seg++;
    }
  } else {
    // ======================================================== dequant + epilogue warps
    const int q = warp & 3;               // TMEM lane quarter this warp may access
    const int h = (warp >> 2) & 1;        // which 64-k half of every group this warp owns
    const int team = warp >> 3;           // which units this warp converts: local unit index % kTeams
    const int rset = warp >> 2;           // which rows of the block this warp finishes in the epilogue
    const int nl = q * 32 + lane;         // weight column inside the tile == TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t w_off = SC::kWBase + (uint32_t)h * 4096u + (uint32_t)nl * 16u;    // chunks 2h, 2h+1
    const uint32_t sz_off = SC::kWBase + kStageWBytes + (uint32_t)nl * 4u;
    const uint32_t a_addr = lane_base + kTmemABase + h * 32;
    // Loop invariants passed through a self-shuffle (the identity, but one the assembler cannot see
    // through): ptxas otherwise re-derives them from %tid and the shared window base in every iteration
    // to save registers -- 15 of the loop's 171 instructions.
    uint32_t k_w = smem_base + w_off, k_sz = smem_base + sz_off, k_ta = a_addr, k_bf = bar(BAR_FULL_A + h),
             k_be = bar(BAR_EMPTY_A + h), k_fw = bar(BAR_FULL_W), k_ew = bar(BAR_EMPTY_W);
    k_w = __shfl_sync(0xffffffffu, k_w, lane); k_sz = __shfl_sync(0xffffffffu, k_sz, lane);
    k_ta = __shfl_sync(0xffffffffu, k_ta, lane); k_bf = __shfl_sync(0xffffffffu, k_bf, lane);
    k_be = __shfl_sync(0xffffffffu, k_be, lane); k_fw = __shfl_sync(0xffffffffu, k_fw, lane);
    k_ew = __shfl_sync(0xffffffffu, k_ew, lane);
    RingPos wp{team, 0}, ap{team, 0};     // this team's next unit is local unit `team`
    u32 lunit = 0;                        // local index of the next unit of the range (all teams count all units)
    int seg = 0;
    ptx::grid_dependency_wait();          // before the first global write / scratch access
    FB2_TL(2);
    bool tl_first = true;
    SegIter si;
    si.init(u0, u1, upt, n_tiles);
    Seg sg;
    while (si.next(sg, upt, n_tiles)) {
      // One unit: a weight stage through the dequant maths into an A slot.  S >= 0 compiles the A slot in.
      auto unit = [&](auto slot_c) {
        constexpr int S = decltype(slot_c)::value;
        const uint32_t a_slot = S < 0 ? (uint32_t)ap.st : (uint32_t)S;
        const uint32_t stage_off = wp.st * kStageBytes;
        ptx::mbar_wait(k_fw + wp.st * 8, wp.par);
        if (tl_first) { FB2_TL(3); tl_first = false; }
        const uint32_t sz = ptx::lds32(k_sz + stage_off);
        const uint4 w0 = ptx::lds128(k_w + stage_off);
        const uint4 w1 = ptx::lds128(k_w + stage_off + 2048u);

        const __half s = __ushort_as_half((unsigned short)(sz & 0xFFFF));
        const __half zl = __ushort_as_half((unsigned short)(sz >> 16));   // 1024 + z
        const __half2 s2 = __half2half2(s);
        const __half2 zlo = __half2half2(zl);
        const __half2 zhi = __hsub2(__float2half2_rn(960.0f), zlo);       // -(64 + z), exact

        uint32_t r0[16], r1[16];                                  // A columns [0,16) = lo class, [16,32) = hi class
        dequant_word(w0.x, zlo, zhi, s2, r0 + 0, r1 + 0);
        dequant_word(w0.y, zlo, zhi, s2, r0 + 2, r1 + 2);
        dequant_word(w0.z, zlo, zhi, s2, r0 + 4, r1 + 4);
        dequant_word(w0.w, zlo, zhi, s2, r0 + 6, r1 + 6);
        dequant_word(w1.x, zlo, zhi, s2, r0 + 8, r1 + 8);
        dequant_word(w1.y, zlo, zhi, s2, r0 + 10, r1 + 10);
        dequant_word(w1.z, zlo, zhi, s2, r0 + 12, r1 + 12);
        dequant_word(w1.w, zlo, zhi, s2, r0 + 14, r1 + 14);
        // the packed words are in registers (consumed, not merely requested): release the weight
        // stage to the HBM stream now -- it does not wait for the MMA.  The arrive is predicated (no
        // divergence, so no reconvergence point is needed) and tied to the last dequantised words.
        if (lane == 0) ptx::mbar_arrive_after(k_ew + wp.st * 8, r0[15] ^ r1[15]);

        ptx::mbar_wait(k_be + a_slot * 16, ap.par ^ 1u);
        ptx::tc_fence_after();
        const uint32_t ta = k_ta + a_slot * 64;
        ptx::tmem_st_x16(ta, r0);
        ptx::tmem_st_x16(ta + 16, r1);
        ptx::tmem_wait_st();                                      // warp-collective: every lane's stores are done
        ptx::tc_fence_before();
        if (lane == 0) ptx::mbar_arrive(k_bf + a_slot * 16);
        wp.advance(SW);
      };
      if constexpr (kASlots == 2) {
        // Two A slots: two copies of the body alternate, each with its slot (barrier and TMEM addresses)
        // compiled in; the ring position costs one parity flip per pair instead of ten instructions a unit.
        int n = sg.r1 - sg.r0;
#pragma unroll 1
        while (n > 0) {
          if (ap.st == 0) {
            unit(std::integral_constant<int, 0>{});
            ap.st = 1;
            if (--n == 0) break;
          }
          unit(std::integral_constant<int, 1>{});
          ap.st = 0;
          ap.par ^= 1u;
          --n;
        }
      } else {
#pragma unroll 1
        for (int n = sg.r1 - sg.r0; n > 0; n--, lunit++) {
          if (kTeams > 1 && (lunit % (u32)kTeams) != (u32)team) continue;   // the other team's unit
          unit(std::integral_constant<int, -1>{});                          // (advances the weight ring once)
#pragma unroll
          for (int t = 1; t < kTeams; t++) wp.advance(SW);
#pragma unroll
          for (int t = 0; t < kTeams; t++) ap.advance(kASlots);
        }
      }

      // ------------------------------------------------ epilogue of this tile segment
      const int r_begin = sg.r0, r_end = sg.r1;                  // [r_begin, r_end) of [0, upt)
      const int blk = sg.blk, tile = phys_tile(sg.tile), tau = sg.blk * p.n_tiles + tile;   // physical tile / counter index
      FB2_TL(4);
      // Everything that does not need the accumulator is resolved while the last MMAs drain.
      // Real rows of this block: they form a prefix (dense: m; dispatch: entry.m; routed:
      // moe_align_block_size puts real ids first), so one count describes every thread's rows.
      int nv;
      if (p.rm.mode == ROWS_DENSE) {
        nv = min(NB, p.rm.m - blk * NB);
      } else if (p.rm.mode == ROWS_DISPATCH) {
        nv = min(NB, tbl->e[blk].m);
      } else {
        int c = 0;
        for (int i = lane; i < p.rm.block; i += 32)
          c += __ldcg(p.rm.sorted_ids + (size_t)blk * p.rm.block + i) < p.rm.num_valid ? 1 : 0;
        nv = min(NB, __reduce_add_sync(0xffffffffu, c));
      }
      constexpr int CH = HALF < 16 ? HALF : 16;
      constexpr int NCH = HALF / CH;
      constexpr int kPart = NACC * NB * kTileN;
      const int col = tile * kTileN + nl;
      const float bias = (!kSilu && p.w.bias) ? p.w.bias[col] : 0.0f;
      const bool whole = (r_begin == 0 && r_end == upt);         // this CTA covered the whole tile
      // Fused all-reduce (row-parallel linear): this rank's finished tiles go straight from the
      // epilogue into slot [epoch parity][this rank] of EVERY rank's receive buffer (peer memory over
      // NVLink, IPC-mapped); p2p_reduce_kernel on each rank then sums the world slots.
      const size_t push_off = push_world(p) ? ((size_t)(*(volatile const uint32_t *)p.push.epoch & 1u) * p.push.world +
                                              p.push.rank) * p.push.slot_elems : 0;
      // A tile cut by CTA range boundaries has contributors c_first .. c_last (consecutive CTAs).
      // c_first -- whose piece of the tile is the LAST segment of its range -- owns the tile: it
      // keeps its own accumulator in registers, waits until the others have published their fp32
      // partials and sums them in contributor order (deterministic).  The others (for them the
      // tile is the FIRST segment of their range) publish a partial and never stall in mid-stream.
      int c_first = cta, c_last = cta;
      if (!whole) {
        const u32 tile_u0 = (u32)(sg.blk * n_tiles + sg.tile) * (u32)upt;      // in scheduling units
        c_first = cta_of_unit(tile_u0, T, G);
        c_last = cta_of_unit(tile_u0 + upt - 1, T, G);
      }
      const bool owner = !whole && cta == c_first;
      // exchange protocol of this tile (identical in all of its contributors: nv depends on the block only)
      const bool ll = !whole && NB <= 64 && nv <= kLlRows && G <= p.ll_ctas;
      auto store_out = [&](int row_in_block, float v0, float v1) {
        int ar, cr;
        if (!rowmap_resolve(p.rm, tbl, NB, blk, row_in_block, ar, cr)) return;
        const float o = kSilu ? silu_mul(v0, v1) : v0 + bias;
        if (push_world(p) == 0) {
          p.y[(size_t)cr * p.ld_out + col] = __float2half_rn(o);
        } else {
          const __half hv = __float2half_rn(o);
          for (int r = 0; r < push_world(p); r++) p.push.data[r][push_off + (size_t)cr * p.ld_out + col] = hv;
        }
      };

      ptx::mbar_wait(bar(BAR_ACC), seg & 1);

      FB2_TL(5);
      seg++;
      ptx::tc_fence_after();
      if (ll) {
        // ---- flag-in-value exchange, four rows at a time.  Row groups alternate between the two warp
        // sets (any warp may read any accumulator COLUMN of its lane quarter), so eight rows still take
        // a single polling round.
#pragma unroll 1
        for (int jr = rset * kLlGroup; jr < nv; jr += 2 * kLlGroup) {   // warp-uniform loop bounds
          const int nr = nv - jr;                                // rows of this group: j < nr
          float v[NACC][kLlGroup];
#pragma unroll
          for (int a = 0; a < NACC; a++) {
            const bool touched = r_begin < (a + 1) * groups && r_end > a * groups;   // CTA-uniform
            if (touched) {
              uint32_t r[kLlGroup];
              ptx::tmem_ld_x4(lane_base + a * NB + jr, r);
              ptx::tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < kLlGroup; j++) v[a][j] = __uint_as_float(r[j]);
            } else {
#pragma unroll
              for (int j = 0; j < kLlGroup; j++) v[a][j] = 0.0f;
            }
          }
          auto slot_of = [&](int c, int a, int j) {
            return p.ll + (size_t)phys_cta(c) * kLlFloatsPerCta + (size_t)((a * kLlRows + jr + j) * kTileN + nl);
          };
          if (!owner) {
#pragma unroll
            for (int a = 0; a < NACC; a++)
#pragma unroll
              for (int j = 0; j < kLlGroup; j++)
                if (j < nr) {
                  const uint32_t u = __float_as_uint(v[a][j]);
                  ptx::st_relaxed_u32(slot_of(cta, a, j), u == kLlEmpty ? 0x7fffffffu : u);
                }
          } else {
            constexpr int UB = 8 / NACC;                         // contributors polled together
#pragma unroll 1
            for (int c0 = c_first + 1; c0 <= c_last; c0 += UB) {
              uint32_t t[UB][NACC][kLlGroup];
#pragma unroll
              for (int i = 0; i < UB; i++)
#pragma unroll
                for (int a = 0; a < NACC; a++)
#pragma unroll
                  for (int j = 0; j < kLlGroup; j++)
                    t[i][a][j] = (j < nr && c0 + i <= c_last) ? ptx::ld_relaxed_u32(slot_of(c0 + i, a, j)) : 0u;
#ifdef FB2_WAIT_TIMEOUT
              const long long t0 = clock64();
#endif
              for (;;) {
                // all slots are tested first and all missing ones re-requested together: a test
                // between two loads would serialise the round trips
                bool pending = false;
#pragma unroll
                for (int i = 0; i < UB; i++)
#pragma unroll
                  for (int a = 0; a < NACC; a++)
#pragma unroll
                    for (int j = 0; j < kLlGroup; j++) pending |= t[i][a][j] == kLlEmpty;
                if (!pending) break;
#pragma unroll
                for (int i = 0; i < UB; i++)
#pragma unroll
                  for (int a = 0; a < NACC; a++)
#pragma unroll
                    for (int j = 0; j < kLlGroup; j++)
                      if (t[i][a][j] == kLlEmpty)                // only ever true for slots that were requested
                        t[i][a][j] = ptx::ld_relaxed_u32(slot_of(c0 + i, a, j));
#ifdef FB2_WAIT_TIMEOUT
                if (clock64() - t0 > 4000000000LL) { printf("fb2: ll slot wait timed out (cta %d tile %d)\n", cta, tau); __trap(); }
#endif
              }
#pragma unroll
              for (int i = 0; i < UB; i++)
#pragma unroll
                for (int a = 0; a < NACC; a++)
#pragma unroll
                  for (int j = 0; j < kLlGroup; j++)
                    if (j < nr && c0 + i <= c_last) {
                      v[a][j] += __uint_as_float(t[i][a][j]);
                      ptx::st_relaxed_u32(slot_of(c0 + i, a, j), kLlEmpty);   // consumed: empty again for the next launch
                    }
            }
#pragma unroll
            for (int j = 0; j < kLlGroup; j++)
              if (j < nr) store_out(jr + j, v[0][j], v[NACC - 1][j]);
          }
        }
      } else {
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
          __syncwarp();                                            // reconverge before .sync.aligned tcgen05.ld
          ptx::named_bar_sync(1, kDequantWarps * 32);
        }
        float *part = p.partials + (size_t)phys_cta(cta) * kPart;
        // The accumulator is read in chunks of CH columns (activation rows) so that the same code
        // serves NB = 16 ... 256 without holding NB/2 floats per thread.
#pragma unroll 1
        for (int ch = 0; ch < NCH; ch++) {
          const int nvc = nv - (rset * HALF + ch * CH);               // real rows of this chunk (warp-uniform)
          if (nvc <= 0) break;                                     // a prefix: later chunks are empty too
          float v[NACC][CH];
#pragma unroll
          for (int a = 0; a < NACC; a++) {
            const bool touched = r_begin < (a + 1) * groups && r_end > a * groups;   // CTA-uniform
            if (touched) {
              load_acc_chunk<CH>(lane_base + a * NB + rset * HALF + ch * CH, v[a]);
            } else {
#pragma unroll
              for (int j = 0; j < CH; j++) v[a][j] = 0.0f;
            }
          }
          const size_t row_off = (size_t)(rset * HALF + ch * CH) * kTileN + nl;
          if (owner) {
            // UC contributors x CH rows x NACC accumulators of independent L2 loads in flight; the sums
            // are taken in contributor order (deterministic)
            constexpr int UC = NACC * CH <= 8 ? 4 : (NACC * CH <= 16 ? 2 : 1);
#pragma unroll 1
            for (int c0 = c_first + 1; c0 <= c_last; c0 += UC) {
              float t[UC][NACC][CH];
#pragma unroll
              for (int i = 0; i < UC; i++) {
                const float *src = p.partials + (size_t)phys_cta(c0 + i) * kPart + row_off;
#pragma unroll
                for (int a = 0; a < NACC; a++)
#pragma unroll
                  for (int j = 0; j < CH; j++)
                    t[i][a][j] = (j < nvc && c0 + i <= c_last) ? __ldcg(src + (size_t)(a * NB + j) * kTileN) : 0.0f;
              }
#pragma unroll
              for (int i = 0; i < UC; i++)
#pragma unroll
                for (int a = 0; a < NACC; a++)
#pragma unroll
                  for (int j = 0; j < CH; j++) v[a][j] += t[i][a][j];
            }
          }
          if (whole || owner) {
#pragma unroll
            for (int j = 0; j < CH; j++)
              if (j < nvc) store_out(rset * HALF + ch * CH + j, v[0][j], v[NACC - 1][j]);
          } else {
#pragma unroll
            for (int a = 0; a < NACC; a++)
#pragma unroll
              for (int j = 0; j < CH; j++)
                if (j < nvc) part[row_off + (size_t)(a * NB + j) * kTileN] = v[a][j];
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar(BAR_ACC_EMPTY));      // accumulators may be overwritten
      FB2_TL(6);
      if (!whole && !owner && !ll) {
        ptx::named_bar_sync(1, kDequantWarps * 32);              // every partial store of the CTA is issued
        if (threadIdx.x == 0) ptx::red_release_add(p.counters + tau, 1);
        __syncwarp();
      }
    }
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
  if constexpr (kPair) ptx::cluster_sync_all();                  // the peer no longer reaches this CTA's memory
  if (warp == kWarpMma) ptx::tmem_dealloc<kTmemCols>(tmem_base);
  FB2_TL(7);
}

template <int NB, bool kSilu, bool kPush = false, bool kPair = false>
__global__ void __launch_bounds__(StageCfg<NB>::kBoundThreads, StageCfg<NB>::kCtasPerSm)
tc_gemm_kernel(const TcParams p, const __grid_constant__ CUtensorMap xmap) {
  ptx::launch_dependents();
  tc_body<NB, kSilu, kPush, kPair>(p, nullptr, &xmap);
}
template <int NB>
__global__ void __launch_bounds__(StageCfg<NB>::kBoundThreads, StageCfg<NB>::kCtasPerSm)
tc_gemm_dispatch_kernel(const TcParams p, const __grid_constant__ CUtensorMap xmap,
                        const __grid_constant__ DispatchTable tbl) {
  ptx::launch_dependents();
  tc_body<NB, false>(p, &tbl, &xmap);
}

// One-shot all-reduce, receive side: wait until every rank's CTAs have signalled this epoch's slot,
// sum the world slots in rank order (fp32, deterministic), write y.  The last CTA re-arms the flags
// and advances the epoch (slots and flags are double-buffered by epoch parity, so a peer that is one
// call ahead never touches what this call still reads).
__global__ void __launch_bounds__(256) p2p_reduce_kernel(const __half *__restrict__ data, uint32_t *flags,
                                                         uint32_t *epoch, uint32_t *done, __half *__restrict__ y,
                                                         size_t elems, size_t slot_elems, int world,
                                                         const PushExpect expect) {
  ptx::launch_dependents();
  ptx::grid_dependency_wait();
  const uint32_t e = *(volatile uint32_t *)epoch, par = e & 1u;
  if ((int)threadIdx.x < world) {
    const uint32_t need = expect.ctas[threadIdx.x];
    const long long t0 = clock64();
    while (ptx::ld_acquire_sys(flags + par * world + threadIdx.x) < need) {
      if (clock64() - t0 > 20000000000LL) { printf("fb2: peer %d never completed its push\n", (int)threadIdx.x); __trap(); }
    }
  }
  __syncthreads();
  const size_t n8 = elems / 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) acc[j] = 0.f;
    for (int r = 0; r < world; r++) {
      const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(data + ((size_t)par * world + r) * slot_elems) + i);
      const __half2 *hp = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
      for (int j = 0; j < 4; j++) { const float2 f = __half22float2(hp[j]); acc[2 * j] += f.x; acc[2 * j + 1] += f.y; }
    }
    uint4 o;
    __half2 *op = reinterpret_cast<__half2 *>(&o);
#pragma unroll
    for (int j = 0; j < 4; j++) op[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
    reinterpret_cast<uint4 *>(y)[i] = o;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t old = atomicAdd(done, 1u);
    if (old == gridDim.x - 1) {
      for (int r = 0; r < world; r++) flags[par * world + r] = 0u;
      *done = 0u;
      __threadfence();
      *(volatile uint32_t *)epoch = e + 1u;
    }
  }
}

#include "tc_decode.cuh"

// ------------------------------------------------------------------------ host planning
struct TcPlan {
  int nb, row_blocks, n_tiles, nacc, grid, whole_tiles;
  long long units;
  size_t partial_floats, counters;
  uint32_t smem;
  int paired;             // prefill row blocks: clusters of two CTAs on adjacent tiles (tc_body kPair)
};

int pick_nb(int rows_per_block, bool allow_wide) {
  if (rows_per_block <= 16) return 16;
  if (rows_per_block <= 32) return 32;
  if (rows_per_block <= 64 || !allow_wide) return 64;
  static const int force = [] { const char *e = getenv("FB2_DEBUG_WIDE_NB"); return e ? atoi(e) : 0; }();
  if (force == 128 || force == 256) return force;
  return rows_per_block <= 128 ? 128 : 256;
}

void plan_rows(const RowMap &rows, int epilogue, int &nb, int &row_blocks) {
  if (rows.mode == ROWS_DENSE) {
    nb = pick_nb(rows.m, !epilogue);
    row_blocks = (rows.m + nb - 1) / nb;
  } else if (rows.mode == ROWS_DISPATCH) {
    nb = rows.block;           // host pre-split the entries into chunks of <= rows.block rows
    row_blocks = rows.num_dispatch;
  } else {
    nb = pick_nb(rows.block, false);
    row_blocks = rows.max_blocks;
  }
}

template <int NB> constexpr uint32_t smem_of() { return (uint32_t)StageCfg<NB>::kSmemBytes; }

TcPlan make_plan(const WeightRef &w, const RowMap &rows, int epilogue, int sm_count) {
  TcPlan pl{};
  plan_rows(rows, epilogue, pl.nb, pl.row_blocks);
  pl.nacc = epilogue ? 2 : 1;
  pl.n_tiles = (epilogue ? w.n / 2 : w.n) / kTileN;
  const long long units = (long long)pl.row_blocks * pl.n_tiles * pl.nacc * (w.k / kGroup);
  pl.units = units;
  static const int ctas_per_sm_dbg = [] { const char *e = getenv("FB2_DEBUG_CTAS_PER_SM"); return e ? atoi(e) : 0; }();
  const int per_sm = ctas_per_sm_dbg ? ctas_per_sm_dbg : (pl.nb >= 128 ? 1 : 2);
  const int slots = per_sm * (sm_count > 0 ? sm_count : 148);
  const int upt = pl.nacc * (w.k / kGroup);
  // A prefill batch over a narrow layer (Llama o_proj / down_proj at M = 512: 32 tiles): 128-row blocks give one whole
  // tile per CTA -- no split tiles, whose partials are 128 KB each at 256 rows (o_proj: 26.2 -> 19.6 us).
  if (rows.mode == ROWS_DENSE && pl.nb == 256 && !epilogue) {
    static const bool forced = getenv("FB2_DEBUG_WIDE_NB") != nullptr;
    const long long nt128 = (long long)((rows.m + 127) / 128) * pl.n_tiles;
    if (!forced && whole_tile_mode_wide(nt128, slots)) {
      pl.nb = 128;
      pl.row_blocks = (rows.m + 127) / 128;
    }
  }
  const long long units2 = (long long)pl.row_blocks * pl.n_tiles * pl.nacc * (w.k / kGroup);
  pl.units = units2;
  static const bool no_align = [] { const char *e = getenv("FB2_DEBUG_NO_TILE_ALIGN"); return e && e[0] == '1'; }();
  static const bool no_pair = [] { const char *e = getenv("FB2_DEBUG_NO_PAIR"); return e && e[0] == '1'; }();
  pl.paired = !no_pair && pl.nb >= 128 && rows.mode == ROWS_DENSE && !epilogue && pl.n_tiles % 2 == 0 && slots >= 2;
  GridPlan gp;
  if (pl.paired) {
    // schedule PAIRS: half the tiles (tile pairs), half the slots; the grid is twice the number of scheduling CTAs
    gp = plan_grid(pl.units / 2, upt, slots / 2, false, !no_align, true);
    gp.grid *= 2;
  } else {
    gp = plan_grid(pl.units, upt, slots, rows.mode == ROWS_ROUTED, !no_align, pl.nb >= 128);   // sched.cuh
  }
  pl.whole_tiles = gp.whole_tiles;
  pl.grid = gp.grid;
  pl.partial_floats = (size_t)pl.grid * pl.nacc * pl.nb * kTileN;
  pl.counters = (size_t)pl.row_blocks * pl.n_tiles;
  pl.smem = pl.nb == 16 ? smem_of<16>() : pl.nb == 32 ? smem_of<32>() : pl.nb == 64 ? smem_of<64>()
            : pl.nb == 128 ? smem_of<128>() : smem_of<256>();
  return pl;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess ||
        qr != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// x viewed as a [rows, K] fp16 tensor; box = 64 k (128 B, SWIZZLE_128B) x nb rows; OOB rows -> 0.
int make_x_tensor_map(CUtensorMap *map, const __half *x, size_t rows, size_t k, int nb) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return FB2_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)k * sizeof(__half)};
  const cuuint32_t box[2] = {64u, (cuuint32_t)nb};
  const cuuint32_t estride[2] = {1u, 1u};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half *>(x), gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) for x [%zu, %zu], box rows %d", (int)r, rows, k, nb);
    return FB2_ERR_CUDA;
  }
  return FB2_OK;
}

void launch_cfg(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr, const TcPlan &pl, cudaStream_t stream) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3((unsigned)pl.grid);
  cfg.blockDim = dim3(pl.nb >= 128 ? StageCfg<256>::kCtaThreads : kThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

template <int NB>
int launch_paired(const TcPlan &pl, cudaStream_t stream, const TcParams &p, const CUtensorMap &xmap) {
  static const cudaError_t attr_rc = cudaFuncSetAttribute(tc_gemm_kernel<NB, false, false, true>,
                                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of<NB>());
  FB2_CUDA_TRY(attr_rc);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  launch_cfg(cfg, attr, pl, stream);
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.numAttrs = 2;
  FB2_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<NB, false, false, true>, p, xmap));
  return FB2_OK;
}
template <int NB, bool kSilu, bool kPush = false>
int launch_plain(const TcPlan &pl, cudaStream_t stream, const TcParams &p, const CUtensorMap &xmap) {
  FB2_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_kernel<NB, kSilu, kPush>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  launch_cfg(cfg, attr, pl, stream);
  FB2_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<NB, kSilu, kPush>, p, xmap));
  return FB2_OK;
}
// decode-batch kernel (tc_decode.cuh): dense rows through the tensor map, act-order rows gathered inside the kernel,
// or routed blocks of an expert stack (with or without the SwiGLU epilogue)
template <int R, int kMode, bool kSilu, bool kPush>
int launch_dk(const TcPlan &pl, cudaStream_t stream, const TcParams &p, const CUtensorMap &xmap) {
  static const cudaError_t attr_rc = cudaFuncSetAttribute(dk::dk_gemm_kernel<R, kMode, kSilu, kPush>,
                                                          cudaFuncAttributeMaxDynamicSharedMemorySize, dk::kSmemBudget);
  FB2_CUDA_TRY(attr_rc);
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  launch_cfg(cfg, attr, pl, stream);
  cfg.blockDim = dim3(dk::kDkThreads);
  cfg.dynamicSmemBytes = dk::kSmemBudget;
  FB2_CUDA_TRY(cudaLaunchKernelEx(&cfg, dk::dk_gemm_kernel<R, kMode, kSilu, kPush>, p, xmap));
  return FB2_OK;
}
template <int kMode>
int launch_dk_rows_silu(int m, const TcPlan &pl, cudaStream_t stream, const TcParams &p, const CUtensorMap &xmap) {
  return m <= 1 ? launch_dk<1, kMode, true, false>(pl, stream, p, xmap)
         : m <= 2 ? launch_dk<2, kMode, true, false>(pl, stream, p, xmap)
         : m <= 4 ? launch_dk<4, kMode, true, false>(pl, stream, p, xmap) : launch_dk<8, kMode, true, false>(pl, stream, p, xmap);
}
template <int kMode, bool kPush>
int launch_dk32(const TcPlan &pl, cudaStream_t stream, const TcParams &p, const CUtensorMap &xmap) {
  if constexpr (kMode == dk::DK_DENSE) {
    return launch_dk<32, kMode, false, kPush>(pl, stream, p, xmap);
  } else {
    set_error("internal: the 32-row decode kernel takes its rows through the tensor map only");
    return FB2_ERR_ARG;
  }
}
template <int kMode, bool kPush>
int launch_dk_rows(int m, const TcPlan &pl, cudaStream_t stream, const TcParams &p, const CUtensorMap &xmap) {
  return m <= 1 ? launch_dk<1, kMode, false, kPush>(pl, stream, p, xmap)
         : m <= 2 ? launch_dk<2, kMode, false, kPush>(pl, stream, p, xmap)
         : m <= 4 ? launch_dk<4, kMode, false, kPush>(pl, stream, p, xmap)
         : m <= 8 ? launch_dk<8, kMode, false, kPush>(pl, stream, p, xmap)
         : m <= 16 ? launch_dk<16, kMode, false, kPush>(pl, stream, p, xmap)
         : launch_dk32<kMode, kPush>(pl, stream, p, xmap);
  // (17..32 rows: R = 32, UMMA N = 32.  With 80 registers per thread its 32 running sums spilled and it lost to
  // tc_body<32>; with the dequant warpgroups at 96 registers it wins: 9.97 / 9.41 / 20.5 / 15.9 us on the four Llama
  // shapes at M = 32 against 11.8 / 11.2 / 22.8 / 17.7 us.)
}

template <int NB>
int launch_dispatch(const TcPlan &pl, cudaStream_t stream, const TcParams &p, const CUtensorMap &xmap,
                    const DispatchTable &tbl) {
  FB2_CUDA_TRY(cudaFuncSetAttribute(tc_gemm_dispatch_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  launch_cfg(cfg, attr, pl, stream);
  FB2_CUDA_TRY(cudaLaunchKernelEx(&cfg, tc_gemm_dispatch_kernel<NB>, p, xmap, tbl));
  return FB2_OK;
}

}  // namespace

// The decode kernel (tc_decode.cuh) serves dense launches of at most 32 rows (8 with the SwiGLU epilogue: two
// accumulators' worth of running sums per thread).
static bool dk_eligible(const WeightRef &w, const RowMap &rows, const DispatchTable *tbl, int epilogue) {
  static const bool no_dk = [] { const char *e = getenv("FB2_DEBUG_NO_DK"); return e && e[0] == '1'; }();
  static const int max_m = [] { const char *e = getenv("FB2_DEBUG_DK_MAXM"); return e ? atoi(e) : 32; }();
  return !no_dk && rows.mode == ROWS_DENSE && rows.m >= 1 && rows.m <= max_m && rows.m <= (epilogue ? 8 : 32) && !tbl;
}
// ... and gathers act-order activations itself when the m rows of x fit next to a weight ring of kMinWStages
bool dk_gather_ok(const WeightRef &w, const RowMap &rows, const DispatchTable *tbl, int epilogue) {
  static const bool no_gather = [] { const char *e = getenv("FB2_DEBUG_NO_DK_GATHER"); return e && e[0] == '1'; }();
  return !no_gather && w.perm && dk_eligible(w, rows, tbl, epilogue) && w.k % 8 == 0 &&
         rows.m <= 16 && dk::w_stages_for(epilogue != 0, rows.m * w.k * 2) >= dk::kMinWStages;
}
// ... and the routed blocks of an expert stack when they hold at most 8 rows (16 without the SwiGLU epilogue behind
// FB2_DEBUG_DK_ROUTED_MAXBLOCK=16)
bool dk_routed_ok(const WeightRef &w, const RowMap &rows, int epilogue) {
  static const bool no_dk = [] { const char *e = getenv("FB2_DEBUG_NO_DK"); return e && e[0] == '1'; }();
  static const int max_block = [] { const char *e = getenv("FB2_DEBUG_DK_ROUTED_MAXBLOCK"); return e ? atoi(e) : 8; }();
  if (no_dk || rows.mode != ROWS_ROUTED) return false;   // (act-order: run_tc permutes x before the launch)
  return rows.block <= (epilogue ? 8 : (max_block < 16 ? max_block : 16));
}

bool tc_supported(const WeightRef &w) {
  return w.tiles != nullptr && w.group_size == kGroup && w.k % kGroup == 0 && w.n % kTileN == 0 &&
         w.col0 % kTileN == 0 && w.ld % kTileN == 0;
}

const char *tc_kernel_name(const WeightRef &w, size_t m) {
  if (!tc_supported(w)) return "generic";
  return m > 64 ? "tc-streamk-prefill" : "tc-streamk";
}

void tc_requirements(const WeightRef &w, const RowMap &rows, int epilogue, int sm_count,
                     size_t *partial_bytes, size_t *counters) {
  *partial_bytes = 0;
  *counters = 0;
  int nb, row_blocks;
  plan_rows(rows, epilogue, nb, row_blocks);
  if (row_blocks <= 0) return;
  TcPlan pl = make_plan(w, rows, epilogue, sm_count);
  *partial_bytes = pl.partial_floats * sizeof(float);
  *counters = pl.counters;
}

int launch_p2p_reduce(const P2pLocal &loc, const PushExpect &expect, __half *y, size_t elems, cudaStream_t stream) {
  if (elems % 8 || ((uintptr_t)y % 16)) {
    set_error("fused all-reduce: output must be 16-byte aligned with a multiple of 8 elements");
    return FB2_ERR_ARG;
  }
  cudaLaunchConfig_t cfg{};
  const size_t n8 = elems / 8;
  cfg.gridDim = dim3((unsigned)std::min<size_t>((n8 + 255) / 256, 148));
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FB2_CUDA_TRY(cudaLaunchKernelEx(&cfg, p2p_reduce_kernel, (const __half *)loc.data, loc.flags, loc.epoch, loc.done, y,
                                  elems, loc.slot_elems, loc.world, expect));
  count_launch();
  return FB2_OK;
}

int tc_grid_size(const WeightRef &w, const RowMap &rows, int epilogue, int sm_count) {
  int nb, row_blocks;
  plan_rows(rows, epilogue, nb, row_blocks);
  if (row_blocks <= 0) return 0;
  return make_plan(w, rows, epilogue, sm_count).grid;
}

int launch_tc_gemm(const WeightRef &w, const RowMap &rows, const DispatchTable *tbl, const __half *x,
                   size_t x_rows, __half *y, int epilogue, const TcScratch &scratch, int sm_count,
                   cudaStream_t stream, const PushTargets *push, int dk_flags, const int32_t *out_map) {
  if (!tc_supported(w)) {
    set_error("tensor-core path needs the tiled weight format (K,N multiples of 128, group 128)");
    return FB2_ERR_UNSUPPORTED;
  }
  int nb, row_blocks;
  plan_rows(rows, epilogue, nb, row_blocks);
  if (row_blocks <= 0) return FB2_OK;
  if (epilogue && (w.n / 2) % kTileN) {
    set_error("SiLU epilogue needs n_per_expert/2 to be a multiple of 128");
    return FB2_ERR_SHAPE;
  }
  if ((rows.mode == ROWS_DISPATCH) != (tbl != nullptr) || (tbl && epilogue)) {
    set_error("internal: dispatch table / row mode mismatch");
    return FB2_ERR_ARG;
  }
  TcPlan pl = make_plan(w, rows, epilogue, sm_count);
  nb = pl.nb;                                  // (make_plan may choose narrower prefill row blocks than plan_rows)
  // the kernels' unit bookkeeping is 32-bit: units x CTAs stays below 2^31
  // (FB2_DEBUG_MAX_UNIT_CTAS lowers the limit so that a test can reach the chunked path with a small batch)
  static const long long unit_cta_limit = [] {
    const char *e = getenv("FB2_DEBUG_MAX_UNIT_CTAS");
    const long long v = e ? atoll(e) : 0;
    return v > 0 && v < (1LL << 31) ? v : (1LL << 31);
  }();
  if (pl.units * pl.grid >= unit_cta_limit) {
    if (rows.mode != ROWS_DENSE || push) {
      set_error("launch too large for one grid (%lld units x %d CTAs)", pl.units, pl.grid);
      return FB2_ERR_UNSUPPORTED;
    }
    // dense batch: cut the rows into several launches of whole row blocks
    const long long per_block = pl.units / pl.row_blocks;
    long long blocks = (unit_cta_limit - 1) / ((long long)pl.grid * per_block);
    if (blocks < 1) { set_error("weight matrix too large for one grid"); return FB2_ERR_UNSUPPORTED; }
    const size_t chunk = (size_t)blocks * pl.nb;
    const size_t ld_out = epilogue ? w.n / 2 : w.n;
    for (size_t r0 = 0; r0 < (size_t)rows.m; r0 += chunk) {
      RowMap sub = rows;
      sub.m = (int)std::min(chunk, (size_t)rows.m - r0);
      int rc = launch_tc_gemm(w, sub, nullptr, x + r0 * w.k, (size_t)sub.m, y + r0 * ld_out, epilogue, scratch, sm_count,
                              stream, nullptr, 0, nullptr);
      if (rc) return rc;
    }
    return FB2_OK;
  }
  if (pl.partial_floats * 4 > scratch.partial_bytes || pl.counters > scratch.num_counters) {
    set_error("split-K scratch too small (%zu > %zu bytes, %zu > %zu counters): call "
              "fb2_pregrow_scratch outside graph capture",
              pl.partial_floats * 4, scratch.partial_bytes, pl.counters, scratch.num_counters);
    return FB2_ERR_ARG;
  }
  TcParams p{};
  p.w = w;
  p.rm = rows;
  p.x = x;
  p.y = y;
  p.partials = scratch.partials;
  p.counters = scratch.counters;
  p.ll = scratch.ll;
  p.ll_ctas = scratch.ll ? scratch.ll_ctas : 0;
  p.n_tiles = pl.n_tiles;
  p.groups = w.k / kGroup;
  p.tiles_per_expert = w.ld / kTileN;
  p.ld_out = epilogue ? w.n / 2 : w.n;
  p.whole_tiles = pl.whole_tiles;
  if (push) p.push = *push; else p.push = PushTargets{};
  // Rows that are contiguous in x (dense batches, dispatch entries) come through a 2-D TMA tensor
  // map: one instruction per [NB x 64] box instead of NB*16 cp.async.  Routed rows are gathered.
  CUtensorMap xmap;
  memset(&xmap, 0, sizeof(xmap));
  p.x_tma = 0;
  if (w.perm) {
    // act-order gather inside the decode kernel (run_tc only leaves perm set when dk_gather_ok said so)
    if (!dk_gather_ok(w, rows, tbl, epilogue) || ((uintptr_t)x % 16) != 0) {
      set_error("internal: act-order activations must be gathered before this launch");
      return FB2_ERR_ARG;
    }
  } else if (rows.mode != ROWS_ROUTED && x_rows > 0 && ((uintptr_t)x % 16) == 0) {
    int rc = make_x_tensor_map(&xmap, x, x_rows, (size_t)w.k, pl.paired ? nb / 2 : nb);   // paired CTAs load half a block each
    if (rc) return rc;
    p.x_tma = 1;
  } else if (nb > 64) {
    set_error("internal: wide row blocks need the tensor-map path");
    return FB2_ERR_ARG;
  }

#ifdef FB2_TRACE
  {
    static int seq = 0;
    static bool init = false;
    if (!init) {
      init = true;
      if (const char *tb = getenv("FB2_DEBUG_BUF")) {
        long long *hp = (long long *)strtoull(tb, nullptr, 10);
        cudaMemcpyToSymbol(g_trace, &hp, sizeof(hp));
      }
    }
    p.x_tma |= (seq++ & 7) << 16;
  }
#endif
  int rc;
  const bool use_dk_dense = dk_eligible(w, rows, tbl, epilogue) && ((p.x_tma & 1) || w.perm) && !(epilogue && push && push->world);
  const bool use_dk_routed = dk_routed_ok(w, rows, epilogue) && !(push && push->world);
  if (out_map && !use_dk_routed) {
    set_error("internal: an output column map needs the routed decode kernel");
    return FB2_ERR_ARG;
  }
  if (use_dk_dense || use_dk_routed) {
    p.l2_prefetch = 0;
    p.dk_flags = use_dk_routed ? dk_flags : 0;
    p.out_map = use_dk_routed ? out_map : nullptr;
    p.fd_grid = make_fastdiv((u32)pl.grid);
    p.fd_upt = make_fastdiv((u32)(pl.nacc * p.groups));
    p.fd_tiles = make_fastdiv((u32)pl.n_tiles);
    p.fd_units = make_fastdiv((u32)pl.units);
    const bool pushing = push && push->world;
    if (use_dk_routed) {
      p.w_stages = dk::w_stages_for(epilogue != 0, (dk::kXSlotsRouted - dk::kXSlots) * dk::Lay<false, 16>::kXBytes);
      if (rows.block <= 8) {
        rc = epilogue ? launch_dk<8, dk::DK_ROUTED, true, false>(pl, stream, p, xmap)
                      : launch_dk<8, dk::DK_ROUTED, false, false>(pl, stream, p, xmap);
      } else if (epilogue) {
        set_error("internal: SwiGLU decode kernel is instantiated for blocks of <= 8 rows");
        return FB2_ERR_ARG;
      } else {
        rc = launch_dk<16, dk::DK_ROUTED, false, false>(pl, stream, p, xmap);
      }
    } else if (w.perm) {
      p.w_stages = dk::w_stages_for(epilogue != 0, rows.m * w.k * 2);
      rc = epilogue ? launch_dk_rows_silu<dk::DK_GATHER>(rows.m, pl, stream, p, xmap)
           : pushing ? launch_dk_rows<dk::DK_GATHER, true>(rows.m, pl, stream, p, xmap)
                     : launch_dk_rows<dk::DK_GATHER, false>(rows.m, pl, stream, p, xmap);
    } else {
      p.w_stages = dk::w_stages_for(epilogue != 0, 0, nb);
      rc = epilogue ? launch_dk_rows_silu<dk::DK_DENSE>(rows.m, pl, stream, p, xmap)
           : pushing ? launch_dk_rows<dk::DK_DENSE, true>(rows.m, pl, stream, p, xmap)
                     : launch_dk_rows<dk::DK_DENSE, false>(rows.m, pl, stream, p, xmap);
    }
  } else if (push && push->world) {
    if (tbl || epilogue || nb > 64) {
      set_error("fused all-reduce: decode-sized dense launches only (<= 64 rows per block)");
      return FB2_ERR_UNSUPPORTED;
    }
    rc = nb == 16 ? launch_plain<16, false, true>(pl, stream, p, xmap)
         : nb == 32 ? launch_plain<32, false, true>(pl, stream, p, xmap) : launch_plain<64, false, true>(pl, stream, p, xmap);
  } else if (tbl) {
    rc = nb == 16 ? launch_dispatch<16>(pl, stream, p, xmap, *tbl)
         : nb == 32 ? launch_dispatch<32>(pl, stream, p, xmap, *tbl) : launch_dispatch<64>(pl, stream, p, xmap, *tbl);
  } else if (epilogue) {
    rc = nb == 16 ? launch_plain<16, true>(pl, stream, p, xmap)
         : nb == 32 ? launch_plain<32, true>(pl, stream, p, xmap) : launch_plain<64, true>(pl, stream, p, xmap);
  } else {
    rc = nb == 16 ? launch_plain<16, false>(pl, stream, p, xmap)
         : nb == 32 ? launch_plain<32, false>(pl, stream, p, xmap)
         : nb == 64 ? launch_plain<64, false>(pl, stream, p, xmap)
         : nb == 128 ? (pl.paired ? launch_paired<128>(pl, stream, p, xmap) : launch_plain<128, false>(pl, stream, p, xmap))
                     : (pl.paired ? launch_paired<256>(pl, stream, p, xmap) : launch_plain<256, false>(pl, stream, p, xmap));
  }
  if (rc) return rc;
  count_launch();
  return FB2_OK;
}

}  // namespace fb2
