// sched.cuh -- the stream-K bookkeeping of tc_gemm.cu: unit ranges per CTA, the owner / contributors of a split tile and
// the segment iterator.  Pure integer functions, __host__ __device__ so that tests/sched_check.cpp (g++, no GPU) can
// check their invariants over random shapes: ranges partition the units, cta_of_unit inverts range_start, segments
// never cross a tile, a tile's contributors are consecutive CTAs of which the first holds the tile as the LAST segment
// of its range and every other one as the FIRST.
#pragma once
#if defined(__CUDACC__)
#define FB2_HD __host__ __device__ __forceinline__
#else
#define FB2_HD inline
#endif

namespace fb2 {
namespace sched {

// ---- stream-K bookkeeping (identical in every role) -------------------------------------
// All of it is 32-bit unsigned arithmetic: the host keeps (units x CTAs) below 2^31 (launch_tc_gemm cuts
// larger dense batches into several launches).  The emulated 64-bit divisions this replaces sat on the
// critical path twice: before the first weight request of every CTA and in front of every fix-up.
typedef unsigned int u32;
FB2_HD u32 range_start(u32 c, u32 T, u32 G) { return c * T / G; }
// Division by a launch constant without the ~40-instruction emulated divide: the host passes floor(2^32 / d); the
// estimate hi32(n * magic) is the quotient or one below it for every 32-bit n.  (The divides sat on the path from CTA
// entry to the first weight request of the decode kernel: four of them were ~0.4 us of a single warp's time.)
struct FastDiv {
  u32 d, magic;
};
FB2_HD FastDiv make_fastdiv(u32 d) {
  FastDiv f;
  f.d = d;
  f.magic = d > 1 ? (u32)(0x100000000ULL / d) : 0xFFFFFFFFu;
  return f;
}
FB2_HD u32 fdiv(u32 n, FastDiv f) {
#if defined(__CUDA_ARCH__)
  u32 q = __umulhi(n, f.magic);
#else
  u32 q = (u32)(((unsigned long long)n * f.magic) >> 32);
#endif
  if (n - q * f.d >= f.d) q++;
  return q;
}
// Tiles of only a few units (K <= 1024: the MoE down projection has 6) are never split: cutting one
// costs a partial-tile fix-up (~3 us) that is longer than streaming the whole tile.  CTA c then owns
// the whole tiles [c * NT / G, (c + 1) * NT / G) -- when that balances: at most one tile per CTA, or so
// many that a difference of one tile is small (measured on the Qwen3 MoE down projection).
constexpr int kWholeTileMaxUnits = 8;
FB2_HD bool whole_tile_mode(int upt, long long n_tiles_total, int slots) {
  return upt <= kWholeTileMaxUnits && (n_tiles_total <= slots || n_tiles_total >= 4LL * slots);
}
// Prefill row blocks (128 / 256 rows): a partial tile is 64 / 128 KB of fp32 per contributor, so a launch whose tiles
// nearly fill the grid one-to-one keeps them whole whatever their length.
FB2_HD bool whole_tile_mode_wide(long long n_tiles_total, int slots) {
  return n_tiles_total <= slots && n_tiles_total * 5 >= slots * 4LL;
}
FB2_HD u32 tile_range_start(u32 c, u32 NT, u32 G, u32 upt) { return (c * NT / G) * upt; }
FB2_HD int cta_of_unit(u32 u, u32 T, u32 G) {
  u32 c = u * G / T;
  while (c + 1 < G && range_start(c + 1, T, G) <= u) c++;
  while (c > 0 && range_start(c, T, G) > u) c--;
  return (int)c;
}

FB2_HD int cta_of_unit_fd(u32 u, FastDiv T, FastDiv G) {
  u32 c = fdiv(u * G.d, T);
  while (c + 1 < G.d && fdiv((c + 1) * T.d, G) <= u) c++;
  while (c > 0 && fdiv(c * T.d, G) > u) c--;
  return (int)c;
}

// One CTA's unit range, cut into SEGMENTS: maximal runs of units that belong to one output
// tile (row block, tile).  r0/r1 index the tile's unit list [gate groups..., up groups...].
struct Seg {
  int blk, tile, r0, r1;
  int at_start;              // the segment starts at the CTA's range start (partial slot 0)
};
struct SegIter {
  int blk, tile, r, left, first;
  FB2_HD void init(u32 u0, u32 u1, int upt, int n_tiles) {
    const u32 tau = u0 / (u32)upt;
    r = (int)(u0 - tau * (u32)upt);
    blk = (int)(tau / (u32)n_tiles);
    tile = (int)(tau - (u32)blk * (u32)n_tiles);
    left = (int)(u1 - u0);
    first = 1;
  }
  FB2_HD void init_fd(u32 u0, u32 u1, FastDiv upt, FastDiv n_tiles) {
    const u32 tau = fdiv(u0, upt);
    r = (int)(u0 - tau * upt.d);
    blk = (int)fdiv(tau, n_tiles);
    tile = (int)(tau - (u32)blk * n_tiles.d);
    left = (int)(u1 - u0);
    first = 1;
  }
  FB2_HD bool next(Seg &s, int upt, int n_tiles) {
    if (left <= 0) return false;
    const int n = left < upt - r ? left : upt - r;
    s.blk = blk; s.tile = tile; s.r0 = r; s.r1 = r + n; s.at_start = first;
    first = 0;
    left -= n;
    r += n;
    if (r == upt) {
      r = 0;
      if (++tile == n_tiles) { tile = 0; blk++; }
    }
    return true;
  }
};

// Grid of one launch (host side, make_plan): one persistent wave of at most `slots` CTAs.
struct GridPlan {
  int grid, whole_tiles;
};
FB2_HD GridPlan plan_grid(long long units, int upt, int slots, bool routed, bool tile_align, bool wide = false) {
  GridPlan gp;
  gp.whole_tiles = !routed && (whole_tile_mode(upt, units / upt, slots) || (wide && whole_tile_mode_wide(units / upt, slots)));
  long long g = gp.whole_tiles ? units / upt : units / 2;         // whole tiles | >= 2 units per CTA
  if (g < 1) g = 1;
  gp.grid = (int)(g < slots ? g : slots);
  // Few tiles (decode qkv / o / down: 32-48 of them): give every tile the same whole number s of
  // CTAs.  With G = tiles * s the ranges c*T/G fall on tile boundaries, so no CTA straddles two tiles
  // (a straddling CTA pays a mid-stream epilogue AND owns the next tile's fix-up: it was the
  // straggler of these launches in the in-kernel timeline).
  const long long nt = units / upt;
  if (tile_align && !gp.whole_tiles && !routed && nt > 0 && nt * 2 <= slots) {
    long long s = slots / nt;
    if (s > upt / 2) s = upt / 2;                                   // >= 2 units per CTA
    if (s >= 1) {
      // the slowest CTA of a tile streams ceil(upt / s) units: the smallest s that reaches the minimum has the
      // fewest contributors per fix-up (o_proj: 8 CTAs x 4 units instead of 9 x 3..4) and leaves slots free for
      // the next kernel's early CTAs
      const long long best = (upt + s - 1) / s;
      while (s > 1 && (upt + s - 2) / (s - 1) == best) s--;
      gp.grid = (int)(nt * s);
    }
  }
  return gp;
}

}  // namespace sched
}  // namespace fb2
