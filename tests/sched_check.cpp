// Host check of the stream-K bookkeeping in ferrum-infer-rs_b200/csrc/sched.cuh (compiled with g++ by
// tests/test_sched_cpu.py; no GPU).  Random shapes plus the shapes of BASELINE.json; for each launch plan:
//   1. the CTA ranges partition [0, T) in order, and every participating CTA's range is non-empty;
//   2. cta_of_unit() returns the CTA whose range holds the unit;
//   3. a CTA's segments tile its range, never cross an output tile, and name the right (block, tile, r);
//   4. the contributors of a split tile are the consecutive CTAs c_first .. c_last; the tile is the LAST segment of
//      c_first's range and the FIRST segment (at_start) of every other contributor's -- the ownership rule of the fix-up;
//      4b. (distributed owners, tc_decode.cuh) the tile is also the LAST segment of every contributor but c_last, so an
//      owner never waits in the middle of its range; and a CTA holds at most one piece as "last contributor" and at most one
//      other split piece -- one per slot area;
//   5. in whole-tile mode no tile is split; with the tile-aligned grid no CTA holds pieces of two tiles;
//   6. every product stays below 2^31 when units * grid does (the 32-bit arithmetic the kernel uses).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../ferrum-infer-rs_b200/csrc/sched.cuh"

using namespace fb2::sched;

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static unsigned rnd(unsigned n) {
  rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
  return (unsigned)((rng_state >> 33) % n);
}

#define CHECK(cond, ...)                                                    \
  do {                                                                      \
    if (!(cond)) {                                                          \
      fprintf(stderr, "FAILED %s (line %d): ", #cond, __LINE__);            \
      fprintf(stderr, __VA_ARGS__);                                         \
      fprintf(stderr, "\n");                                                \
      exit(1);                                                              \
    }                                                                       \
  } while (0)

struct SegRec { int cta, blk, tile, r0, r1, at_start, last; };

static long long cases = 0, split_pieces = 0, whole_plans = 0, aligned_plans = 0;

static void check_plan(int row_blocks, int n_tiles, int upt, int slots, bool routed, bool align) {
  const long long units = (long long)row_blocks * n_tiles * upt;
  GridPlan gp = plan_grid(units, upt, slots, routed, align);
  int G = gp.grid;
  bool whole = gp.whole_tiles != 0;
  if (units * G >= (1LL << 31)) return;                       // launch_tc_gemm chunks or refuses these
  const u32 T = (u32)units;
  if (routed) {                                               // the in-kernel reduction of the grid (tc_body)
    whole = whole_tile_mode(upt, T / (u32)upt, G);
    const u32 cap = whole ? T / (u32)upt : T / 2;
    const u32 g_eff = cap > 1 ? cap : 1;
    if (g_eff < (u32)G) G = (int)g_eff;
  }
  cases++;
  whole_plans += whole ? 1 : 0;
  CHECK(G >= 1 && G <= slots, "G=%d slots=%d", G, slots);
  std::vector<u32> start(G + 1);
  for (int c = 0; c <= G; c++) {
    const unsigned long long exact = whole ? ((unsigned long long)c * (T / upt) / G) * upt : (unsigned long long)c * T / G;
    start[c] = whole ? tile_range_start((u32)c, T / (u32)upt, (u32)G, (u32)upt) : range_start((u32)c, T, (u32)G);
    CHECK(start[c] == exact, "32-bit overflow: c=%d T=%u G=%d", c, T, G);                               // 6
  }
  CHECK(start[0] == 0 && start[G] == T, "ends: %u %u T=%u", start[0], start[G], T);                      // 1
  for (int c = 0; c < G; c++)
    CHECK(start[c] < start[c + 1] || (routed && T < 2), "empty range: c=%d T=%u G=%d upt=%d whole=%d", c, T, G, upt, (int)whole);
  std::vector<SegRec> segs;
  for (int c = 0; c < G; c++) {
    SegIter si;
    si.init(start[c], start[c + 1], upt, n_tiles);
    Seg sg;
    u32 u = start[c];
    int nseg = 0;
    const size_t first_idx = segs.size();
    while (si.next(sg, upt, n_tiles)) {
      CHECK(sg.r0 >= 0 && sg.r0 < sg.r1 && sg.r1 <= upt, "segment bounds r0=%d r1=%d upt=%d", sg.r0, sg.r1, upt);   // 3
      const u32 expect = ((u32)sg.blk * (u32)n_tiles + (u32)sg.tile) * (u32)upt + (u32)sg.r0;
      CHECK(expect == u, "segment position: c=%d expect %u at %u", c, expect, u);
      CHECK(sg.blk < row_blocks && sg.tile < n_tiles, "tile index out of range");
      CHECK(sg.at_start == (nseg == 0), "at_start flag");
      segs.push_back(SegRec{c, sg.blk, sg.tile, sg.r0, sg.r1, sg.at_start, 0});
      u += (u32)(sg.r1 - sg.r0);
      nseg++;
    }
    CHECK(u == start[c + 1], "segments do not cover the range: c=%d", c);
    if (nseg) segs.back().last = 1;
    if (whole)
      for (size_t i = first_idx; i < segs.size(); i++)
        CHECK(segs[i].r0 == 0 && segs[i].r1 == upt, "whole-tile mode split a tile");                     // 5
    if (!whole && !routed && align && (long long)(T / upt) * 2 <= slots && G % (int)(T / upt) == 0) {
      aligned_plans += c == 0 ? 1 : 0;
      CHECK(nseg == 1, "tile-aligned grid: CTA %d holds pieces of %d tiles", c, nseg);                   // 5
    }
  }
  if (!whole) {
    // units to sample for (2): every range boundary and a few random ones
    for (int c = 0; c < G; c++) {
      CHECK(cta_of_unit(start[c], T, (u32)G) == c, "cta_of_unit(start) c=%d", c);                        // 2
      CHECK(cta_of_unit(start[c + 1] - 1, T, (u32)G) == c, "cta_of_unit(end-1) c=%d", c);
    }
    for (int i = 0; i < 16; i++) {
      const u32 u = rnd(T);
      const int c = cta_of_unit(u, T, (u32)G);
      CHECK(start[c] <= u && u < start[c + 1], "cta_of_unit(%u) = %d", u, c);
    }
    // (4): ownership of split tiles
    std::vector<int> as_last, as_other;
    for (size_t i = 0; i < segs.size(); i++) {
      const SegRec &s = segs[i];
      if (s.r0 == 0 && s.r1 == upt) continue;                 // this CTA covered the whole tile
      const u32 tau = (u32)s.blk * (u32)n_tiles + (u32)s.tile;
      const int c_first = cta_of_unit(tau * (u32)upt, T, (u32)G), c_last = cta_of_unit(tau * (u32)upt + (u32)upt - 1, T, (u32)G);
      split_pieces++;
      CHECK(c_first < c_last, "split tile with one contributor");
      CHECK(c_first <= s.cta && s.cta <= c_last, "contributor %d outside [%d, %d]", s.cta, c_first, c_last);
      if (s.cta == c_first) CHECK(s.last, "owner's piece is not the last segment of its range");
      else CHECK(s.at_start, "contributor's piece is not the first segment of its range");
      if (s.cta == c_first) CHECK(s.r0 == 0, "owner does not start the tile");
      if (s.cta == c_last) CHECK(s.r1 == upt, "last contributor does not end the tile");
      if (s.cta != c_last) CHECK(s.last, "a contributor other than the last would wait in the middle of its range");   // 4b
      if ((int)as_last.size() <= s.cta) { as_last.resize(s.cta + 1, 0); as_other.resize(s.cta + 1, 0); }
      if (s.cta == c_last) as_last[s.cta]++; else as_other[s.cta]++;
      CHECK(as_last[s.cta] <= 1 && as_other[s.cta] <= 1, "CTA %d needs more than two slot areas", s.cta);
    }
  }
}

int main() {
  const int slot_choices[] = {148, 296, 132, 264, 8, 2};
  // BASELINE.json shapes: (K/128 groups, N/128 tiles); decode (1 row block) and prefill row blocks
  const int shapes[][2] = {{32, 32}, {32, 48}, {32, 224}, {112, 32}, {16, 40}, {32, 16}, {16, 12}, {6, 16},
                           {42, 32}, {42, 16}, {32, 42}, {168, 42}, {42, 336}};
  for (auto &sh : shapes)
    for (int rb : {1, 2, 4, 16, 32})
      for (int nacc : {1, 2})
        for (int slots : {148, 296})
          for (int routed = 0; routed < 2; routed++) check_plan(rb, sh[1], sh[0] * nacc, slots, routed != 0, true);
  for (int it = 0; it < 60000; it++) {
    const int row_blocks = 1 + (int)rnd(rnd(4) == 0 ? 600 : 6);
    const int n_tiles = 1 + (int)rnd(rnd(3) == 0 ? 400 : 48);
    const int upt = 1 + (int)rnd(rnd(3) == 0 ? 336 : 12);
    const int slots = slot_choices[rnd(6)];
    check_plan(row_blocks, n_tiles, upt, slots, rnd(3) == 0, rnd(4) != 0);
  }
  printf("sched ok: %lld launch plans checked (%lld whole-tile, %lld tile-aligned), %lld pieces of split tiles\n", cases, whole_plans,
         aligned_plans, split_pieces);
  return 0;
}
