// common.cuh -- shared declarations of libfb2_w4a16.so (host + device).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/fb2_w4a16.h"

struct fb2_linear;   // the C handle (defined in api.cu)

namespace fb2 {

// ----------------------------------------------------------------------------- errors
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define FB2_CUDA_TRY(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      fb2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                     __LINE__);                                                         \
      return FB2_ERR_CUDA;                                                              \
    }                                                                                   \
  } while (0)

#define FB2_LAUNCH_CHECK()                                                              \
  do {                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                \
    if (_e != cudaSuccess) {                                                            \
      fb2::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),        \
                     __FILE__, __LINE__);                                               \
      return FB2_ERR_CUDA;                                                              \
    }                                                                                   \
    fb2::count_launch();                                                                \
  } while (0)

// Programmatic dependent launch is on by default; FB2_DEBUG_NO_PDL=1 turns it off (a
// measurement aid, read once -- it changes launch attributes only, never the kernels chosen).
bool pdl_enabled();

// ------------------------------------------------------------------ tiled weight format
// The "tile stripe" layout the tensor-core kernels stream (data layout):
//   for each 128-column tile t, for each 128-row group g (rows in act-order-SORTED k):
//     8192 B  packed int4: 4 chunks of 16 B per column; chunk c = 128 columns(n) x uint4.  A dequant thread
//             (column n, 64-k half h) reads chunks 2h, 2h+1 = words J = 0..7; nibble p of word J holds
//             k = 64h + 32 (p & 1) + 4J + {0,2,1,3}[p >> 1]  (tile_nibble_k below): even nibbles ("lo class",
//             bits 3:0 / 19:16 after an optional shift by 8) fill MMA k-steps 0,1 of the half, odd nibbles
//             ("hi class", bits 7:4 / 23:20) k-steps 2,3 -- so `w & 0x000f000f` and `w & 0x00f000f0` are
//             half2 pairs of consecutive k, and every k-step of the UMMA A operand is class-pure (the decode
//             kernel feeds the raw nibbles to the tensor core as fp16 subnormals, q * 2^-24 / q * 2^-20).
//      512 B  128 x u32: lo half = f16 scale, hi half = f16 (1024 + zero)
constexpr int kTileN = 128;          // weight columns per tile  (UMMA M)
constexpr int kGroup = 128;          // k rows per stage == GPTQ group_size of the fast path
constexpr int kChunkK = 32;          // k per uint4
constexpr int kStageWBytes = kTileN * kGroup / 2;           // 8192
constexpr int kStageSZBytes = kTileN * 4;                   // 512
constexpr int kStageBytes = kStageWBytes + kStageSZBytes;   // 8704
// k offset inside the group of nibble p (0..7) of word j (0..3) of chunk c (0..3)
__host__ __device__ __forceinline__ int tile_nibble_k(int c, int j, int p) {
  const int h = c >> 1, J = (c & 1) * 4 + j;
  return 64 * h + 32 * (p & 1) + 4 * J + ((p >> 1) & 1) * 2 + (p >> 2);
}

// ------------------------------------------------------------------------ row-block map
// Every GEMM kernel works on "row blocks": up to NB activation rows that share one weight
// matrix (one expert).  Three sources describe them.
enum RowMode : int { ROWS_DENSE = 0, ROWS_DISPATCH = 1, ROWS_ROUTED = 2 };

struct DispatchEntry {  // one (expert, in_row, out_row, m) of gemm_phase_batched, <= NB rows
  int expert, in_row, out_row, m;
};
// Dispatch entries travel BY VALUE in the kernel parameter block (graph-capturable, no
// H2D copy); longer lists are launched in chunks.
constexpr int kMaxDispatchBlocks = 240;
struct DispatchTable {
  DispatchEntry e[kMaxDispatchBlocks];
};

struct RowMap {
  int mode;
  int m;                             // DENSE: number of rows
  // ROUTED (moe_align_block_size outputs)
  const int32_t *sorted_ids;         // [capacity]
  const int32_t *block_experts;      // [capacity / block]
  const int32_t *num_post_padded;    // [1]
  int block;                         // moe_block_size
  int top_k;
  int num_valid;                     // prob_m * top_k: ids >= this are padding
  // DISPATCH: entries come from the DispatchTable kernel parameter
  int num_dispatch;
  int max_blocks;                    // host-side upper bound of the number of row blocks
};

// Weight descriptor shared by all kernels (dense linear, expert view, or stack).
struct WeightRef {
  // tiled format (fast path); null when the layer only has the raw format
  const uint8_t *tiles;              // [experts][n_tiles_total][groups][kStageBytes]
  const int32_t *perm;               // [K] sorted position -> disk row, or null
  const int32_t *inv_perm;           // [K] disk row -> sorted position (stacks only; a producer's column map), or null
  // raw format (generic path)
  const int32_t *qweight;            // [experts][K/8][ld]
  const __half *scales;              // [experts][G][ld]
  const int32_t *qzeros;             // [experts][G][ld/8]
  const int32_t *g_idx;              // [K] or null
  const float *bias;                 // [n] or null (dense / expert view only)
  int k, n;                          // logical GEMM K and N of one expert (or view)
  int ld;                            // columns per expert in storage (>= col0 + n)
  int col0;                          // first column of the view inside the expert
  int group_size;
  int num_experts;
};

// ---------------------------------------------------------------------- kernel launchers
// generic_kernels.cu
int launch_generic_gemm(const WeightRef &w, const RowMap &rows, const __half *x, __half *y,
                        int epilogue /*0 none, 1 silu_mul*/, cudaStream_t stream);
int launch_generic_gemm_dispatch(const WeightRef &w, const RowMap &rows, const DispatchTable &tbl,
                                 const __half *x, __half *y, cudaStream_t stream);
int launch_dequant(const WeightRef &w, int expert, float *out_f32, __half *out_f16,
                   cudaStream_t stream);
int launch_permute_cols(const __half *x, const int32_t *perm, __half *xp, int m, int k,
                        cudaStream_t stream);
int launch_silu_mul_split(const __half *gu, __half *out, size_t tokens, size_t im, int gelu,
                          cudaStream_t stream);
int launch_add_rms_norm(const __half *input, __half *residual, const __half *weight, __half *output,
                        const int32_t *out_perm, int tokens, int hidden, float eps, cudaStream_t stream);
int launch_argmax_rows(const __half *logits, int m, int n, int32_t *out, cudaStream_t stream);
// moe_kernels.cu
int launch_route_topk_softmax(const __half *logits, int32_t *ids, float *wts, int batch,
                              int num_experts, int top_k, int norm, cudaStream_t stream);
// routing + block alignment (pair ids) of a decode batch in one single-CTA launch
bool route_align_small_supported(int batch, int num_experts, int top_k);
// x / perm / xp non-null: CTAs 1..batch of the same launch also gather xp[r, j] = x[r, perm[j]] (act-order expert stacks)
int launch_route_align_small(const __half *logits, int32_t *ids, float *wts, int32_t *sorted_ids,
                             int32_t *block_ids, int32_t *total_post_pad, int batch, int num_experts, int top_k,
                             int norm, int block_size, int sorted_max, cudaStream_t stream,
                             const __half *x = nullptr, const int32_t *perm = nullptr, __half *xp = nullptr, int hidden = 0);
// expert parallelism, sender side: routing + dispatch plan in one launch; row scatter into the send buffer
bool ep_route_plan_supported(int batch, int num_experts, int top_k, int world);
int launch_ep_route_plan(const __half *logits, int32_t *ids, float *wts, int32_t *slot_of_pair, int32_t *send_meta,
                         int batch, int num_experts, int top_k, int norm, int world, int cap, cudaStream_t stream);
int launch_ep_scatter_rows(const __half *x, const int32_t *slot_of_pair, __half *send_x, int pairs, int hidden, int top_k,
                           cudaStream_t stream);
int launch_build_pairs(const int32_t *expert_ids, int32_t *pairs_by_token,
                       int32_t *packed_token_idx, int32_t *expert_offsets, int pairs,
                       int num_experts, int top_k, cudaStream_t stream);
int launch_align_block_size(const int32_t *expert_ids, int32_t *sorted_ids, int32_t *block_ids,
                            int32_t *total_post_pad, int pairs, int num_experts, int block_size,
                            int sorted_max, int pair_ids, cudaStream_t stream);
int launch_moe_combine(const __half *packed_down, const int32_t *pairs_by_token,
                       const float *weights, __half *out, int batch, int hidden, int top_k,
                       cudaStream_t stream);
// Fused all-reduce of a row-parallel linear (parallel.cu owns the buffers; tc_gemm.cu pushes).
constexpr int kMaxRanks = 8;
struct PushTargets {        // kernel parameter of the pushing GEMM; world == 0: plain store to y
  __half *data[kMaxRanks];           // rank r's receive buffer [2 epochs][world][slot_elems] (peer-mapped)
  uint32_t *flags[kMaxRanks];        // rank r's arrival counters [2 epochs][world]
  const uint32_t *epoch;             // this rank's call counter (device memory)
  size_t slot_elems;
  int world, rank;
};
struct PushExpect { uint32_t ctas[kMaxRanks]; };   // signalling CTAs per source rank
struct P2pLocal {           // this rank's side of the exchange
  __half *data; uint32_t *flags, *epoch, *done;
  size_t slot_elems;
  int world;
};
int forward_linear_push(const ::fb2_linear *lin, cudaStream_t stream, const __half *x, __half *y_unused, size_t m,
                        const PushTargets *push, uint32_t *ctas);
int launch_p2p_reduce(const P2pLocal &loc, const PushExpect &expect, __half *y, size_t elems, cudaStream_t stream);

// tc_gemm.cu
struct TcScratch {
  float *partials;                   // split-K partial sums
  size_t partial_bytes;
  int32_t *counters;                 // self-resetting tile arrival counters
  size_t num_counters;
  uint32_t *ll;                      // flag-in-value partial slots of small-batch split tiles (tc_gemm.cu)
  int ll_ctas;                       // CTAs the area serves; 0 = exchange disabled
};
#ifndef FB2_LL_MAX_CTAS
#define FB2_LL_MAX_CTAS 1536
#endif
constexpr int kLlMaxCtas = FB2_LL_MAX_CTAS;                 // tc_body: one area per CTA; the decode kernel: up to four.  (NOT 2048: a 16 MB area
                                                 // shifted the weight allocations behind it so that the widest decode GEMM lost 2.4 us: measured, cause not pinned down)
constexpr size_t kLlBytes = (size_t)kLlMaxCtas * 2 * 8 * 128 * sizeof(uint32_t);   // [ctas][2 acc][8 rows][128]
bool tc_supported(const WeightRef &w);
// act-order launches the decode kernel gathers itself (no permute launch, w.perm stays set)
bool dk_gather_ok(const WeightRef &w, const RowMap &rows, const DispatchTable *tbl, int epilogue);
// routed launches the decode kernel serves (and therefore honours the DK_FLAG_* chain hints below)
bool dk_routed_ok(const WeightRef &w, const RowMap &rows, int epilogue);
constexpr int DK_FLAG_META_EARLY = 1;     // == dk::DKF_META_EARLY (tc_decode.cuh)
constexpr int DK_FLAG_LATE_TRIGGER = 2;   // == dk::DKF_LATE_TRIGGER
// Split-K scratch one launch with these rows needs (0 / 0 when it does not split).
void tc_requirements(const WeightRef &w, const RowMap &rows, int epilogue, int sm_count,
                     size_t *partial_bytes, size_t *counters);
// tbl != nullptr  <=>  rows.mode == ROWS_DISPATCH
// x_rows: rows of x that may be addressed (bounds of the activation tensor map)
int launch_tc_gemm(const WeightRef &w, const RowMap &rows, const DispatchTable *tbl, const __half *x,
                   size_t x_rows, __half *y, int epilogue, const TcScratch &scratch, int sm_count,
                   cudaStream_t stream, const PushTargets *push = nullptr, int dk_flags = 0,
                   const int32_t *out_map = nullptr);
// CTAs the launch above uses (each signals the peers once when it pushes)
int tc_grid_size(const WeightRef &w, const RowMap &rows, int epilogue, int sm_count);
const char *tc_kernel_name(const WeightRef &w, size_t m);

}  // namespace fb2
