// moe_kernels.cu -- MoE routing and post-op kernels (BackendMoeFused surface).
//
// Semantics follow the reference kernels; the implementations are new:
//   router        crates/ferrum-kernels/kernels/moe_router.cu:32        (expf instead of __expf so
//                 the result tracks the host router, ferrum-models/src/moe/router.rs:113-195)
//   build_pairs   crates/ferrum-kernels/kernels/moe_build_pairs.cu:24
//   align         crates/ferrum-kernels/kernels/moe_align_block_size.cu:46, _pair_ids.cu:13
//   combine       crates/ferrum-kernels/kernels/moe_combine.cu:29,62
// The reference sorts with shared-memory atomics in ONE block, so the order of pairs
// inside an expert's run depends on scheduling.  Here one CTA per expert ranks its pairs
// with an ordered block scan: the output is deterministic (ascending pair id, the order
// the host plan MoeBucketPlan::rebuild_into produces, moe/dispatch.rs:1408-1461) and the
// sort scales over SMs instead of one.
#include "common.cuh"

namespace fb2 {
namespace {

constexpr int kSortThreads = 256;
constexpr int kMaxExperts = 1024;

// Ordered rank of this thread's matches inside the block.  Returns the exclusive prefix of
// `mine` over threads and the block total.
__device__ __forceinline__ int block_exclusive_scan(int mine, int *smem_warp, int &total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int v = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  if (lane == 31) smem_warp[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int wv = lane < (int)(blockDim.x >> 5) ? smem_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, wv, o);
      if (lane >= o) wv += t;
    }
    smem_warp[32 + lane] = wv;  // inclusive over warps
  }
  __syncthreads();
  const int warp_base = warp ? smem_warp[32 + warp - 1] : 0;
  total = smem_warp[32 + (blockDim.x >> 5) - 1];
  __syncthreads();
  return warp_base + v - mine;
}

__global__ void __launch_bounds__(32)
router_kernel(const __half *__restrict__ logits, int32_t *__restrict__ out_ids,
              float *__restrict__ out_w, int batch, int num_experts, int top_k, int norm) {
  extern __shared__ float probs[];
  // PDL: let the next kernel start its prologue; wait for the previous kernel's results
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int row = blockIdx.x, lane = threadIdx.x;
  if (row >= batch) return;
  float tmax = -INFINITY;
  for (int i = lane; i < num_experts; i += 32) {
    float v = __half2float(logits[(size_t)row * num_experts + i]);
    probs[i] = v;
    tmax = v > tmax ? v : tmax;
  }
  for (int o = 16; o; o >>= 1) {
    float t = __shfl_xor_sync(0xffffffffu, tmax, o);
    tmax = t > tmax ? t : tmax;
  }
  __syncwarp();
  // The host router sums exp() left to right in one f32 accumulator (router.rs:150-155);
  // lane 0 does the same so that the normaliser is the same number.
  float sum = 0.f;
  for (int i = lane; i < num_experts; i += 32) probs[i] = expf(probs[i] - tmax);
  __syncwarp();
  if (lane == 0)
    for (int i = 0; i < num_experts; i++) sum += probs[i];
  sum = __shfl_sync(0xffffffffu, sum, 0);
  const float inv = 1.0f / sum;
  for (int i = lane; i < num_experts; i += 32) probs[i] *= inv;
  __syncwarp();

  float sel_sum = 0.f;
  for (int k = 0; k < top_k; k++) {
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int i = lane; i < num_experts; i += 32) {
      float v = probs[i];
      if (v > best) { best = v; best_i = i; }
    }
    // max value, ties -> smaller index
    for (int o = 16; o; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (best_i == 0x7fffffff) best_i = 0;  // all -inf: reference keeps best_idx = 0
    if (lane == 0) {
      out_ids[(size_t)row * top_k + k] = best_i;
      out_w[(size_t)row * top_k + k] = best;
      probs[best_i] = -INFINITY;
    }
    sel_sum += best;
    __syncwarp();
  }
  if (norm && lane < top_k) {
    // K <= 32 handled by one pass; larger K loops
    for (int k = lane; k < top_k; k += 32) {
      if (sel_sum > 0.f) out_w[(size_t)row * top_k + k] *= (1.0f / sel_sum);
      else out_w[(size_t)row * top_k + k] = 1.0f / (float)top_k;
    }
  }
}

// grid = num_experts CTAs.  CTA e ranks the pairs routed to expert e.
__global__ void __launch_bounds__(kSortThreads)
build_pairs_kernel(const int32_t *__restrict__ expert_ids, int32_t *__restrict__ pairs_by_token,
                   int32_t *__restrict__ packed_token_idx, int32_t *__restrict__ expert_offsets,
                   int pairs, int num_experts, int top_k) {
  __shared__ int warp_sums[64];
  // PDL: let the next kernel start its prologue; wait for the previous kernel's results
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int e = blockIdx.x;
  const int per = (pairs + kSortThreads - 1) / kSortThreads;
  const int p0 = min(pairs, (int)threadIdx.x * per), p1 = min(pairs, p0 + per);
  int lt = 0, eq = 0;
  for (int p = p0; p < p1; p++) {
    int id = expert_ids[p];
    bool valid = id >= 0 && id < num_experts;
    lt += (valid && id < e);
    eq += (valid && id == e);
    if (e == 0 && !valid) pairs_by_token[p] = -1;
  }
  int total_lt, total_eq;
  block_exclusive_scan(lt, warp_sums, total_lt);
  int my_base = block_exclusive_scan(eq, warp_sums, total_eq);
  if (threadIdx.x == 0) {
    expert_offsets[e] = total_lt;
    if (e == num_experts - 1) expert_offsets[num_experts] = total_lt + total_eq;
  }
  int slot = total_lt + my_base;
  for (int p = p0; p < p1; p++) {
    if (expert_ids[p] == e) {
      pairs_by_token[p] = slot;
      packed_token_idx[slot] = p / top_k;
      slot++;
    }
  }
}

// grid = num_experts CTAs.  Every CTA builds the full histogram (cheap: pairs ints from L2)
// so it can place its own expert's padded run without a second launch.
__global__ void __launch_bounds__(kSortThreads)
align_block_size_kernel(const int32_t *__restrict__ expert_ids,
                        int32_t *__restrict__ sorted_ids, int32_t *__restrict__ block_ids,
                        int32_t *__restrict__ total_post_pad, int pairs, int num_experts,
                        int block_size, int sorted_max, int pair_ids) {
  __shared__ int hist[kMaxExperts];
  __shared__ int warp_sums[64];
  __shared__ int s_pad_off, s_unpad_off, s_total_pad;
  // PDL: let the next kernel start its prologue; wait for the previous kernel's results
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int e = blockIdx.x;
  for (int i = threadIdx.x; i < num_experts; i += kSortThreads) hist[i] = 0;
  __syncthreads();
  const int per = (pairs + kSortThreads - 1) / kSortThreads;
  const int p0 = min(pairs, (int)threadIdx.x * per), p1 = min(pairs, p0 + per);
  int eq = 0;
  for (int p = p0; p < p1; p++) {
    int id = expert_ids[p];
    if (id >= 0 && id < num_experts) {
      atomicAdd(&hist[id], 1);  // counts only: order-independent
      eq += (id == e);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int pad = 0, unpad = 0, tot = 0;
    for (int i = 0; i < num_experts; i++) {
      int padded = ((hist[i] + block_size - 1) / block_size) * block_size;
      if (i < e) { pad += padded; unpad += hist[i]; }
      tot += padded;
    }
    s_pad_off = pad; s_unpad_off = unpad; s_total_pad = tot;
  }
  __syncthreads();
  int total_eq;
  int my_base = block_exclusive_scan(eq, warp_sums, total_eq);
  const int pad_off = s_pad_off, unpad_off = s_unpad_off, total_pad = s_total_pad;
  const int padded_e = ((total_eq + block_size - 1) / block_size) * block_size;
  int rank = my_base;
  for (int p = p0; p < p1; p++) {
    if (expert_ids[p] == e) {
      int slot = pad_off + rank;
      if (slot < sorted_max) sorted_ids[slot] = pair_ids ? p : unpad_off + rank;
      rank++;
    }
  }
  // padding slots of this expert's run -> sentinel
  for (int i = total_eq + threadIdx.x; i < padded_e; i += kSortThreads)
    if (pad_off + i < sorted_max) sorted_ids[pad_off + i] = pairs;
  // everything is clamped to the caller's capacity: an undersized sorted_max truncates the plan (the GEMM then
  // sees only whole blocks below it) instead of writing past block_ids / walking past sorted_ids
  const int max_blocks = sorted_max / block_size;
  for (int b = threadIdx.x; b < padded_e / block_size; b += kSortThreads)
    if (pad_off / block_size + b < max_blocks) block_ids[pad_off / block_size + b] = e;
  if (e == num_experts - 1) {
    for (int i = total_pad + threadIdx.x; i < sorted_max; i += kSortThreads) sorted_ids[i] = pairs;
    if (threadIdx.x == 0) total_post_pad[0] = min(total_pad, max_blocks * block_size);
  }
}

// Decode batches (<= 32 tokens): routing AND block alignment in ONE single-CTA launch.
// Warp r routes token r exactly as router_kernel does (same arithmetic, same tie rule), then the
// block builds the moe_align_block_size outputs with pair ids from a per-expert bit mask of the
// tokens that selected it: the rank of pair (r, e) inside expert e's run is popc(mask[e] & ((1 << r) - 1)),
// i.e. ascending pair order (the deterministic order of the host plan), with no sort and no
// cross-CTA scan.  Replaces two dependent launches (router_kernel + align_block_size_kernel).
constexpr int kRouteAlignThreads = 1024;
template <int PER>                                    // experts per lane: num_experts <= 32 * PER
__global__ void __launch_bounds__(kRouteAlignThreads)
route_align_small_kernel(const __half *__restrict__ logits, int32_t *__restrict__ out_ids, float *__restrict__ out_w,
                         int32_t *__restrict__ sorted_ids, int32_t *__restrict__ block_ids,
                         int32_t *__restrict__ total_post_pad, int batch, int num_experts, int top_k, int norm,
                         int block_size, int sorted_max, const __half *__restrict__ x, const int32_t *__restrict__ perm,
                         __half *__restrict__ xp, int hidden) {
  extern __shared__ float s_probs[];                  // [32 warps][num_experts]
  __shared__ unsigned s_mask[kMaxExperts];            // tokens that selected expert e
  __shared__ int s_off[kMaxExperts + 1];              // padded start of expert e's run
  __shared__ int s_ids[32 * 32];                      // expert of pair (row, k) at [row * 32 + k]
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (blockIdx.x > 0) {
    // act-order expert stack: CTA r + 1 gathers row r of the activations into sorted-k order while CTA 0 routes
    // (reference: gather_columns.cu:15 -- a separate launch in front of every grouped GEMM)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int r = blockIdx.x - 1;
    for (int j = threadIdx.x; j < hidden; j += (int)blockDim.x) xp[(size_t)r * hidden + j] = x[(size_t)r * hidden + perm[j]];
    return;
  }
  const int nthr = (int)blockDim.x;
  for (int i = threadIdx.x; i < num_experts; i += nthr) s_mask[i] = 0u;
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pairs = batch * top_k;
  if (warp < batch) {
    // The probabilities of a row live in registers (expert i in lane i % 32, register i / 32); shared memory only
    // serves the left-to-right sum.  Ids and weights leave through ONE store each (the earlier form rescaled the weights
    // in global memory and re-read the ids there: two L2 round trips on the critical path of a decode layer).
    const int row = warp;
    float *probs = s_probs + (size_t)warp * num_experts;
    float pv[PER];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < PER; j++) {
      const int i = j * 32 + lane;
      pv[j] = i < num_experts ? __half2float(logits[(size_t)row * num_experts + i]) : -INFINITY;
      tmax = pv[j] > tmax ? pv[j] : tmax;
    }
    for (int o = 16; o; o >>= 1) {
      float t = __shfl_xor_sync(0xffffffffu, tmax, o);
      tmax = t > tmax ? t : tmax;
    }
#pragma unroll
    for (int j = 0; j < PER; j++) {
      const int i = j * 32 + lane;
      pv[j] = expf(pv[j] - tmax);
      if (i < num_experts) probs[i] = pv[j];
    }
    __syncwarp();
    // left to right, as router.rs:150-155 -- eight independent shared-memory loads in flight per step
    float sum = 0.f;
    if (lane == 0) {
      int i = 0;
      for (; i + 8 <= num_experts; i += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = probs[i + j];
#pragma unroll
        for (int j = 0; j < 8; j++) sum += v[j];
      }
      for (; i < num_experts; i++) sum += probs[i];
    }
    sum = __shfl_sync(0xffffffffu, sum, 0);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < PER; j++) pv[j] = (j * 32 + lane < num_experts) ? pv[j] * inv : -INFINITY;
    float sel_sum = 0.f, my_w = 0.f;
    int my_id = 0;
    for (int k = 0; k < top_k; k++) {                    // top_k <= 32 (route_align_small_supported)
      // probabilities are >= 0 (removed ones -inf): their bit patterns order like signed integers, so the row maximum
      // is ONE warp reduction (redux.sync) instead of five dependent shuffle pairs; the lowest index holding it wins
      // (the reference's tie rule): lowest register first, lowest lane inside it
      float lbest = pv[0];
#pragma unroll
      for (int j = 1; j < PER; j++) lbest = fmaxf(lbest, pv[j]);
      const float best = __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(lbest)));
      int best_i = 0;
      bool found = false;
#pragma unroll
      for (int j = 0; j < PER; j++) {
        const unsigned hit = __ballot_sync(0xffffffffu, pv[j] == best);
        if (!found && hit) { best_i = j * 32 + (__ffs(hit) - 1); found = true; }
      }
#pragma unroll
      for (int j = 0; j < PER; j++)
        if (best_i == j * 32 + lane) pv[j] = -INFINITY;
      if (lane == 0) atomicOr(&s_mask[best_i], 1u << row);
      if (k == lane) { my_w = best; my_id = best_i; }    // lane k keeps pair k
      sel_sum += best;
    }
    if (lane < top_k) {
      float w = my_w;
      if (norm) w = sel_sum > 0.f ? w * (1.0f / sel_sum) : 1.0f / (float)top_k;
      out_ids[(size_t)row * top_k + lane] = my_id;
      out_w[(size_t)row * top_k + lane] = w;
      s_ids[row * 32 + lane] = my_id;
    }
  }
  __syncthreads();
  // padded offsets: warp 0 scans the experts 32 at a time
  if (warp == 0) {
    int carry = 0;
    for (int e0 = 0; e0 < num_experts; e0 += 32) {
      const int e = e0 + lane;
      const int cnt = e < num_experts ? __popc(s_mask[e]) : 0;
      const int padded = ((cnt + block_size - 1) / block_size) * block_size;
      int incl = padded;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (e < num_experts) s_off[e] = carry + incl - padded;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) { s_off[num_experts] = carry; total_post_pad[0] = min(carry, (sorted_max / block_size) * block_size); }
  }
  __syncthreads();
  // every slot of the padded runs is written exactly once: a pair id or the sentinel
  const int total = min(s_off[num_experts], sorted_max);
  for (int e = threadIdx.x; e < num_experts; e += nthr) {
    const int cnt = __popc(s_mask[e]);
    for (int i = s_off[e] + cnt; i < s_off[e + 1]; i++)
      if (i < total) sorted_ids[i] = pairs;
    for (int b = s_off[e] / block_size; b < s_off[e + 1] / block_size; b++)
      if (b < sorted_max / block_size) block_ids[b] = e;
  }
  for (int pidx = threadIdx.x; pidx < pairs; pidx += nthr) {
    const int row = pidx / top_k;
    const int e = s_ids[row * 32 + pidx % top_k];
    const int slot = s_off[e] + __popc(s_mask[e] & ((1u << row) - 1u));
    if (slot < sorted_max) sorted_ids[slot] = pidx;
  }
}

// Expert parallelism, sender side: route this rank's tokens and plan the dispatch in ONE single-CTA
// launch.  Pair p = (token, k) goes to rank dest = id / e_loc, into slot (ordered count of earlier
// pairs with the same destination) of the fixed-capacity send buffer [world][cap]; send_meta holds
// the expert id local to the destination (-1 = empty slot).  slot_of_pair feeds the combine.
__global__ void __launch_bounds__(kRouteAlignThreads)
ep_route_plan_kernel(const __half *__restrict__ logits, int32_t *__restrict__ out_ids, float *__restrict__ out_w,
                     int32_t *__restrict__ slot_of_pair, int32_t *__restrict__ send_meta, int batch, int num_experts,
                     int top_k, int norm, int world, int cap) {
  extern __shared__ float s_probs[];                  // [32 warps][num_experts]
  __shared__ int s_dest_id[kRouteAlignThreads];       // expert id of pair p
  __shared__ int warp_sums[64];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kRouteAlignThreads / 32;
  const int pairs = batch * top_k, e_loc = num_experts / world;
  for (int i = threadIdx.x; i < world * cap; i += kRouteAlignThreads) send_meta[i] = -1;
  for (int row = warp; row < batch; row += nwarps) {
    float *probs = s_probs + (size_t)warp * num_experts;
    float tmax = -INFINITY;
    for (int i = lane; i < num_experts; i += 32) {
      float v = __half2float(logits[(size_t)row * num_experts + i]);
      probs[i] = v;
      tmax = v > tmax ? v : tmax;
    }
    for (int o = 16; o; o >>= 1) {
      float t = __shfl_xor_sync(0xffffffffu, tmax, o);
      tmax = t > tmax ? t : tmax;
    }
    __syncwarp();
    for (int i = lane; i < num_experts; i += 32) probs[i] = expf(probs[i] - tmax);
    __syncwarp();
    // left to right, as router.rs:150-155 -- from registers: four independent shared-memory loads in flight per
    // step instead of one dependent load per add (the serial form cost 2 us of a 128-expert row)
    float sum = 0.f;
    if (lane == 0) {
      int i = 0;
      for (; i + 8 <= num_experts; i += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = probs[i + j];
#pragma unroll
        for (int j = 0; j < 8; j++) sum += v[j];
      }
      for (; i < num_experts; i++) sum += probs[i];
    }
    sum = __shfl_sync(0xffffffffu, sum, 0);
    const float inv = 1.0f / sum;
    for (int i = lane; i < num_experts; i += 32) probs[i] *= inv;
    __syncwarp();
    float sel_sum = 0.f;
    for (int k = 0; k < top_k; k++) {
      float best = -INFINITY;
      int best_i = 0x7fffffff;
      for (int i = lane; i < num_experts; i += 32) {
        float v = probs[i];
        if (v > best) { best = v; best_i = i; }
      }
      for (int o = 16; o; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
      }
      if (best_i == 0x7fffffff) best_i = 0;
      if (lane == 0) {
        out_ids[(size_t)row * top_k + k] = best_i;
        out_w[(size_t)row * top_k + k] = best;
        s_dest_id[row * top_k + k] = best_i;
        probs[best_i] = -INFINITY;
      }
      sel_sum += best;
      __syncwarp();
    }
    if (norm) {
      for (int k = lane; k < top_k; k += 32) {
        if (sel_sum > 0.f) out_w[(size_t)row * top_k + k] *= (1.0f / sel_sum);
        else out_w[(size_t)row * top_k + k] = 1.0f / (float)top_k;
      }
    }
    __syncwarp();
  }
  __syncthreads();
  // ordered slot of every pair inside its destination's run: one block scan per destination rank
  const int pidx = threadIdx.x;
  const int my_id = pidx < pairs ? s_dest_id[pidx] : -1;
  const int my_dest = my_id >= 0 ? my_id / e_loc : -1;
  for (int r = 0; r < world; r++) {
    int total;
    const int rank_in = block_exclusive_scan(my_dest == r ? 1 : 0, warp_sums, total);
    if (my_dest == r) {
      const int slot = r * cap + rank_in;                        // rank_in < cap: cap = batch * top_k
      slot_of_pair[pidx] = slot;
      send_meta[slot] = my_id - r * e_loc;
    }
  }
}

// One CTA per pair: x[pair / top_k] -> send_x[slot_of_pair[pair]] (16-byte vectors).
__global__ void __launch_bounds__(128)
ep_scatter_rows_kernel(const __half *__restrict__ x, const int32_t *__restrict__ slot_of_pair,
                       __half *__restrict__ send_x, int hidden, int top_k) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int p = blockIdx.x;
  const uint4 *src = reinterpret_cast<const uint4 *>(x + (size_t)(p / top_k) * hidden);
  uint4 *dst = reinterpret_cast<uint4 *>(send_x + (size_t)slot_of_pair[p] * hidden);
  for (int i = threadIdx.x; i < hidden / 8; i += blockDim.x) dst[i] = src[i];
}

template <bool kPairs>
__global__ void combine_kernel(const __half *__restrict__ rows, const int32_t *__restrict__ pairs_by_token,
                               const float *__restrict__ weights, __half *__restrict__ out,
                               int batch, int hidden, int top_k) {
  // PDL: let the next kernel start its prologue; wait for the previous kernel's results
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int b = blockIdx.y;
  const int h2 = blockIdx.x * blockDim.x + threadIdx.x;   // pair of columns
  if (b >= batch || h2 * 2 >= hidden) return;
  float a0 = 0.f, a1 = 0.f;
  for (int k = 0; k < top_k; k++) {
    int row = kPairs ? pairs_by_token[b * top_k + k] : b * top_k + k;
    if (row < 0) continue;
    float w = weights[b * top_k + k];
    __half2 v = *reinterpret_cast<const __half2 *>(rows + (size_t)row * hidden + h2 * 2);
    a0 += w * __low2float(v);
    a1 += w * __high2float(v);
  }
  *reinterpret_cast<__half2 *>(out + (size_t)b * hidden + h2 * 2) =
      __halves2half2(__float2half_rn(a0), __float2half_rn(a1));
}

// Launch with the programmatic-stream-serialisation attribute (PDL).
template <typename... KArgs, typename... Args>
int launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FB2_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  count_launch();
  return FB2_OK;
}

}  // namespace

int launch_route_topk_softmax(const __half *logits, int32_t *ids, float *wts, int batch,
                              int num_experts, int top_k, int norm, cudaStream_t stream) {
  if (batch <= 0) return FB2_OK;
  return launch_pdl(router_kernel, dim3(batch), dim3(32), (size_t)num_experts * sizeof(float), stream, logits, ids,
                    wts, batch, num_experts, top_k, norm);
}

bool route_align_small_supported(int batch, int num_experts, int top_k) {
  return batch >= 1 && batch <= 32 && num_experts <= kMaxExperts && top_k <= num_experts && top_k <= 32 &&
         (size_t)32 * num_experts * sizeof(float) <= 160 * 1024;
}

int launch_route_align_small(const __half *logits, int32_t *ids, float *wts, int32_t *sorted_ids,
                             int32_t *block_ids, int32_t *total_post_pad, int batch, int num_experts, int top_k,
                             int norm, int block_size, int sorted_max, cudaStream_t stream, const __half *x,
                             const int32_t *perm, __half *xp, int hidden) {
  const size_t smem = (size_t)32 * num_experts * sizeof(float);
  const bool gather = x && perm && xp && hidden > 0;
  // one warp per token; at least four warps for the alignment passes (and the act-order row gather CTAs)
  const int threads = gather ? kRouteAlignThreads : std::max(128, 32 * batch);
  auto go = [&](auto kernel) -> int {
    // (the kernel also holds ~12.3 KB of static shared memory: the opt-in is needed from 48 KB in total)
    if (smem + 13 * 1024 > 48 * 1024) FB2_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    return launch_pdl(kernel, dim3(gather ? 1 + batch : 1), dim3(threads), smem, stream, logits, ids, wts, sorted_ids,
                      block_ids, total_post_pad, batch, num_experts, top_k, norm, block_size, sorted_max, x, perm, xp, hidden);
  };
  if (num_experts <= 128) return go(route_align_small_kernel<4>);
  if (num_experts <= 256) return go(route_align_small_kernel<8>);
  return go(route_align_small_kernel<kMaxExperts / 32>);
}

bool ep_route_plan_supported(int batch, int num_experts, int top_k, int world) {
  return batch >= 1 && batch * top_k <= kRouteAlignThreads && num_experts <= 256 && world >= 1 &&
         num_experts % world == 0 && top_k <= num_experts;
}

int launch_ep_route_plan(const __half *logits, int32_t *ids, float *wts, int32_t *slot_of_pair, int32_t *send_meta,
                         int batch, int num_experts, int top_k, int norm, int world, int cap, cudaStream_t stream) {
  const size_t smem = (size_t)32 * num_experts * sizeof(float);
  return launch_pdl(ep_route_plan_kernel, dim3(1), dim3(kRouteAlignThreads), smem, stream, logits, ids, wts, slot_of_pair,
                    send_meta, batch, num_experts, top_k, norm, world, cap);
}

int launch_ep_scatter_rows(const __half *x, const int32_t *slot_of_pair, __half *send_x, int pairs, int hidden, int top_k,
                           cudaStream_t stream) {
  if (pairs <= 0) return FB2_OK;
  return launch_pdl(ep_scatter_rows_kernel, dim3(pairs), dim3(128), 0, stream, x, slot_of_pair, send_x, hidden, top_k);
}

int launch_build_pairs(const int32_t *expert_ids, int32_t *pairs_by_token,
                       int32_t *packed_token_idx, int32_t *expert_offsets, int pairs,
                       int num_experts, int top_k, cudaStream_t stream) {
  return launch_pdl(build_pairs_kernel, dim3(num_experts), dim3(kSortThreads), 0, stream, expert_ids, pairs_by_token,
                    packed_token_idx, expert_offsets, pairs, num_experts, top_k);
}

int launch_align_block_size(const int32_t *expert_ids, int32_t *sorted_ids, int32_t *block_ids,
                            int32_t *total_post_pad, int pairs, int num_experts, int block_size,
                            int sorted_max, int pair_ids, cudaStream_t stream) {
  if (num_experts > kMaxExperts) {
    set_error("moe_align_block_size: num_experts %d > %d", num_experts, kMaxExperts);
    return FB2_ERR_UNSUPPORTED;
  }
  return launch_pdl(align_block_size_kernel, dim3(num_experts), dim3(kSortThreads), 0, stream, expert_ids, sorted_ids,
                    block_ids, total_post_pad, pairs, num_experts, block_size, sorted_max, pair_ids);
}

int launch_moe_combine(const __half *packed_down, const int32_t *pairs_by_token,
                       const float *weights, __half *out, int batch, int hidden, int top_k,
                       cudaStream_t stream) {
  if (batch <= 0) return FB2_OK;
  dim3 grid((hidden / 2 + 127) / 128, batch);
  if (pairs_by_token)
    return launch_pdl(combine_kernel<true>, grid, dim3(128), 0, stream, packed_down, pairs_by_token, weights, out,
                      batch, hidden, top_k);
  return launch_pdl(combine_kernel<false>, grid, dim3(128), 0, stream, packed_down, (const int32_t *)nullptr, weights,
                    out, batch, hidden, top_k);
}

}  // namespace fb2
