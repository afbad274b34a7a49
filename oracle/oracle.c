/*
 * oracle.c -- CPU restatement of ferrum-infer-rs's GPTQ-Int4 W4A16 reference path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (ferrum-infer-rs_b200/,
 * libfb2_w4a16.so) may call, link or import this file.  It is used by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, as
 * the checker and as the timed CPU baseline ("port").
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  The reference is Rust and cannot be compiled in this image
 * (no rustc/cargo), so this is a "port" oracle.  Pinning: the reference holds no
 * numeric golden outputs for the GPTQ GEMM itself (its tests are identity tests:
 * two independent implementations must agree); those identities are reproduced in
 * tests/test_oracle.py, and the router / MoE known-answer tests, which do hold
 * hand-computed numbers, are pinned from the JSON files under tests/golden/.
 *
 * Build: see oracle/Makefile (gcc -O2 -shared -pthread).  -ffast-math is NOT used:
 * summation order and f64 accumulation must match the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Test PRNG: crates/ferrum-quantization/tests/gptq_parity_test.rs:28-38       */
/* ------------------------------------------------------------------------- */
ORACLE_API uint32_t oracle_rnd_u32(uint64_t *state) {
  *state = (*state) * 6364136223846793005ULL + 1442695040888963407ULL;
  return (uint32_t)(*state >> 33);
}

ORACLE_API float oracle_rnd_f32(uint64_t *state, float lo, float hi) {
  float u = (float)(oracle_rnd_u32(state) & 0x00FFFFFFu) / 16777216.0f;
  return lo + u * (hi - lo);
}

/* make_synthetic: gptq_parity_test.rs:61-93.  Draw order: all qweight words, all
 * scales, then for each qzeros word 8 nibble draws low nibble first.
 * `symmetric` != 0 applies make_synthetic_symmetric (:95-103): qzeros = 0x77777777. */
ORACLE_API void oracle_make_synthetic(int64_t k, int64_t n, int64_t group_size, uint64_t seed,
                                      int symmetric, int32_t *qweight, float *scales,
                                      int32_t *qzeros) {
  uint64_t rs = seed;
  int64_t groups = k / group_size;
  for (int64_t i = 0; i < (k / 8) * n; i++) qweight[i] = (int32_t)oracle_rnd_u32(&rs);
  for (int64_t i = 0; i < groups * n; i++) scales[i] = oracle_rnd_f32(&rs, 0.01f, 0.1f);
  for (int64_t i = 0; i < groups * (n / 8); i++) {
    uint32_t word = 0;
    for (int bi = 0; bi < 8; bi++) word |= (oracle_rnd_u32(&rs) & 0xFu) << (bi * 4);
    qzeros[i] = (int32_t)word;
  }
  if (symmetric)
    for (int64_t i = 0; i < groups * (n / 8); i++) qzeros[i] = (int32_t)0x77777777u;
}

/* make_desc_act_g_idx: gptq_parity_test.rs:50-59 (balanced, non-monotonic). */
ORACLE_API void oracle_make_desc_act_g_idx(int64_t k, int64_t group_size, int32_t *g_idx) {
  int64_t groups = k / group_size;
  for (int64_t i = 0; i < k; i++) g_idx[i] = (int32_t)(i % groups);
}

/* ------------------------------------------------------------------------- */
/* Dequant: crates/ferrum-kernels/src/backend/cpu.rs:2283-2315 (g_idx == NULL, */
/* sequential groups) and crates/ferrum-quantization/src/native_safetensors.rs */
/* :1391-1420 (g_idx != NULL, act-order).  Output [n, k] row-major f32.       */
/* Returns 0, or 1 for bits != 4 (the reference returns Err(unsupported)).    */
/* ------------------------------------------------------------------------- */
ORACLE_API int oracle_dequant_gptq(const int32_t *qweight, const float *scales,
                                   const int32_t *qzeros, const int32_t *g_idx, uint32_t bits,
                                   int64_t group_size, int64_t k, int64_t n, float *w) {
  if (bits != 4) return 1;
  int64_t packed_rows = k / 8;
  for (int64_t pr = 0; pr < packed_rows; pr++) {
    for (int64_t col = 0; col < n; col++) {
      uint32_t packed = (uint32_t)qweight[pr * n + col];
      for (int bi = 0; bi < 8; bi++) {
        int64_t ki = pr * 8 + bi;
        int32_t q = (int32_t)((packed >> (bi * 4)) & 0xFu);
        int64_t grp = g_idx ? (int64_t)g_idx[ki] : ki / group_size;
        float scale = scales[grp * n + col];
        uint32_t z_packed = (uint32_t)qzeros[grp * (n / 8) + (col / 8)];
        int32_t zero = (int32_t)((z_packed >> ((col % 8) * 4)) & 0xFu) + 1;
        w[col * k + ki] = (float)(q - zero) * scale;
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* GEMM: crates/ferrum-kernels/src/backend/cpu.rs:481-492 (Linux branch):     */
/* out[i,j] = f32( sum_p f64(a[i,p]) * f64(b[j,p]) ), p ascending.            */
/* threads <= 1 reproduces the reference's single-thread loop; threads > 1    */
/* parallelises over (i,j) only -- every dot product keeps the same order, so */
/* the result is bit-identical for any thread count.                          */
/* ------------------------------------------------------------------------- */
typedef struct {
  const float *a, *b;
  float *out;
  int64_t n, k, lo, hi; /* flat (i*n + j) range [lo, hi) */
} gemm_job_t;

static void *gemm_worker(void *arg) {
  gemm_job_t *job = (gemm_job_t *)arg;
  for (int64_t f = job->lo; f < job->hi; f++) {
    int64_t i = f / job->n, j = f % job->n;
    double sum = 0.0;
    const float *ar = job->a + i * job->k;
    const float *br = job->b + j * job->k;
    for (int64_t p = 0; p < job->k; p++) sum += (double)ar[p] * (double)br[p];
    job->out[f] = (float)sum;
  }
  return NULL;
}

ORACLE_API void oracle_gemm(const float *a, const float *b, float *out, int64_t m, int64_t n,
                            int64_t k, int threads) {
  int64_t total = m * n;
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if ((int64_t)threads > total) threads = (int)(total > 0 ? total : 1);
  if (threads == 1) {
    gemm_job_t job = {a, b, out, n, k, 0, total};
    gemm_worker(&job);
    return;
  }
  pthread_t tids[256];
  gemm_job_t jobs[256];
  for (int t = 0; t < threads; t++) {
    jobs[t] = (gemm_job_t){a, b, out, n, k, total * t / threads, total * (t + 1) / threads};
    pthread_create(&tids[t], NULL, gemm_worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(tids[t], NULL);
}

/* add_bias: crates/ferrum-kernels/src/backend/cpu.rs:2065-2080 */
ORACLE_API void oracle_add_bias(float *data, const float *bias, int64_t rows, int64_t cols) {
  for (int64_t r = 0; r < rows; r++)
    for (int64_t c = 0; c < cols; c++) data[r * cols + c] += bias[c];
}

/* CpuGptqLinear::forward: crates/ferrum-kernels/src/quant_linear/cpu_dequant.rs:32-52 */
ORACLE_API void oracle_linear_forward(const float *x, const float *w_nk, const float *bias,
                                      float *out, int64_t m, int64_t n, int64_t k, int threads) {
  oracle_gemm(x, w_nk, out, m, n, k, threads);
  if (bias) oracle_add_bias(out, bias, m, n);
}

/* fused_silu_mul_split: crates/ferrum-kernels/src/backend/cpu.rs:1666-1680 */
ORACLE_API void oracle_silu_mul_split(const float *gate_up, float *out, int64_t tokens,
                                      int64_t im) {
  for (int64_t t = 0; t < tokens; t++)
    for (int64_t i = 0; i < im; i++) {
      float g = gate_up[t * 2 * im + i];
      float u = gate_up[t * 2 * im + im + i];
      out[t * im + i] = (g / (1.0f + expf(-g))) * u;
    }
}

/* fused_gelu_tanh_mul_split: crates/ferrum-kernels/src/backend/cpu.rs:1682-1698 */
ORACLE_API void oracle_gelu_tanh_mul_split(const float *gate_up, float *out, int64_t tokens,
                                           int64_t im) {
  const float SQRT_2_OVER_PI = 0.79788456f;
  for (int64_t t = 0; t < tokens; t++)
    for (int64_t i = 0; i < im; i++) {
      float g = gate_up[t * 2 * im + i];
      float u = gate_up[t * 2 * im + im + i];
      float inner = SQRT_2_OVER_PI * (g + 0.044715f * g * g * g);
      out[t * im + i] = 0.5f * g * (1.0f + tanhf(inner)) * u;
    }
}

/* ------------------------------------------------------------------------- */
/* Router: crates/ferrum-models/src/moe/router.rs:113-195 (route_into).       */
/* Returns 0, or a negative code for the reference's assert! failures.        */
/* ------------------------------------------------------------------------- */
ORACLE_API int oracle_route_into(const float *logits, int64_t logits_len, int64_t batch,
                                 int64_t num_experts, int64_t top_k, int norm_topk_prob,
                                 uint32_t *expert_ids, float *expert_weights) {
  if (logits_len != batch * num_experts) return -1; /* "router logits shape mismatch" */
  if (top_k <= 0) return -2;                        /* "top_k must be > 0" */
  if (top_k > num_experts) return -3;               /* "top_k .. exceeds num_experts" */
  float *probs = (float *)malloc(sizeof(float) * (size_t)num_experts);
  for (int64_t b = 0; b < batch; b++) {
    const float *row = logits + b * num_experts;
    float max = -INFINITY;
    for (int64_t i = 0; i < num_experts; i++)
      if (row[i] > max) max = row[i];
    float sum = 0.0f;
    for (int64_t i = 0; i < num_experts; i++) {
      float e = expf(row[i] - max);
      probs[i] = e;
      sum += e;
    }
    float inv_sum = 1.0f / sum;
    for (int64_t i = 0; i < num_experts; i++) probs[i] *= inv_sum;
    float sel_sum = 0.0f;
    int64_t dst_lo = b * top_k;
    for (int64_t kk = 0; kk < top_k; kk++) {
      float best = -INFINITY;
      int64_t best_idx = 0;
      for (int64_t i = 0; i < num_experts; i++)
        if (probs[i] > best) {
          best = probs[i];
          best_idx = i;
        }
      expert_ids[dst_lo + kk] = (uint32_t)best_idx;
      expert_weights[dst_lo + kk] = best;
      sel_sum += best;
      probs[best_idx] = -INFINITY;
    }
    if (norm_topk_prob) {
      if (sel_sum > 0.0f) {
        float scale = 1.0f / sel_sum;
        for (int64_t kk = 0; kk < top_k; kk++) expert_weights[dst_lo + kk] *= scale;
      } else {
        float uniform = 1.0f / (float)top_k;
        for (int64_t kk = 0; kk < top_k; kk++) expert_weights[dst_lo + kk] = uniform;
      }
    }
  }
  free(probs);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Bucket plan: crates/ferrum-models/src/moe/dispatch.rs:1408-1461            */
/* (MoeBucketPlan::rebuild_into): counting sort of (b,k) pairs by expert,     */
/* walking pairs in (b,k) order.                                              */
/* ------------------------------------------------------------------------- */
ORACLE_API void oracle_bucket_plan(const uint32_t *expert_ids, int64_t batch, int64_t num_experts,
                                   int64_t top_k, int32_t *expert_offsets /*[E+1]*/,
                                   uint32_t *packed_token_idx /*[pairs]*/,
                                   int32_t *pairs_by_token /*[pairs]*/) {
  int64_t total = batch * top_k;
  for (int64_t e = 0; e <= num_experts; e++) expert_offsets[e] = 0;
  for (int64_t i = 0; i < total; i++) pairs_by_token[i] = -1;
  for (int64_t i = 0; i < total; i++) expert_offsets[expert_ids[i] + 1] += 1;
  for (int64_t e = 0; e < num_experts; e++) expert_offsets[e + 1] += expert_offsets[e];
  int32_t *cursors = (int32_t *)malloc(sizeof(int32_t) * (size_t)num_experts);
  memcpy(cursors, expert_offsets, sizeof(int32_t) * (size_t)num_experts);
  for (int64_t b = 0; b < batch; b++)
    for (int64_t kk = 0; kk < top_k; kk++) {
      int64_t pair_flat = b * top_k + kk;
      uint32_t eid = expert_ids[pair_flat];
      int32_t slot = cursors[eid]++;
      packed_token_idx[slot] = (uint32_t)b;
      pairs_by_token[pair_flat] = slot;
    }
  free(cursors);
}

/* ------------------------------------------------------------------------- */
/* moe_align_block_size, deterministic host statement of                      */
/* crates/ferrum-kernels/kernels/moe_align_block_size.cu:46-160 (packed-row   */
/* ids) and moe_align_block_size_pair_ids.cu:13-96 (pair ids).  The GPU       */
/* kernels claim slots with atomics, so the order of ids inside one expert's  */
/* run is unspecified there; this oracle fills them in ascending pair order,  */
/* and tests compare per-expert runs as sets.                                 */
/* ------------------------------------------------------------------------- */
ORACLE_API void oracle_moe_align_block_size(const int32_t *expert_ids_per_pair, int64_t pairs,
                                            int64_t num_experts, int64_t block_size,
                                            int64_t sorted_max_size, int pair_ids,
                                            int32_t *sorted_token_ids, int32_t *block_ids,
                                            int32_t *total_post_pad) {
  int32_t *counts = (int32_t *)calloc((size_t)num_experts, sizeof(int32_t));
  int32_t *offs = (int32_t *)calloc((size_t)num_experts + 1, sizeof(int32_t));
  int32_t *uoffs = (int32_t *)calloc((size_t)num_experts + 1, sizeof(int32_t));
  for (int64_t i = 0; i < sorted_max_size; i++) sorted_token_ids[i] = (int32_t)pairs;
  for (int64_t p = 0; p < pairs; p++) {
    int32_t e = expert_ids_per_pair[p];
    if (e >= 0 && e < num_experts) counts[e]++;
  }
  int32_t acc_pad = 0, acc_unpad = 0;
  for (int64_t e = 0; e < num_experts; e++) {
    offs[e] = acc_pad;
    uoffs[e] = acc_unpad;
    acc_pad += (int32_t)(((counts[e] + block_size - 1) / block_size) * block_size);
    acc_unpad += counts[e];
  }
  offs[num_experts] = acc_pad;
  total_post_pad[0] = acc_pad;
  int32_t *cur = (int32_t *)calloc((size_t)num_experts, sizeof(int32_t));
  for (int64_t p = 0; p < pairs; p++) {
    int32_t e = expert_ids_per_pair[p];
    if (e >= 0 && e < num_experts) {
      int32_t pos = cur[e]++;
      sorted_token_ids[offs[e] + pos] = pair_ids ? (int32_t)p : uoffs[e] + pos;
    }
  }
  int64_t total_blocks = acc_pad / block_size;
  for (int64_t b = 0; b < total_blocks; b++) {
    int64_t row = b * block_size;
    int32_t e = 0;
    for (int64_t ei = 0; ei < num_experts; ei++)
      if (offs[ei] <= row && row < offs[ei + 1]) {
        e = (int32_t)ei;
        break;
      }
    block_ids[b] = e;
  }
  free(counts);
  free(offs);
  free(uoffs);
  free(cur);
}

/* moe_combine: crates/ferrum-kernels/src/backend/capabilities.rs:684-723 (default
 * host loop; out is OVERWRITTEN) */
ORACLE_API void oracle_moe_combine(const float *packed_down, const int32_t *pairs_by_token,
                                   const float *pair_weights, float *out, int64_t batch,
                                   int64_t hidden, int64_t top_k) {
  for (int64_t i = 0; i < batch * hidden; i++) out[i] = 0.0f;
  for (int64_t b = 0; b < batch; b++)
    for (int64_t kk = 0; kk < top_k; kk++) {
      int32_t row = pairs_by_token[b * top_k + kk];
      if (row < 0) continue;
      float w = pair_weights[b * top_k + kk];
      const float *src = packed_down + (int64_t)row * hidden;
      float *dst = out + b * hidden;
      for (int64_t h = 0; h < hidden; h++) dst[h] += w * src[h];
    }
}

/* ------------------------------------------------------------------------- */
/* moe_forward_cpu: crates/ferrum-models/src/moe/dispatch.rs:2208-2288.       */
/* Experts are given as dense [E][2I,H] (gate||up rows) and [E][H,I] f32      */
/* matrices (what CpuGptqLinear holds after load-time dequant).               */
/* Returns 0, -1 for an expert id out of range.                               */
/* ------------------------------------------------------------------------- */
ORACLE_API int oracle_moe_forward_cpu(const float *x, int64_t batch, int64_t hidden,
                                      int64_t inter, int64_t top_k, const uint32_t *expert_ids,
                                      const float *expert_weights, int64_t n_experts,
                                      const float *gate_up_w /*[E][2I][H]*/,
                                      const float *down_w /*[E][H][I]*/, float *out,
                                      int threads) {
  for (int64_t i = 0; i < batch * hidden; i++) out[i] = 0.0f;
  float *gu = (float *)malloc(sizeof(float) * (size_t)(2 * inter));
  float *act = (float *)malloc(sizeof(float) * (size_t)inter);
  float *dn = (float *)malloc(sizeof(float) * (size_t)hidden);
  int rc = 0;
  for (int64_t b = 0; b < batch && rc == 0; b++) {
    const float *xb = x + b * hidden;
    for (int64_t kk = 0; kk < top_k; kk++) {
      int64_t pair = b * top_k + kk;
      int64_t e = (int64_t)expert_ids[pair];
      float w = expert_weights[pair];
      if (e >= n_experts) {
        rc = -1;
        break;
      }
      oracle_gemm(xb, gate_up_w + e * 2 * inter * hidden, gu, 1, 2 * inter, hidden, threads);
      oracle_silu_mul_split(gu, act, 1, inter);
      oracle_gemm(act, down_w + e * hidden * inter, dn, 1, hidden, inter, threads);
      float *orow = out + b * hidden;
      for (int64_t h = 0; h < hidden; h++) orow[h] += w * dn[h];
    }
  }
  free(gu);
  free(act);
  free(dn);
  return rc;
}

/* validate_cuda_marlin_desc_act_g_idx: crates/ferrum-quantization/src/native_safetensors.rs
 * :1327-1372.  0 ok; 1 group_size==0; 2 K % group_size != 0; 3 value out of range;
 * 4 unbalanced group. */
ORACLE_API int oracle_validate_desc_act_g_idx(const int32_t *g_idx, int64_t k,
                                              int64_t group_size) {
  if (group_size == 0) return 1;
  if (k % group_size != 0) return 2;
  int64_t groups = k / group_size;
  int64_t *counts = (int64_t *)calloc((size_t)groups, sizeof(int64_t));
  int rc = 0;
  for (int64_t i = 0; i < k; i++) {
    if (g_idx[i] < 0 || g_idx[i] >= groups) {
      rc = 3;
      break;
    }
    counts[g_idx[i]]++;
  }
  if (rc == 0)
    for (int64_t g = 0; g < groups; g++)
      if (counts[g] != group_size) {
        rc = 4;
        break;
      }
  free(counts);
  return rc;
}

ORACLE_API int oracle_max_threads(void) {
  long nc = sysconf(_SC_NPROCESSORS_ONLN);
  return nc > 0 ? (int)nc : 1;
}
