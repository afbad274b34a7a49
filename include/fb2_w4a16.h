/*
 * fb2_w4a16.h -- C ABI of libfb2_w4a16.so: a B200-native (sm_100a) GPTQ-Int4 W4A16
 * dequant-GEMM path that sits behind ferrum-infer-rs's own operator traits.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * reference checkout).  Conventions follow the reference's existing FFI
 * (crates/ferrum-kernels/src/backend/cuda/marlin.rs:293-388): raw pointers and sizes,
 * a CUDA stream, an `int` status (0 = ok), the caller owns every buffer it passes,
 * and no call synchronises the stream unless it says so.  `*_forward*`/`*_gemm_*` and
 * the `fb2_moe_*` launchers never allocate and never touch the host after launch, so
 * they can be captured into a CUDA graph (reference requirement:
 * crates/ferrum-kernels/src/backend/capabilities.rs:205-214).
 *
 * Device activations are IEEE fp16 (`uint16_t` bit patterns here so that the header
 * needs no CUDA include); routing indices are int32; combine weights are fp32.
 * There is no CPU fallback: every entry point that computes fails with
 * FB2_ERR_CUDA when no sm_100 device is usable.
 *
 * Threads and streams.  Handles are immutable after create and may be shared between
 * threads; create / destroy / pregrow calls are serialised inside the library; the error
 * message is thread-local.  Launches on one device share that device's scratch (split-K
 * partials, fix-up slots, the act-order gather buffer), exactly as the reference's Marlin
 * path shares its gather scratch and workspace per device (quant.rs:228-322): work
 * submitted through this library to ONE device must be ordered on ONE stream at a time
 * (or on streams the caller orders with events).  Different devices are independent.
 * The `*_forward_host*` calls run on a private stream, return after synchronising it and
 * serialise themselves with a mutex; do not overlap them with other work of this library
 * on the same device.
 */
#ifndef FB2_W4A16_H
#define FB2_W4A16_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB2_API __attribute__((visibility("default")))

typedef uint16_t fb2_half;          /* IEEE binary16 bit pattern                       */
typedef void *fb2_stream_t;         /* cudaStream_t / CUstream; NULL = default stream  */
typedef struct fb2_linear fb2_linear_t; /* one GPTQ linear      (ref: Box<dyn Linear<B>>)          */
typedef struct fb2_stack fb2_stack_t;   /* E stacked experts    (ref: Arc<dyn MarlinExpertStack<B>>) */

/* Status codes.  FerrumError::unsupported -> FB2_ERR_UNSUPPORTED, FerrumError::model ->
 * FB2_ERR_MODEL (crates/ferrum-types error kinds used by load_gptq). */
enum {
  FB2_OK = 0,
  FB2_ERR_SHAPE = -1,       /* a size/stride argument violates the documented contract */
  FB2_ERR_UNSUPPORTED = -2, /* valid GPTQ, but outside what this path implements        */
  FB2_ERR_CUDA = -3,        /* CUDA runtime/driver error, or no sm_100 device           */
  FB2_ERR_NCCL = -4,
  FB2_ERR_ARG = -5,         /* NULL / inconsistent pointer arguments                    */
  FB2_ERR_MODEL = -6        /* malformed tensors (g_idx out of range, length mismatch)  */
};

/* Library version string, e.g. "fb2-w4a16 0.1.0 (sm_100a)". */
FB2_API const char *fb2_version(void);
/* Thread-local message of the last failing call on this thread ("" if none). */
FB2_API const char *fb2_last_error(void);
/* Number of kernels this library has launched since load / since the last reset
 * (bench.py's gpu_launches claim is read from here). */
FB2_API uint64_t fb2_launch_count(void);
FB2_API void fb2_reset_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Dense GPTQ linear.
 * Replaces BackendQuantMarlin::load_gptq (crates/ferrum-kernels/src/backend/
 * capabilities.rs:136; CUDA impl crates/ferrum-kernels/src/backend/cuda/quant.rs:612)
 * argument for argument, as called by GptqLinear::from_raw[_with_metadata]
 * (crates/ferrum-quantization/src/gptq.rs:58,84):
 *   qweight [K/8, N] i32   nibble i of word (pr, n) = row 8*pr+i
 *   scales  [K/g, N] f32   (f16 on disk, widened by the loader; re-rounded to f16 here,
 *                           as quant.rs:717 does)
 *   qzeros  [K/g, N/8] i32 nibble n%8 of word (g, n/8) stores zero-1
 *   g_idx   [K] i32 or NULL; non-trivial g_idx (act-order) must be balanced: every
 *           group exactly `group_size` rows (native_safetensors.rs:1327-1372)
 *   bias    [N] f32 or NULL
 * Host pointers are borrowed for the duration of the call only.  Unlike the reference's
 * Marlin path (quant.rs:676 rejects any zero code other than 7) arbitrary zero codes
 * are accepted, matching the CPU reference (backend/cpu.rs:2283-2315).
 * Errors: bits != 4 -> UNSUPPORTED; N % 8 or K % 8 or K % group_size != 0 -> SHAPE;
 * g_idx value outside [0, K/g) -> MODEL; unbalanced act-order groups -> UNSUPPORTED.
 * ---------------------------------------------------------------------------------- */
FB2_API int fb2_gptq_linear_create(const int32_t *qweight, const float *scales,
                                   const int32_t *qzeros, const int32_t *g_idx,
                                   const float *bias, uint32_t bits, size_t group_size, size_t k,
                                   size_t n, int device, fb2_linear_t **out);
FB2_API void fb2_gptq_linear_destroy(fb2_linear_t *lin);
/* Linear::in_features / out_features (crates/ferrum-kernels/src/linear.rs:107-108). */
FB2_API size_t fb2_gptq_linear_in_features(const fb2_linear_t *lin);
FB2_API size_t fb2_gptq_linear_out_features(const fb2_linear_t *lin);
/* 1 when the layer was loaded with a non-trivial g_idx (quant.rs:685-711). */
FB2_API int fb2_gptq_linear_is_act_order(const fb2_linear_t *lin);
/* Name of the kernel family forward() will use for `m` rows ("tc-streamk": the tensor-core kernels of up to 64-row
 * blocks -- the decode kernel of csrc/tc_decode.cuh up to 32 rows; "tc-streamk-prefill": 128 / 256-row blocks;
 * "generic": shapes outside the tile format): introspection for tests and the bench, not a dispatch knob. */
FB2_API const char *fb2_gptq_linear_kernel_name(const fb2_linear_t *lin, size_t m);

/* Linear::forward (crates/ferrum-kernels/src/linear.rs:123; CUDA impl
 * quant_linear/cuda_marlin.rs:83-181): y[m, N] = x[m, K] . W^T (+ bias), fp16 in/out,
 * fp32 accumulate.  x and y are device pointers, row-major, contiguous; x must be 16-byte aligned (ARG otherwise:
 * every activation path moves 16-byte vectors).  Work is appended to `stream`; no sync, no allocation (the
 * per-device split-K scratch grows outside graph capture only, and outgrown buffers are retired, not freed, so
 * graphs captured earlier stay valid).  m == 0 is a no-op. */
FB2_API int fb2_gptq_linear_forward(const fb2_linear_t *lin, fb2_stream_t stream,
                                    const fb2_half *x, fb2_half *y, size_t m);
/* The same call for HOST buffers: copies x to the device, runs forward, copies y back
 * and synchronises; it uses a pinned staging area owned by the handle.  All host-buffer calls of one device
 * (this one and the batch form below) are serialised by one per-device mutex: they share the split-K scratch. */
FB2_API int fb2_gptq_linear_forward_host(fb2_linear_t *lin, const fb2_half *x_host,
                                         fb2_half *y_host, size_t m);

/* Host-buffer batch form of the same call: y_host[i] = lins[i].forward(x_host[i], m) for
 * i < n (all on one device).  The linears are processed in chunks of 16 on three streams: the H2D copies of chunk
 * c+1 and the D2H copies of chunk c-1 overlap the GEMMs of chunk c; page-locked caller buffers are DMA'd directly,
 * pageable ones are staged through a pinned arena; the call returns after one synchronisation.  This is what a
 * host-side caller that owns host tensors (the reference's CPU-backend tests, gptq_parity_test.rs:109-145) would
 * call, and what the end-to-end bench times. */
FB2_API int fb2_gptq_linear_forward_host_batch(fb2_linear_t *const *lins, size_t n,
                                               const fb2_half *const *x_host,
                                               fb2_half *const *y_host, size_t m);
/* The same with ms[i] rows for entry i (a decode step over requests of different batch sizes in ONE call: the copy
 * pipeline fills and drains once).  Repeated calls with the same linears and rows replay cached launch graphs. */
FB2_API int fb2_gptq_linear_forward_host_batch_m(fb2_linear_t *const *lins, size_t n,
                                                 const fb2_half *const *x_host,
                                                 fb2_half *const *y_host, const size_t *ms);

/* Dequantise the stored weights back to the reference's [N, K] row-major matrix, in
 * ORIGINAL (disk) k order, exactly as cpu_dequant_gptq / dequantize_gptq_with_g_idx
 * produce it (backend/cpu.rs:2283; native_safetensors.rs:1391).  f32 output is
 * bit-exact; f16 output equals f16_rn of it.  Device pointers. */
FB2_API int fb2_gptq_linear_dequant_f32(const fb2_linear_t *lin, fb2_stream_t stream,
                                        float *w_nk);
FB2_API int fb2_gptq_linear_dequant_f16(const fb2_linear_t *lin, fb2_stream_t stream,
                                        fb2_half *w_nk);

/* Algorithmic bytes / flops of one forward (SURVEY.md 8d formula; used for roofline
 * accounting only). */
FB2_API uint64_t fb2_gemm_algorithmic_bytes(size_t k, size_t n, size_t m, size_t group_size,
                                            int act_order);
FB2_API uint64_t fb2_gemm_flops(size_t k, size_t n, size_t m);

/* BackendQuantMarlin::pregrow_marlin_gather_scratch (capabilities.rs:212): make sure
 * the per-device scratch (act-order permuted activations, split-K partials) can serve
 * forwards of up to `max_m` rows on every linear/stack created so far on `device`.
 * Must be called outside graph capture; create() already sizes it for m <= 64. */
FB2_API int fb2_pregrow_scratch(int device, size_t max_m);

/* ------------------------------------------------------------------------------------
 * Stacked MoE experts.
 * Replaces BackendQuantMarlin::load_gptq_stacked (capabilities.rs:180; CUDA impl
 * quant.rs:768) and the MarlinExpertStack<B> trait
 * (crates/ferrum-kernels/src/marlin_expert_stack.rs:31-109).
 * qweights[e] / scales[e] / qzeros[e]: expert e's raw tensors, all [K, n_per_expert];
 * one shared g_idx (or NULL).
 * ---------------------------------------------------------------------------------- */
FB2_API int fb2_gptq_stack_create(const int32_t *const *qweights, const float *const *scales,
                                  const int32_t *const *qzeros, const int32_t *g_idx,
                                  uint32_t bits, size_t group_size, size_t k,
                                  size_t n_per_expert, size_t num_experts, int device,
                                  fb2_stack_t **out);
FB2_API void fb2_gptq_stack_destroy(fb2_stack_t *st);
FB2_API size_t fb2_gptq_stack_n_per_expert(const fb2_stack_t *st);
FB2_API size_t fb2_gptq_stack_k(const fb2_stack_t *st);
FB2_API size_t fb2_gptq_stack_num_experts(const fb2_stack_t *st);
/* MarlinExpertStack::zero_workspace (marlin_expert_stack.rs:57): this design has no
 * lock workspace; kept as a no-op for trait parity. */
FB2_API int fb2_gptq_stack_zero_workspace(const fb2_stack_t *st, fb2_stream_t stream);
/* MarlinExpertStack::make_expert_linear (marlin_expert_stack.rs:103): a Linear view of
 * columns [expert_offset, expert_offset + expert_n) of the stacked [K, E*n_per_expert]
 * matrix.  The view borrows the stack's device memory: destroy it before the stack.
 * expert_offset and expert_n must be multiples of 128 and stay inside one expert. */
FB2_API int fb2_gptq_stack_make_expert_linear(const fb2_stack_t *st, size_t expert_offset,
                                              size_t expert_n, const float *bias,
                                              fb2_linear_t **out);
/* MarlinExpertStack::gemm_phase_batched (marlin_expert_stack.rs:64-70): for each
 * dispatch (expert, in_row, out_row, m): y[out_row .. +m, :n_per_expert] =
 * x[in_row .. +m, :K] . W_expert^T.  `dispatches` is a HOST array of nd 4-tuples of
 * int64; it is read during the call only (the entries travel by value in the kernel parameters, <= 240 per
 * launch: no upload, no sync, graph-capturable). */
typedef struct {
  int64_t expert, in_row, out_row, m;
} fb2_dispatch_t;
FB2_API int fb2_gptq_stack_gemm_batched(const fb2_stack_t *st, fb2_stream_t stream,
                                        const fb2_half *x, const fb2_dispatch_t *dispatches,
                                        size_t nd, fb2_half *y, size_t k);
/* MarlinExpertStack::gemm_phase_vllm (marlin_expert_stack.rs:82-96), the routed grouped
 * GEMM keyed by moe_align_block_size outputs (device int32 arrays): for every position
 * i < num_tokens_post_padded[0] with id = sorted_token_ids[i] < prob_m*top_k:
 *   y[id, :] = x[id / top_k, :] . W_{expert_ids[i / block]}^T.
 * Unlike the reference kernel the output need not be pre-zeroed.  block in {8,16,32,64}. */
FB2_API int fb2_gptq_stack_gemm_routed(const fb2_stack_t *st, fb2_stream_t stream,
                                       const fb2_half *x, const int32_t *sorted_token_ids,
                                       const int32_t *expert_ids,
                                       const int32_t *num_tokens_post_padded, fb2_half *y,
                                       size_t prob_m, size_t block, size_t top_k,
                                       size_t sorted_capacity);
/* Same, with the SwiGLU epilogue fused (stack must be a gate||up stack, n_per_expert =
 * 2*I): act[id, i] = silu(g[id, i]) * g[id, I + i]; replaces gemm_phase_vllm +
 * B::fused_silu_mul_split (crates/ferrum-models/src/moe/dispatch.rs:1999-2048). */
FB2_API int fb2_gptq_stack_gemm_routed_silu(const fb2_stack_t *st, fb2_stream_t stream,
                                            const fb2_half *x, const int32_t *sorted_token_ids,
                                            const int32_t *expert_ids,
                                            const int32_t *num_tokens_post_padded,
                                            fb2_half *act, size_t prob_m, size_t block,
                                            size_t top_k, size_t sorted_capacity);
FB2_API int fb2_gptq_stack_dequant_f32(const fb2_stack_t *st, fb2_stream_t stream,
                                       float *w_enk);

/* ------------------------------------------------------------------------------------
 * MoE routing / post-ops.  Replace BackendMoeFused
 * (crates/ferrum-kernels/src/backend/capabilities.rs:305-723; CUDA launchers
 * crates/ferrum-kernels/src/backend/cuda/moe.rs).  All pointers are device pointers.
 * ---------------------------------------------------------------------------------- */
/* route_topk_softmax (capabilities.rs:337; kernel kernels/moe_router.cu:32): softmax over
 * E in f32, K passes of strict-> argmax (ties -> smaller index), optional renorm.  Uses
 * expf (not __expf) so ids/weights match the host router route_into
 * (crates/ferrum-models/src/moe/router.rs:113-195) on the same f16-rounded logits. */
FB2_API int fb2_moe_route_topk_softmax(fb2_stream_t stream, const fb2_half *logits,
                                       int32_t *out_ids, float *out_weights, size_t batch,
                                       size_t num_experts, size_t top_k, int norm_topk_prob);
/* moe_build_pairs_by_token (capabilities.rs:407; kernels/moe_build_pairs.cu:24).  The
 * order inside an expert's run is ascending pair id (deterministic; the reference's is
 * atomics order). */
FB2_API int fb2_moe_build_pairs_by_token(fb2_stream_t stream, const int32_t *expert_ids,
                                         int32_t *pairs_by_token, int32_t *packed_token_idx,
                                         int32_t *expert_offsets, size_t batch_x_topk,
                                         size_t num_experts, size_t top_k);
/* moe_align_block_size / _pair_ids (capabilities.rs:427,447;
 * kernels/moe_align_block_size.cu:46, moe_align_block_size_pair_ids.cu:13). */
FB2_API int fb2_moe_align_block_size(fb2_stream_t stream, const int32_t *expert_ids_per_pair,
                                     int32_t *sorted_token_ids, int32_t *block_ids,
                                     int32_t *total_tokens_post_pad, size_t batch_x_topk,
                                     size_t num_experts, size_t block_size,
                                     size_t sorted_max_size, int pair_ids);
/* moe_combine (capabilities.rs:684; kernels/moe_combine.cu:29): out[b,h] = sum_k
 * w[b,k] * packed_down[pairs_by_token[b,k], h]; row -1 = skip; out is overwritten. */
FB2_API int fb2_moe_combine(fb2_stream_t stream, const fb2_half *packed_down,
                            const int32_t *pairs_by_token, const float *pair_weights,
                            fb2_half *out, size_t batch, size_t hidden, size_t top_k);
/* weighted_sum_batched (capabilities.rs:565; kernels/moe_combine.cu:62): slots is
 * [batch, top_k, hidden]. */
FB2_API int fb2_moe_weighted_sum_batched(fb2_stream_t stream, const fb2_half *slots,
                                         const float *weights, fb2_half *out, size_t batch,
                                         size_t top_k, size_t hidden);
/* B::fused_silu_mul_split / fused_gelu_tanh_mul_split
 * (crates/ferrum-kernels/src/backend/cpu.rs:1666,1682; kernels/fused_silu_mul.cu:29,66):
 * gate_up [tokens, 2*im] (gate || up) -> out [tokens, im]; f32 math, f16 I/O. */
FB2_API int fb2_fused_silu_mul_split(fb2_stream_t stream, const fb2_half *gate_up,
                                     fb2_half *out, size_t tokens, size_t im);
FB2_API int fb2_fused_gelu_tanh_mul_split(fb2_stream_t stream, const fb2_half *gate_up,
                                          fb2_half *out, size_t tokens, size_t im);

/* ---- adjacent fusions: the ops on either side of the GEMMs, shaped so that no gather launch is
 * left between them and an act-order linear.
 *
 * fb2_fused_add_rms_norm: residual += input (when residual != NULL; it is updated in place, as
 * fused_add_rms_norm_inplace_f16, crates/ferrum-kernels/kernels/fused_add_rms_norm.cu:62-100), then
 * output = rms_norm(residual or input, weight, eps) -- and, with out_perm (the device array
 * fb2_gptq_linear_act_order_perm returns), output[:, j] = normed[:, out_perm[j]]: the activation gather of the
 * act-order linear that follows (reference: kernels/gather_columns.cu:15, a launch of its own) rides on the norm's
 * store; feed the result to fb2_gptq_linear_forward_sorted / _forward_swiglu(x_is_sorted = 1).
 * f16 I/O, f32 arithmetic, hidden <= 65536. */
FB2_API int fb2_fused_add_rms_norm(fb2_stream_t stream, const fb2_half *input, fb2_half *residual,
                                   const fb2_half *weight, fb2_half *output, size_t tokens, size_t hidden,
                                   float eps, const int32_t *out_perm);
/* Device pointer to the layer's act-order column order (sorted position -> input column; K int32), or NULL when the
 * layer reads x as it lies.  Owned by the handle. */
FB2_API int fb2_gptq_linear_act_order_perm(const fb2_linear_t *lin, const int32_t **perm_dev);
/* Linear::forward for activations that already are in that order. */
FB2_API int fb2_gptq_linear_forward_sorted(const fb2_linear_t *lin, fb2_stream_t stream, const fb2_half *x_sorted,
                                           fb2_half *y, size_t m);
/* gate|up linear (out_features = 2 * inter, gate columns first) with SiLU(gate) * up applied in the GEMM epilogue:
 * y [m, out_features / 2].  One launch for the reference's pair Linear::forward + fused_silu_mul_split
 * (crates/ferrum-models/src/models/llama_family.rs:3468-3509); gate and up are rounded to f16 before the activation,
 * as that pair does. */
FB2_API int fb2_gptq_linear_forward_swiglu(const fb2_linear_t *lin, fb2_stream_t stream, const fb2_half *x,
                                           fb2_half *y, size_t m, int x_is_sorted);
/* Row argmax of f16 logits [m, n] -> out_idx [m] (ties -> lowest index; crates/ferrum-kernels/kernels/argmax_rows.cu). */
FB2_API int fb2_argmax_rows_f16(fb2_stream_t stream, const fb2_half *logits, size_t m, size_t n, int32_t *out_idx);

/* The whole MoE FFN of one layer, steps 1-6 of moe_forward_bucketed
 * (crates/ferrum-models/src/moe/dispatch.rs:1558-2192) with gather, SwiGLU and combine
 * fused around the two grouped GEMMs:
 *   out[b, :] = sum_k w[b,k] * down_{e(b,k)}( silu_mul( gate_up_{e(b,k)}( x[b, :] ) ) )
 * router_logits [batch, E] f16; x, out [batch, hidden] f16 (out is overwritten);
 * workspace: fb2_moe_ffn_workspace_bytes() bytes of device memory, caller-owned. */
FB2_API size_t fb2_moe_ffn_workspace_bytes(const fb2_stack_t *gate_up, const fb2_stack_t *down,
                                           size_t max_batch, size_t top_k);
FB2_API int fb2_moe_ffn_forward(const fb2_stack_t *gate_up, const fb2_stack_t *down,
                                fb2_stream_t stream, const fb2_half *x,
                                const fb2_half *router_logits, fb2_half *out, size_t batch,
                                size_t top_k, int norm_topk_prob, void *workspace,
                                size_t workspace_bytes);

/* ------------------------------------------------------------------------------------
 * Tensor parallelism: sharding GPTQ tensors, and the collectives that complete the sharded
 * GEMMs.  Shard maths: crates/ferrum-engine/src/parallel/tensor_parallel.rs:171-195 and
 * crates/ferrum-models/tests/tp_sharding_test.rs:7-170 (the reference shards dense tensors
 * only; quantised sharding -- 8-column qzeros words, 128-row groups, act-order -- is defined
 * here).
 *   FB2_SPLIT_COLUMN  split N.  `segments` (widths of the parts fused along N, e.g. q|k|v or
 *       gate|up; NULL = one part) are split separately and the rank's pieces re-fused, so
 *       SiLU(gate)*up and the head layout stay local.  Every piece must be a multiple of 8
 *       columns.  The shard keeps g_idx (all K rows are local).
 *   FB2_SPLIT_ROW     split K on group boundaries; with act-order the split is taken in
 *       g_idx-SORTED order (every rank owns whole groups and only its own scales / zeros) and
 *       the shard has sequential groups.  x_cols[k_local] = the columns of the full activation
 *       this shard multiplies, in the order it expects them.  The bias stays with rank 0.
 * fb2_gptq_shard_host is pure host code (no GPU needed); outputs are caller-allocated:
 *   qweight_out [k_local/8, n_local], scales_out [k_local/g, n_local],
 *   qzeros_out [k_local/g, n_local/8], bias_out [n_local] (nullable), x_cols_out [k_local]
 *   (nullable).
 * ---------------------------------------------------------------------------------- */
enum { FB2_SPLIT_COLUMN = 0, FB2_SPLIT_ROW = 1 };
FB2_API int fb2_gptq_shard_dims(size_t k, size_t n, size_t group_size, int split, int tp_rank,
                                int tp_size, const size_t *segments, int n_seg, size_t *k_local,
                                size_t *n_local);
FB2_API int fb2_gptq_shard_host(const int32_t *qweight, const float *scales, const int32_t *qzeros,
                                const int32_t *g_idx, const float *bias, size_t group_size,
                                size_t k, size_t n, int split, int tp_rank, int tp_size,
                                const size_t *segments, int n_seg, int32_t *qweight_out,
                                float *scales_out, int32_t *qzeros_out, float *bias_out,
                                int32_t *x_cols_out);
/* load_gptq of this rank's shard: fb2_gptq_shard_host + fb2_gptq_linear_create. */
FB2_API int fb2_gptq_linear_create_sharded(const int32_t *qweight, const float *scales,
                                           const int32_t *qzeros, const int32_t *g_idx,
                                           const float *bias, uint32_t bits, size_t group_size,
                                           size_t k, size_t n, int device, int tp_rank,
                                           int tp_size, int split, const size_t *segments,
                                           int n_seg, int32_t *x_cols_out, fb2_linear_t **out);

/* Collectives.  Replaces BackendCollective::{world_size, rank, all_reduce, all_gather,
 * broadcast} (crates/ferrum-kernels/src/backend/capabilities.rs:85-107) and NcclRank
 * (crates/ferrum-kernels/src/backend/cuda/nccl_comm.rs:12-105): one communicator per process
 * (= per GPU), every call is enqueued on the caller's stream, nothing synchronises.
 * NCCL is bound at run time from the process's libnccl.so.2 (FB2_NCCL_LIB overrides the name);
 * without it the calls return FB2_ERR_NCCL.
 *   fb2_comm_get_unique_id   rank 0 fills 128 bytes (ncclUniqueId) for the caller to distribute
 *   fb2_comm_init_rank       collective over all ranks (ncclCommInitRank)
 *   fb2_all_reduce_f16       in-place sum of `len` halves (nccl_comm.rs:40 all_reduce_f16_inplace)
 *   fb2_all_gather_f16       recv [world, len_per_rank]
 *   fb2_all_to_all_f16       the expert-parallel dispatch / return: peer p receives
 *                            send[send_offsets[p] .. + send_counts[p]) and this rank receives
 *                            recv[recv_offsets[p] .. + recv_counts[p]) from it (elements; host
 *                            arrays of world_size entries, read during the call only)
 *   fb2_gptq_linear_forward_allreduce   row-parallel linear: partial GEMM + all-reduce
 * ---------------------------------------------------------------------------------- */
typedef struct fb2_comm fb2_comm_t;
#define FB2_UNIQUE_ID_BYTES 128
FB2_API int fb2_comm_get_unique_id(void *id_out);
FB2_API int fb2_comm_init_rank(const void *unique_id, int rank, int world_size, int device,
                               fb2_comm_t **out);
FB2_API void fb2_comm_destroy(fb2_comm_t *comm);
FB2_API int fb2_comm_world_size(const fb2_comm_t *comm);
FB2_API int fb2_comm_rank(const fb2_comm_t *comm);
FB2_API int fb2_all_reduce_f16(fb2_comm_t *comm, fb2_stream_t stream, fb2_half *buf, size_t len);
FB2_API int fb2_all_gather_f16(fb2_comm_t *comm, fb2_stream_t stream, const fb2_half *send,
                               fb2_half *recv, size_t len_per_rank);
FB2_API int fb2_broadcast_f16(fb2_comm_t *comm, fb2_stream_t stream, fb2_half *buf, size_t len,
                              int root);
FB2_API int fb2_all_to_all_f16(fb2_comm_t *comm, fb2_stream_t stream, const fb2_half *send,
                               const size_t *send_offsets, const size_t *send_counts,
                               fb2_half *recv, const size_t *recv_offsets,
                               const size_t *recv_counts);
/* Expert-parallel MoE FFN (the reference has no EP -- is_ep is hard-wired false,
 * backend/cuda/quant.rs:1179 -- so parity is defined against fb2_moe_ffn_forward on one GPU).
 * `gate_up` / `down` hold THIS rank's experts [rank*E/R, (rank+1)*E/R); x [batch_local, hidden] and
 * router_logits [batch_local, E] are this rank's tokens (every rank passes the same batch_local).
 * One launch routes the tokens and plans the dispatch, one scatters the rows into a fixed-capacity
 * send buffer [R][batch_local*top_k]; one grouped all-to-all carries rows + expert ids to the expert
 * owners; the owners run the two grouped GEMMs over the received rows; one all-to-all returns the
 * outputs; the token owner combines.  No host sync, fixed message sizes: graph-capturable.
 * Needs batch_local * top_k <= 1024 and E <= 256. */
FB2_API size_t fb2_moe_ep_workspace_bytes(const fb2_stack_t *gate_up, const fb2_stack_t *down,
                                          size_t max_batch_local, size_t top_k, int world_size);
FB2_API int fb2_moe_ep_ffn_forward(const fb2_stack_t *gate_up, const fb2_stack_t *down,
                                   fb2_comm_t *comm, fb2_stream_t stream, const fb2_half *x,
                                   const fb2_half *router_logits, fb2_half *out,
                                   size_t batch_local, size_t top_k, int norm_topk_prob,
                                   void *workspace, size_t workspace_bytes);

/* Collective over all ranks (synchronises `stream`): allocates this rank's receive buffer for
 * one-shot all-reduces of up to `max_elems` halves and maps every peer's buffer (CUDA IPC).
 * Afterwards fb2_gptq_linear_forward_allreduce pushes the GEMM's output tiles to the peers from its
 * epilogue (NVLink stores) and completes the sum with one small reduce kernel instead of an NCCL
 * launch, whenever m * out_features <= max_elems.  All ranks must pass the same max_elems. */
FB2_API int fb2_comm_enable_p2p(fb2_comm_t *comm, fb2_stream_t stream, size_t max_elems);
FB2_API int fb2_gptq_linear_forward_allreduce(const fb2_linear_t *lin, fb2_comm_t *comm,
                                              fb2_stream_t stream, const fb2_half *x,
                                              fb2_half *y, size_t m);

/* ------------------------------------------------------------------------------------
 * Safetensors GPTQ checkpoint ingest (host C++).  Replaces NativeSafetensorsLoader
 * (crates/ferrum-quantization/src/native_safetensors.rs): open + shard index :142-194, typed
 * reads :197-329, read_gptq_raw :342-365, load_stacked_gptq_experts :388-585, the GPTQ branch
 * of load_linear with its fusion shims :597-634, load_gptq_linear :763-880,
 * load_gptq_linear_fused :887-1102, quantize_config.json / config.json :1475-1530.
 * Semantics kept: `model.safetensors` wins over `model.safetensors.index.json`; shard files
 * are the sorted distinct values of weight_map; sym=true 4-bit checkpoints get
 * qzeros = 0x77777777 (:1242-1246); g_idx is validated as validate_gptq_g_idx does
 * (:1288-1325); `<prefix>qkv_proj` / `gate_up_proj` / `in_proj_qkvz` / `in_proj_ba` are
 * synthesised from their parts by concatenation along N when the fused tensor is absent;
 * biases are all-or-none across fused parts.  Errors: missing files / read failures ->
 * FB2_ERR_IO (FerrumError::io); everything malformed -> FB2_ERR_MODEL.
 * The fb2_ckpt_read_* / *_info calls are pure host code (no GPU needed).
 * ---------------------------------------------------------------------------------- */
typedef struct fb2_ckpt fb2_ckpt_t;
enum { FB2_ERR_IO = -7 };
enum { FB2_QUANT_NONE = 0, FB2_QUANT_GPTQ = 1, FB2_QUANT_AWQ = 2, FB2_QUANT_GGUF = 3 };
enum { FB2_DTYPE_OTHER = -1, FB2_DTYPE_F32 = 0, FB2_DTYPE_F16 = 1, FB2_DTYPE_BF16 = 2, FB2_DTYPE_I32 = 3 };
#define FB2_MAX_RANK 8
/* QuantConfig (crates/ferrum-quantization/src/config.rs:26-45); present = 0 when the
 * directory carries no quantisation metadata. */
typedef struct {
  int present;
  int method;
  uint32_t bits;
  size_t group_size;
  int desc_act;
  int sym;
} fb2_quant_config_t;
/* One GPTQ linear as load_gptq will receive it (after fusion): part_n = widths of the fused
 * parts along N (n_parts = 1 for a stored tensor); desc_act = g_idx is a non-trivial order. */
typedef struct {
  size_t k, n;
  int has_g_idx, has_bias, desc_act, n_parts;
  size_t part_n[4];
} fb2_gptq_info_t;

FB2_API int fb2_ckpt_open(const char *model_dir, fb2_ckpt_t **out);
FB2_API void fb2_ckpt_close(fb2_ckpt_t *ckpt);
FB2_API int fb2_ckpt_quant_config(const fb2_ckpt_t *ckpt, fb2_quant_config_t *out);
FB2_API size_t fb2_ckpt_num_tensors(const fb2_ckpt_t *ckpt);
FB2_API const char *fb2_ckpt_tensor_name(const fb2_ckpt_t *ckpt, size_t i);
FB2_API int fb2_ckpt_has_tensor(const fb2_ckpt_t *ckpt, const char *name);
/* shape must have room for FB2_MAX_RANK entries; any out pointer may be NULL. */
FB2_API int fb2_ckpt_tensor_info(const fb2_ckpt_t *ckpt, const char *name, int *dtype, size_t *rank,
                                 size_t *shape, size_t *nbytes);
/* F32 / F16 / BF16 -> f32 (read_f32, :197); I32 only (read_i32, :305).  capacity in elements. */
FB2_API int fb2_ckpt_read_f32(const fb2_ckpt_t *ckpt, const char *name, float *out, size_t capacity);
FB2_API int fb2_ckpt_read_i32(const fb2_ckpt_t *ckpt, const char *name, int32_t *out, size_t capacity);
/* Resolve `name` (stored or fused) and report / copy out the host tensors exactly as they are
 * handed to load_gptq: qweight [k/8, n], scales [k/g, n] f32, qzeros [k/g, n/8], g_idx [k]
 * (when has_g_idx), bias [n] (when has_bias).  Output pointers may be NULL to skip a tensor. */
FB2_API int fb2_ckpt_gptq_info(const fb2_ckpt_t *ckpt, const char *name, fb2_gptq_info_t *info);
FB2_API int fb2_ckpt_read_gptq(const fb2_ckpt_t *ckpt, const char *name, int32_t *qweight, float *scales,
                               int32_t *qzeros, int32_t *g_idx, float *bias);
/* WeightLoader::load_linear for a GPTQ layer: read + fb2_gptq_linear_create. */
FB2_API int fb2_ckpt_load_gptq_linear(const fb2_ckpt_t *ckpt, const char *name, int device,
                                      fb2_linear_t **out);
/* The same for this rank's tensor-parallel shard; for a column split the fused parts are the
 * segments (q|k|v and gate|up are split separately, tp_sharding_test.rs:7-52,114-134). */
FB2_API int fb2_ckpt_load_gptq_linear_sharded(const fb2_ckpt_t *ckpt, const char *name, int device,
                                              int tp_rank, int tp_size, int split,
                                              int32_t *x_cols_out, fb2_linear_t **out);
/* load_stacked_gptq_experts (:388): expert e's tensors are `<fmt with {e} replaced><proj>` for
 * each proj in order, concatenated along N (Qwen3-MoE: fmt "model.layers.L.mlp.experts.{e}.",
 * projs {"gate_proj","up_proj"} and {"down_proj"}, qwen3_moe/load.rs:237-256).  Experts
 * [expert_begin, expert_begin + num_experts) are loaded (an expert-parallel rank passes its
 * own range).  The *_info / read_* forms are host-only views of one expert. */
FB2_API int fb2_ckpt_stacked_gptq_info(const fb2_ckpt_t *ckpt, const char *expert_prefix_fmt,
                                       const char *const *proj_names, size_t n_proj, size_t expert,
                                       size_t *k, size_t *n_per_expert, int *has_g_idx);
FB2_API int fb2_ckpt_read_stacked_gptq_expert(const fb2_ckpt_t *ckpt, const char *expert_prefix_fmt,
                                              const char *const *proj_names, size_t n_proj,
                                              size_t expert, int32_t *qweight, float *scales,
                                              int32_t *qzeros, int32_t *g_idx);
FB2_API int fb2_ckpt_load_stacked_gptq_experts(const fb2_ckpt_t *ckpt, const char *expert_prefix_fmt,
                                               size_t expert_begin, size_t num_experts,
                                               const char *const *proj_names, size_t n_proj,
                                               int device, fb2_stack_t **out, size_t *n_per_expert,
                                               size_t *k);

#ifdef __cplusplus
}
#endif
#endif /* FB2_W4A16_H */
