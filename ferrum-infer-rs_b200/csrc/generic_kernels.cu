// generic_kernels.cu -- shape-agnostic CUDA kernels of the W4A16 path.
//
//  * generic_gemm: W4A16 GEMM straight from the on-disk GPTQ layout (any group_size,
//    N % 8 == 0, act-order through per-row g_idx lookups).  It serves the shapes the
//    tensor-core kernels (tc_gemm.cu) do not take (N or K not a multiple of 128,
//    group_size != 128) and is the in-library cross-check for them.  Weights are
//    dequantised exactly as the CPU reference does (integer subtract, one f32 multiply:
//    crates/ferrum-kernels/src/backend/cpu.rs:2297-2313), accumulated in fp32.
//  * dequant kernels: stored weights -> the reference's [N, K] matrix, bit-exact.
//  * permute_cols: the act-order activation gather A'[m, j] = A[m, perm[j]]
//    (reference: crates/ferrum-kernels/kernels/gather_columns.cu:15).
//  * silu/gelu * up: crates/ferrum-kernels/kernels/fused_silu_mul.cu:29,66.
#include "common.cuh"
#include "rowmap.cuh"

namespace fb2 {

namespace {

constexpr int kGenCols = 128;   // columns (threads) per CTA
constexpr int kGenRows = 8;     // activation rows per pass
constexpr int kGenKC = 512;     // k per shared-memory chunk

__device__ __forceinline__ float silu_mul(float g, float u) { return (g / (1.0f + expf(-g))) * u; }
__device__ __forceinline__ float gelu_tanh_mul(float g, float u) {
  const float c = 0.79788456f;
  float inner = c * (g + 0.044715f * g * g * g);
  return 0.5f * g * (1.0f + tanhf(inner)) * u;
}

struct GenCol {
  const int32_t *qw;       // + column
  const __half *sc;
  const int32_t *qz;
  int zshift;
  int ld;
};

__device__ __forceinline__ GenCol gen_col(const WeightRef &w, int expert, int col) {
  GenCol c;
  int sc_col = w.col0 + col;
  size_t groups = (size_t)w.k / w.group_size;
  c.qw = w.qweight + (size_t)expert * (w.k / 8) * w.ld + sc_col;
  c.sc = w.scales + (size_t)expert * groups * w.ld + sc_col;
  c.qz = w.qzeros + (size_t)expert * groups * (w.ld / 8) + sc_col / 8;
  c.zshift = (sc_col % 8) * 4;
  c.ld = w.ld;
  return c;
}

// kSilu: thread owns gate column c and up column c + n/2; output width n/2.
template <bool kSilu>
__device__ __forceinline__ void generic_gemm_body(const WeightRef &w, const RowMap &rm,
                                                  const __half *__restrict__ x,
                                                  __half *__restrict__ y,
                                                  const DispatchTable *tbl) {
  __shared__ __half xs[kGenRows][kGenKC];
  __shared__ int a_rows[kGenRows], c_rows[kGenRows];

  const int nb = (rm.mode == ROWS_ROUTED) ? rm.block : kGenRows;
  const int b = blockIdx.y;
  if (b >= rowmap_num_blocks(rm, nb)) return;  // uniform per CTA
  const int expert = rowmap_block_expert(rm, tbl, b);
  const int n_out = kSilu ? w.n / 2 : w.n;
  const int col = blockIdx.x * kGenCols + threadIdx.x;
  const bool col_ok = col < n_out;
  GenCol c0 = gen_col(w, expert, col_ok ? col : 0);
  GenCol c1 = gen_col(w, expert, col_ok ? (kSilu ? col + n_out : col) : 0);

  for (int r0 = 0; r0 < nb; r0 += kGenRows) {
    __syncthreads();
    if (threadIdx.x < kGenRows) {
      int ar = 0, cr = 0;
      bool ok = (r0 + threadIdx.x < nb) && rowmap_resolve(rm, tbl, nb, b, r0 + threadIdx.x, ar, cr);
      a_rows[threadIdx.x] = ok ? ar : -1;
      c_rows[threadIdx.x] = ok ? cr : -1;
    }
    __syncthreads();
    bool any = false;
    for (int r = 0; r < kGenRows; r++) any |= (a_rows[r] >= 0);
    if (!any) continue;

    float acc0[kGenRows], acc1[kGenRows];
#pragma unroll
    for (int r = 0; r < kGenRows; r++) acc0[r] = acc1[r] = 0.0f;

    for (int k0 = 0; k0 < w.k; k0 += kGenKC) {
      const int kc = min(kGenKC, w.k - k0);
      __syncthreads();
      for (int idx = threadIdx.x; idx < kGenRows * kc; idx += kGenCols) {
        int r = idx / kc, kk = idx - r * kc;
        xs[r][kk] = a_rows[r] >= 0 ? x[(size_t)a_rows[r] * w.k + k0 + kk] : __float2half(0.0f);
      }
      __syncthreads();
      if (col_ok) {
        int last_g = -1;
        float s0 = 0.f, s1 = 0.f;
        int z0 = 0, z1 = 0;
        for (int pr = 0; pr < kc / 8; pr++) {
          const int kbase = k0 + pr * 8;
          uint32_t p0 = (uint32_t)c0.qw[(size_t)(kbase / 8) * c0.ld];
          uint32_t p1 = kSilu ? (uint32_t)c1.qw[(size_t)(kbase / 8) * c1.ld] : 0u;
#pragma unroll
          for (int bi = 0; bi < 8; bi++) {
            const int k = kbase + bi;
            const int g = w.g_idx ? w.g_idx[k] : k / w.group_size;
            if (g != last_g) {
              last_g = g;
              s0 = __half2float(c0.sc[(size_t)g * c0.ld]);
              z0 = (int)(((uint32_t)c0.qz[(size_t)g * (c0.ld / 8)] >> c0.zshift) & 0xF) + 1;
              if (kSilu) {
                s1 = __half2float(c1.sc[(size_t)g * c1.ld]);
                z1 = (int)(((uint32_t)c1.qz[(size_t)g * (c1.ld / 8)] >> c1.zshift) & 0xF) + 1;
              }
            }
            const float w0 = (float)((int)((p0 >> (bi * 4)) & 0xF) - z0) * s0;
            const float w1 = kSilu ? (float)((int)((p1 >> (bi * 4)) & 0xF) - z1) * s1 : 0.f;
#pragma unroll
            for (int r = 0; r < kGenRows; r++) {
              const float xv = __half2float(xs[r][pr * 8 + bi]);
              acc0[r] = fmaf(w0, xv, acc0[r]);
              if (kSilu) acc1[r] = fmaf(w1, xv, acc1[r]);
            }
          }
        }
      }
    }
    if (col_ok) {
#pragma unroll
      for (int r = 0; r < kGenRows; r++) {
        if (c_rows[r] < 0) continue;
        float v;
        if (kSilu) {
          // the reference rounds gate_up to f16 before SiLU (fused_silu_mul.cu reads f16)
          float g = __half2float(__float2half_rn(acc0[r]));
          float u = __half2float(__float2half_rn(acc1[r]));
          v = silu_mul(g, u);
        } else {
          v = acc0[r] + (w.bias ? w.bias[col] : 0.0f);
        }
        y[(size_t)c_rows[r] * n_out + col] = __float2half_rn(v);
      }
    }
  }
}

template <bool kSilu>
__global__ void __launch_bounds__(kGenCols)
generic_gemm_kernel(const WeightRef w, const RowMap rm, const __half *__restrict__ x,
                    __half *__restrict__ y) {
  generic_gemm_body<kSilu>(w, rm, x, y, nullptr);
}

__global__ void __launch_bounds__(kGenCols)
generic_gemm_dispatch_kernel(const WeightRef w, const RowMap rm, const __half *__restrict__ x,
                             __half *__restrict__ y, const __grid_constant__ DispatchTable tbl) {
  generic_gemm_body<false>(w, rm, x, y, &tbl);
}

// --------------------------------------------------------------------------- dequant
template <typename OutT>
__device__ __forceinline__ void store_w(OutT *p, float v);
template <>
__device__ __forceinline__ void store_w<float>(float *p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_w<__half>(__half *p, float v) { *p = __float2half_rn(v); }

template <typename OutT>
__global__ void dequant_raw_kernel(const WeightRef w, int expert, OutT *__restrict__ out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int pr = blockIdx.y;
  if (col >= w.n) return;
  GenCol c = gen_col(w, expert, col);
  uint32_t packed = (uint32_t)c.qw[(size_t)pr * c.ld];
#pragma unroll
  for (int bi = 0; bi < 8; bi++) {
    const int k = pr * 8 + bi;
    const int g = w.g_idx ? w.g_idx[k] : k / w.group_size;
    const float s = __half2float(c.sc[(size_t)g * c.ld]);
    const int z = (int)(((uint32_t)c.qz[(size_t)g * (c.ld / 8)] >> c.zshift) & 0xF) + 1;
    store_w<OutT>(out + (size_t)col * w.k + k, (float)((int)((packed >> (bi * 4)) & 0xF) - z) * s);
  }
}

// One thread per packed word of the tiled format.
template <typename OutT>
__global__ void dequant_tiled_kernel(const WeightRef w, int expert, OutT *__restrict__ out) {
  const int groups = w.k / kGroup;
  const int tiles_per_expert = w.ld / kTileN;
  const size_t words_per_stage = kStageWBytes / 4;  // 2048
  const size_t total = (size_t)(w.n / kTileN) * groups * words_per_stage;
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int word = (int)(idx % words_per_stage);
  const int g = (int)((idx / words_per_stage) % groups);
  const int t = (int)(idx / words_per_stage / groups);
  const int j = word & 3, nl = (word >> 2) & 127, c = word >> 9;
  const uint8_t *stage = w.tiles +
      ((size_t)expert * tiles_per_expert + (w.col0 / kTileN) + t) * groups * kStageBytes +
      (size_t)g * kStageBytes;
  const uint32_t packed = reinterpret_cast<const uint32_t *>(stage)[word];
  const uint32_t sz = reinterpret_cast<const uint32_t *>(stage + kStageWBytes)[nl];
  const float s = __half2float(__ushort_as_half((unsigned short)(sz & 0xFFFF)));
  const int z = (int)__half2float(__ushort_as_half((unsigned short)(sz >> 16))) - 1024;
  const int col = t * kTileN + nl;
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int sk = g * kGroup + tile_nibble_k(c, j, p);
    const int dk = w.perm ? w.perm[sk] : sk;
    store_w<OutT>(out + (size_t)col * w.k + dk, (float)((int)((packed >> (p * 4)) & 0xF) - z) * s);
  }
}

__global__ void permute_cols_kernel(const __half *__restrict__ x, const int32_t *__restrict__ perm,
                                    __half *__restrict__ xp, int m, int k) {
  // PDL: let the GEMM that follows start streaming its weights now; wait for the previous GEMM
  // (which may still be reading the shared xp scratch) before writing.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int src = j < k ? perm[j] : 0;            // a weight-side constant: in flight before the grid dependency resolves
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (j >= k) return;
  for (int r = blockIdx.y; r < m; r += gridDim.y) xp[(size_t)r * k + j] = x[(size_t)r * k + src];
}

// Batched (prefill) form: one CTA per activation row.  The row is staged in shared memory with
// 16-byte loads and written back permuted with 16-byte stores, so global traffic is exactly one
// read and one write of x (the element-wise kernel above moves 2 bytes per 32-byte sector).
__global__ void __launch_bounds__(256) permute_rows_kernel(const __half *__restrict__ x,
                                                           const int32_t *__restrict__ perm,
                                                           __half *__restrict__ xp, int m, int k) {
  extern __shared__ __align__(16) __half s_row[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int k8 = k >> 3;                                         // k % 8 == 0 (checked by the launcher)
  for (int r = blockIdx.x; r < m; r += gridDim.x) {
    const uint4 *src = reinterpret_cast<const uint4 *>(x + (size_t)r * k);
    for (int i = threadIdx.x; i < k8; i += blockDim.x) reinterpret_cast<uint4 *>(s_row)[i] = src[i];
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(xp + (size_t)r * k);
    for (int i = threadIdx.x; i < k8; i += blockDim.x) {
      const int4 p0 = reinterpret_cast<const int4 *>(perm)[2 * i];
      const int4 p1 = reinterpret_cast<const int4 *>(perm)[2 * i + 1];
      const unsigned short *sr = reinterpret_cast<const unsigned short *>(s_row);
      uint4 o;
      o.x = (uint32_t)sr[p0.x] | ((uint32_t)sr[p0.y] << 16);
      o.y = (uint32_t)sr[p0.z] | ((uint32_t)sr[p0.w] << 16);
      o.z = (uint32_t)sr[p1.x] | ((uint32_t)sr[p1.y] << 16);
      o.w = (uint32_t)sr[p1.z] | ((uint32_t)sr[p1.w] << 16);
      dst[i] = o;
    }
    __syncthreads();
  }
}

template <bool kGelu>
__global__ void act_mul_split_kernel(const __half *__restrict__ gu, __half *__restrict__ out,
                                     size_t tokens, size_t im) {
  size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= tokens * im) return;
  size_t t = idx / im, i = idx - t * im;
  float g = __half2float(gu[t * 2 * im + i]);
  float u = __half2float(gu[t * 2 * im + im + i]);
  out[idx] = __float2half_rn(kGelu ? gelu_tanh_mul(g, u) : silu_mul(g, u));
}

// Residual add + RMSNorm (reference: kernels/fused_add_rms_norm.cu:21-60, rms_norm.cu), one CTA per token, with the
// act-order column gather of the linear that consumes the result folded into the store: out[:, j] = normed[:, perm[j]].
// (The reference runs gather_columns.cu as a separate launch in front of every act-order Marlin call.)
// f16 in/out, f32 arithmetic; the sum x + r is rounded to f16 (it IS the new residual) before it is normalised, as
// the reference does.
__global__ void __launch_bounds__(1024) add_rms_norm_kernel(const __half *__restrict__ input, __half *__restrict__ residual,
                                                            const __half *__restrict__ weight, __half *__restrict__ output,
                                                            const int32_t *__restrict__ out_perm, int hidden, float eps) {
  extern __shared__ __align__(16) __half s_row[];       // the (updated) residual row
  __shared__ float s_part[32];
  __shared__ float s_inv;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const size_t off = (size_t)blockIdx.x * hidden;
  float var = 0.f;
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) {
    float v = __half2float(input[off + i]);
    if (residual) v += __half2float(residual[off + i]);
    const __half hv = __float2half(v);
    if (residual) residual[off + i] = hv;
    s_row[i] = hv;
    const float r = __half2float(hv);
    var += r * r;
  }
  for (int o = 16; o; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = var;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x + 31) / 32 ? s_part[threadIdx.x] : 0.f;
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) s_inv = rsqrtf(v / (float)hidden + eps);
  }
  __syncthreads();
  const float inv = s_inv;
  for (int j = threadIdx.x; j < hidden; j += blockDim.x) {
    const int i = out_perm ? out_perm[j] : j;
    output[off + j] = __float2half(__half2float(s_row[i]) * inv * __half2float(weight[i]));
  }
}

// Row argmax of f16 logits (reference: kernels/argmax_rows.cu: one CTA per row, ties -> lowest index).
__global__ void __launch_bounds__(256) argmax_rows_kernel(const __half *__restrict__ logits, int n, int32_t *__restrict__ out) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const __half *row = logits + (size_t)blockIdx.x * n;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = __half2float(row[i]);
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = best; s_i[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; w++)
      if (s_v[w] > best || (s_v[w] == best && s_i[w] < bi)) { best = s_v[w]; bi = s_i[w]; }
    out[blockIdx.x] = bi == 0x7fffffff ? 0 : bi;
  }
}

}  // namespace

int launch_add_rms_norm(const __half *input, __half *residual, const __half *weight, __half *output,
                        const int32_t *out_perm, int tokens, int hidden, float eps, cudaStream_t stream) {
  if (tokens <= 0) return FB2_OK;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tokens);
  cfg.blockDim = dim3(hidden >= 1024 ? 1024 : ((hidden + 31) / 32) * 32);
  cfg.dynamicSmemBytes = (size_t)hidden * sizeof(__half);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cfg.dynamicSmemBytes > 48 * 1024)
    FB2_CUDA_TRY(cudaFuncSetAttribute(add_rms_norm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes));
  FB2_CUDA_TRY(cudaLaunchKernelEx(&cfg, add_rms_norm_kernel, input, residual, weight, output, out_perm, hidden, eps));
  count_launch();
  return FB2_OK;
}

int launch_argmax_rows(const __half *logits, int m, int n, int32_t *out, cudaStream_t stream) {
  if (m <= 0) return FB2_OK;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(m);
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FB2_CUDA_TRY(cudaLaunchKernelEx(&cfg, argmax_rows_kernel, logits, n, out));
  count_launch();
  return FB2_OK;
}

int launch_generic_gemm(const WeightRef &w, const RowMap &rows, const __half *x, __half *y,
                        int epilogue, cudaStream_t stream) {
  if (!w.qweight) {
    set_error("generic GEMM needs the raw GPTQ tensors on the device (layer was loaded tile-only)");
    return FB2_ERR_UNSUPPORTED;
  }
  // The routed form is launched with the worst-case block count; kernels exit early.
  int nblocks;
  if (rows.mode == ROWS_DENSE) nblocks = (rows.m + kGenRows - 1) / kGenRows;
  else nblocks = rows.max_blocks;
  if (nblocks <= 0) return FB2_OK;
  const int n_out = epilogue ? w.n / 2 : w.n;
  dim3 grid((n_out + kGenCols - 1) / kGenCols, nblocks);
  if (epilogue) generic_gemm_kernel<true><<<grid, kGenCols, 0, stream>>>(w, rows, x, y);
  else generic_gemm_kernel<false><<<grid, kGenCols, 0, stream>>>(w, rows, x, y);
  FB2_LAUNCH_CHECK();
  return FB2_OK;
}

int launch_generic_gemm_dispatch(const WeightRef &w, const RowMap &rows, const DispatchTable &tbl,
                                 const __half *x, __half *y, cudaStream_t stream) {
  if (!w.qweight) {
    set_error("generic GEMM needs the raw GPTQ tensors on the device");
    return FB2_ERR_UNSUPPORTED;
  }
  if (rows.num_dispatch <= 0) return FB2_OK;
  dim3 grid((w.n + kGenCols - 1) / kGenCols, rows.num_dispatch);
  generic_gemm_dispatch_kernel<<<grid, kGenCols, 0, stream>>>(w, rows, x, y, tbl);
  FB2_LAUNCH_CHECK();
  return FB2_OK;
}

int launch_dequant(const WeightRef &w, int expert, float *out_f32, __half *out_f16,
                   cudaStream_t stream) {
  if (w.tiles) {
    const size_t total = (size_t)(w.n / kTileN) * (w.k / kGroup) * (kStageWBytes / 4);
    const int threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    if (out_f32) dequant_tiled_kernel<float><<<blocks, threads, 0, stream>>>(w, expert, out_f32);
    else dequant_tiled_kernel<__half><<<blocks, threads, 0, stream>>>(w, expert, out_f16);
  } else {
    dim3 grid((w.n + 127) / 128, w.k / 8);
    if (out_f32) dequant_raw_kernel<float><<<grid, 128, 0, stream>>>(w, expert, out_f32);
    else dequant_raw_kernel<__half><<<grid, 128, 0, stream>>>(w, expert, out_f16);
  }
  FB2_LAUNCH_CHECK();
  return FB2_OK;
}

int launch_permute_cols(const __half *x, const int32_t *perm, __half *xp, int m, int k,
                        cudaStream_t stream) {
  if (m <= 0) return FB2_OK;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const size_t row_bytes = (size_t)k * sizeof(__half);
  const bool aligned = k % 8 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)xp % 16) == 0 && ((uintptr_t)perm % 16) == 0;
  static const int rows_min = [] { const char *e = getenv("FB2_DEBUG_PERMUTE_ROWS_MIN"); return e ? atoi(e) : 64; }();
  if (m >= rows_min && aligned && row_bytes <= 48 * 1024) {
    // batched rows: one CTA per row through shared memory
    cfg.gridDim = dim3(m < 1184 ? m : 1184);
    cfg.dynamicSmemBytes = row_bytes;
    FB2_CUDA_TRY(cudaLaunchKernelEx(&cfg, permute_rows_kernel, x, perm, xp, m, k));
    count_launch();
    return FB2_OK;
  }
  cfg.gridDim = dim3((k + 255) / 256, m < 64 ? m : 64);
  FB2_CUDA_TRY(cudaLaunchKernelEx(&cfg, permute_cols_kernel, x, perm, xp, m, k));
  count_launch();
  return FB2_OK;
}

int launch_silu_mul_split(const __half *gu, __half *out, size_t tokens, size_t im, int gelu,
                          cudaStream_t stream) {
  const size_t total = tokens * im;
  if (total == 0) return FB2_OK;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (gelu) act_mul_split_kernel<true><<<blocks, 256, 0, stream>>>(gu, out, tokens, im);
  else act_mul_split_kernel<false><<<blocks, 256, 0, stream>>>(gu, out, tokens, im);
  FB2_LAUNCH_CHECK();
  return FB2_OK;
}

}  // namespace fb2
