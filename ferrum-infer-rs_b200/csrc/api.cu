// api.cu -- the extern "C" surface of libfb2_w4a16.so (include/fb2_w4a16.h): handle
// lifetime, GPTQ ingest + device-side repack into the tile-stripe format, per-device
// scratch, and the forward / grouped-GEMM / MoE entry points that pick a kernel family.
//
// Reference behaviour followed (paths relative to the reference checkout):
//   load_gptq            crates/ferrum-kernels/src/backend/cuda/quant.rs:612-766
//   load_gptq_stacked    crates/ferrum-kernels/src/backend/cuda/quant.rs:768-990
//   Linear::forward      crates/ferrum-kernels/src/quant_linear/cuda_marlin.rs:83-181
//   MarlinExpertStack    crates/ferrum-kernels/src/marlin_expert_stack.rs:31-109
//   moe_forward_bucketed crates/ferrum-models/src/moe/dispatch.rs:1558-2192
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace fb2 {

// ------------------------------------------------------------------------------ errors
static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static const bool on = [] {
    const char *e = getenv("FB2_DEBUG_NO_PDL");
    return !(e && e[0] == '1');
  }();
  return on;
}

// ---------------------------------------------------------------------- device scratch
// Split-K partial sums and tile arrival counters live in one per-device pool (the
// reference keeps its Marlin gather scratch the same way: process-global, keyed by device
// ordinal, pre-grown outside graph capture -- quant.rs:228-322).
constexpr int kMaxDevices = 64;
struct DeviceScratch {
  std::mutex mu;
  bool probed = false;
  int sm_count = 0;
  float *partials = nullptr;
  size_t partial_bytes = 0;
  int32_t *counters = nullptr;
  size_t num_counters = 0;
  size_t want_partial_bytes = 0;   // high-water mark requested by create()/pregrow
  size_t want_counters = 0;
  __half *xperm = nullptr;         // act-order: activations with columns gathered into sorted-k order
  size_t xperm_bytes = 0;
  uint32_t *ll = nullptr;          // flag-in-value partial slots (fixed size, all-ones = empty)
  // Outgrown buffers are RETIRED, never freed: a captured CUDA graph (or a thread that has already read the pool)
  // may still hold their addresses in kernel parameters.  The pool only grows to the high-water mark of the
  // process, so what is kept is bounded by the sum of a geometric-ish series of earlier sizes.
  std::vector<void *> retired;
};
static DeviceScratch g_scratch[kMaxDevices];

static int probe_device(int device) {
  if (device < 0 || device >= kMaxDevices) {
    set_error("device ordinal %d out of range", device);
    return FB2_ERR_ARG;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    set_error("no CUDA device usable (%s): this library has no CPU fallback",
              e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return FB2_ERR_CUDA;
  }
  if (device >= count) {
    set_error("device %d requested but only %d visible", device, count);
    return FB2_ERR_CUDA;
  }
  DeviceScratch &s = g_scratch[device];
  std::lock_guard<std::mutex> lock(s.mu);
  if (!s.probed) {
    cudaDeviceProp prop;
    FB2_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
      set_error("device %d is sm_%d%d; libfb2_w4a16 is built for sm_100a only", device, prop.major,
                prop.minor);
      return FB2_ERR_CUDA;
    }
    s.sm_count = prop.multiProcessorCount;
    // FB2_DEBUG_NO_LL=1 keeps every split tile on the counter exchange (a measurement aid)
    const char *no_ll = getenv("FB2_DEBUG_NO_LL");
    if (!(no_ll && no_ll[0] == '1')) {
      FB2_CUDA_TRY(cudaSetDevice(device));
      FB2_CUDA_TRY(cudaMalloc((void **)&s.ll, kLlBytes));
      FB2_CUDA_TRY(cudaMemset(s.ll, 0xff, kLlBytes));
      FB2_CUDA_TRY(cudaDeviceSynchronize());
    }
    s.probed = true;
  }
  FB2_CUDA_TRY(cudaSetDevice(device));
  return FB2_OK;
}

// Grow the pool (never inside graph capture: cudaMalloc/cudaFree).
static int grow_scratch(int device, size_t partial_bytes, size_t counters) {
  DeviceScratch &s = g_scratch[device];
  std::lock_guard<std::mutex> lock(s.mu);
  s.want_partial_bytes = std::max(s.want_partial_bytes, partial_bytes);
  s.want_counters = std::max(s.want_counters, counters);
  if (s.want_partial_bytes > s.partial_bytes) {
    if (s.partials) s.retired.push_back(s.partials);
    s.partials = nullptr;
    s.partial_bytes = 0;
    FB2_CUDA_TRY(cudaMalloc(&s.partials, s.want_partial_bytes));
    FB2_CUDA_TRY(cudaMemset(s.partials, 0, s.want_partial_bytes));
    s.partial_bytes = s.want_partial_bytes;
  }
  if (s.want_counters > s.num_counters) {
    if (s.counters) s.retired.push_back(s.counters);
    s.counters = nullptr;
    s.num_counters = 0;
    FB2_CUDA_TRY(cudaMalloc(&s.counters, s.want_counters * sizeof(int32_t)));
    FB2_CUDA_TRY(cudaMemset(s.counters, 0, s.want_counters * sizeof(int32_t)));
    s.num_counters = s.want_counters;
  }
  return FB2_OK;
}

static int grow_xperm(int device, size_t bytes) {
  DeviceScratch &s = g_scratch[device];
  std::lock_guard<std::mutex> lock(s.mu);
  if (bytes <= s.xperm_bytes) return FB2_OK;
  if (s.xperm) s.retired.push_back(s.xperm);
  s.xperm = nullptr;
  s.xperm_bytes = 0;
  FB2_CUDA_TRY(cudaMalloc((void **)&s.xperm, bytes));
  s.xperm_bytes = bytes;
  return FB2_OK;
}

static TcScratch current_scratch(int device, int *sm_count) {
  DeviceScratch &s = g_scratch[device];
  std::lock_guard<std::mutex> lock(s.mu);
  if (sm_count) *sm_count = s.sm_count;
  return TcScratch{s.partials, s.partial_bytes, s.counters, s.num_counters, s.ll, s.ll ? kLlMaxCtas : 0};
}

// -------------------------------------------------------------------- repack kernels
// Raw GPTQ tensors (device copies of what load_gptq receives) -> tile-stripe format.
// One thread per output word: 8 nibble gathers (tile_nibble_k gives the sorted k of each nibble).
__global__ void repack_tiles_kernel(const int32_t *__restrict__ qweight,   // [K/8][n]
                                    const float *__restrict__ scales,      // [G][n]
                                    const int32_t *__restrict__ qzeros,    // [G][n/8]
                                    const int32_t *__restrict__ perm,      // [K] or null
                                    uint8_t *__restrict__ tiles, int k, int n) {
  const int groups = k / kGroup;
  const int per_stage = kStageBytes / 4;                 // 2176 words: 2048 packed + 128 sz
  const size_t total = (size_t)(n / kTileN) * groups * per_stage;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int word = (int)(idx % per_stage);
  const int g = (int)((idx / per_stage) % groups);
  const int t = (int)(idx / per_stage / groups);
  uint32_t *stage = reinterpret_cast<uint32_t *>(tiles + ((size_t)t * groups + g) * kStageBytes);
  if (word >= kStageWBytes / 4) {
    const int nl = word - kStageWBytes / 4;
    const int col = t * kTileN + nl;
    const __half s = __float2half_rn(scales[(size_t)g * n + col]);   // quant.rs:717 re-rounds
    const uint32_t zn = ((uint32_t)qzeros[(size_t)g * (n / 8) + col / 8] >> ((col % 8) * 4)) & 0xFu;
    const __half z = __float2half_rn(1024.0f + (float)(zn + 1));     // exact
    stage[word] = (uint32_t)__half_as_ushort(s) | ((uint32_t)__half_as_ushort(z) << 16);
    return;
  }
  const int j = word & 3, nl = (word >> 2) & 127, c = word >> 9;
  const int col = t * kTileN + nl;
  uint32_t out = 0;
#pragma unroll
  for (int p = 0; p < 8; p++) {
    const int sk = g * kGroup + tile_nibble_k(c, j, p);          // sorted k of this nibble
    const int dk = perm ? perm[sk] : sk;                          // disk row
    const uint32_t src = (uint32_t)qweight[(size_t)(dk >> 3) * n + col];
    out |= ((src >> ((dk & 7) * 4)) & 0xFu) << (p * 4);
  }
  stage[word] = out;
}

__global__ void f32_to_f16_kernel(const float *__restrict__ in, __half *__restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}

// ------------------------------------------------------------------------- validation
struct Ingest {
  bool act_order = false;
  bool tc = false;                  // tile-stripe format applies
  std::vector<int32_t> perm;        // sorted position -> disk row (act-order only)
};

static int validate_common(uint32_t bits, size_t group_size, size_t k, size_t n, const char *what) {
  if (bits != 4) {
    set_error("%s: only bits=4 supported (got %u)", what, bits);
    return FB2_ERR_UNSUPPORTED;
  }
  if (group_size == 0 || k == 0 || n == 0 || k % 8 != 0 || n % 8 != 0 || k % group_size != 0) {
    set_error("%s: shape K=%zu N=%zu group_size=%zu violates GPTQ packing (K%%8, N%%8, "
              "K%%group_size must be 0)", what, k, n, group_size);
    return FB2_ERR_SHAPE;
  }
  if (k > (size_t)INT32_MAX || n > (size_t)INT32_MAX) {
    set_error("%s: shape too large", what);
    return FB2_ERR_SHAPE;
  }
  return FB2_OK;
}

// g_idx handling of load_gptq (quant.rs:685-711) + validate_cuda_marlin_desc_act_g_idx
// (native_safetensors.rs:1327-1372).
static int analyse_g_idx(const int32_t *g_idx, size_t group_size, size_t k, size_t n, Ingest *ing) {
  const size_t groups = k / group_size;
  ing->act_order = false;
  if (g_idx) {
    for (size_t i = 0; i < k; i++) {
      if (g_idx[i] < 0 || (size_t)g_idx[i] >= groups) {
        set_error("g_idx[%zu] = %d outside [0, %zu)", i, g_idx[i], groups);
        return FB2_ERR_MODEL;
      }
      if ((size_t)g_idx[i] != i / group_size) ing->act_order = true;
    }
  }
  bool balanced = true;
  if (ing->act_order) {
    std::vector<size_t> counts(groups, 0);
    for (size_t i = 0; i < k; i++) counts[g_idx[i]]++;
    for (size_t g = 0; g < groups; g++) balanced &= (counts[g] == group_size);
  }
  ing->tc = group_size == (size_t)kGroup && k % kGroup == 0 && n % kTileN == 0 && balanced;
  if (ing->act_order && ing->tc) {
    ing->perm.resize(k);
    std::iota(ing->perm.begin(), ing->perm.end(), 0);
    std::stable_sort(ing->perm.begin(), ing->perm.end(),
                     [&](int32_t a, int32_t b) { return g_idx[a] < g_idx[b]; });
  }
  return FB2_OK;
}

template <typename T>
static int upload(T **dst, const T *src, size_t count) {
  *dst = nullptr;
  FB2_CUDA_TRY(cudaMalloc((void **)dst, count * sizeof(T)));
  FB2_CUDA_TRY(cudaMemcpy(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice));
  return FB2_OK;
}

}  // namespace fb2

using namespace fb2;

// ------------------------------------------------------------------------------ handles
struct fb2_stack {
  WeightRef w{};              // num_experts = E, n = ld = n_per_expert
  int device = 0;
  bool act_order = false;
  std::vector<void *> allocs;
};

struct fb2_linear {
  WeightRef w{};
  int device = 0;
  bool act_order = false;
  int expert = 0;             // view: expert index inside the parent stack
  std::vector<void *> allocs; // owned device memory (empty for views, except bias)
  // forward_host staging
  std::mutex host_mu;
  void *h_x = nullptr, *h_y = nullptr, *d_x = nullptr, *d_y = nullptr;
  size_t stage_rows = 0;
  cudaStream_t host_stream = nullptr;
};

// Bumped whenever a linear is created or destroyed: invalidates the launch graphs cached by the host-buffer batch call
// (a new handle may reuse a destroyed one's address).
static std::atomic<uint64_t> g_handle_gen{1};

static void free_all(std::vector<void *> &v) {
  for (void *p : v)
    if (p) cudaFree(p);
  v.clear();
}

// Upload one [K, n] GPTQ matrix and produce its device representation:
//   tc:      tiles (+perm handled by caller)     raw tensors are temporary
//   generic: raw tensors kept (qweight, f16 scales, qzeros)
struct RawDev {
  int32_t *qweight = nullptr;
  float *scales_f32 = nullptr;
  int32_t *qzeros = nullptr;
};

static int upload_raw(const int32_t *qweight, const float *scales, const int32_t *qzeros, size_t k,
                      size_t n, size_t group_size, RawDev *r) {
  const size_t groups = k / group_size;
  int rc;
  if ((rc = upload(&r->qweight, qweight, (k / 8) * n))) return rc;
  if ((rc = upload(&r->scales_f32, scales, groups * n))) return rc;
  if ((rc = upload(&r->qzeros, qzeros, groups * (n / 8)))) return rc;
  return FB2_OK;
}
static void free_raw(RawDev *r) {
  if (r->qweight) cudaFree(r->qweight);
  if (r->scales_f32) cudaFree(r->scales_f32);
  if (r->qzeros) cudaFree(r->qzeros);
  *r = RawDev{};
}

static int repack_into(const RawDev &r, const int32_t *d_perm, uint8_t *tiles, size_t k, size_t n) {
  const size_t total = (n / kTileN) * (k / kGroup) * (size_t)(kStageBytes / 4);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  repack_tiles_kernel<<<blocks, 256>>>(r.qweight, r.scales_f32, r.qzeros, d_perm, tiles, (int)k,
                                       (int)n);
  FB2_LAUNCH_CHECK();
  return FB2_OK;
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t tile_bytes(size_t k, size_t n) { return (n / kTileN) * (k / kGroup) * (size_t)kStageBytes; }

// Upper bound of row blocks moe_align_block_size can emit for `pairs` pairs.
static int routed_block_bound(size_t pairs, size_t num_experts, size_t block) {
  const size_t nonempty = std::min(pairs, num_experts);
  return (int)((pairs + nonempty * (block - 1)) / block);
}

// Worst-case split-K scratch of any launch shape this weight can see with <= rows
// activation rows (dense, dispatch and routed forms; with and without the SiLU epilogue).
static void worst_case_requirements(const WeightRef &w, size_t rows, int sms, size_t *pb, size_t *nc) {
  *pb = *nc = 0;
  if (!tc_supported(w) || rows == 0) return;
  auto consider = [&](const RowMap &rm, int epi) {
    size_t a = 0, b = 0;
    tc_requirements(w, rm, epi, sms, &a, &b);
    *pb = std::max(*pb, a);
    *nc = std::max(*nc, b);
  };
  const bool silu_ok = w.n % 2 == 0 && (w.n / 2) % kTileN == 0;
  for (size_t m = 1; m <= rows; m = (m < 16 ? 16 : m * 2)) {     // 1, 16, 32, 64, ... : every nb class
    RowMap rm{};
    rm.mode = ROWS_DENSE;
    rm.m = (int)std::min(m, rows);
    consider(rm, 0);
  }
  {
    RowMap rm{};
    rm.mode = ROWS_DENSE;
    rm.m = (int)rows;
    consider(rm, 0);
  }
  if (w.num_experts > 1) {
    for (int block : {8, 16, 32, 64}) {
      RowMap rm{};
      rm.mode = ROWS_ROUTED;
      rm.block = block;
      rm.max_blocks = routed_block_bound(rows, w.num_experts, block);
      consider(rm, 0);
      if (silu_ok) consider(rm, 1);
    }
    for (int nb : {16, 32, 64}) {
      RowMap rm{};
      rm.mode = ROWS_DISPATCH;
      rm.block = nb;
      rm.num_dispatch = (int)std::min<size_t>(kMaxDispatchBlocks, std::min<size_t>(rows, w.num_experts) + rows / nb);
      consider(rm, 0);
    }
  }
}

static int reserve_scratch_for(const WeightRef &w, int device, size_t rows) {
  if (w.perm) {
    int rc = grow_xperm(device, rows * (size_t)w.k * sizeof(__half));
    if (rc) return rc;
  }
  size_t pb, nc;
  worst_case_requirements(w, rows, g_scratch[device].sm_count, &pb, &nc);
  if (pb == 0 && nc == 0) return FB2_OK;
  return grow_scratch(device, pb, nc);
}

// Launch the tensor-core kernel family, growing the scratch pool first when this launch
// shape needs more than any before it (impossible while the stream is being captured:
// then the caller must have called fb2_pregrow_scratch, as the reference requires of
// pregrow_marlin_gather_scratch, capabilities.rs:205-214).
static int run_tc(const WeightRef &w_in, const RowMap &rm, const DispatchTable *tbl, const __half *x,
                  size_t x_rows, __half *y, int epilogue, int device, cudaStream_t stream,
                  const PushTargets *push = nullptr, int dk_flags = 0, const int32_t *out_map = nullptr) {
  if (((uintptr_t)x % 16) != 0 || ((uintptr_t)y % 2) != 0) {
    // every activation path of the tensor-core kernels moves 16-byte vectors (TMA boxes, cp.async rows, the act-order
    // gathers): a misaligned sub-tensor view is refused here instead of faulting on the device
    set_error("activations must be 16-byte aligned (x = %p) and the output 2-byte aligned (y = %p)", (const void *)x, (void *)y);
    return FB2_ERR_ARG;
  }
  int sms = 0;
  TcScratch sc = current_scratch(device, &sms);
  size_t pb = 0, nc = 0;
  WeightRef w = w_in;
  tc_requirements(w, rm, epilogue, sms, &pb, &nc);
  size_t xb = w.perm ? x_rows * (size_t)w.k * sizeof(__half) : 0;
  size_t have_xb;
  {
    DeviceScratch &s = g_scratch[device];
    std::lock_guard<std::mutex> lock(s.mu);
    have_xb = s.xperm_bytes;
  }
  if (pb > sc.partial_bytes || nc > sc.num_counters || xb > have_xb) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    FB2_CUDA_TRY(cudaStreamIsCapturing(stream, &cs));
    if (cs != cudaStreamCaptureStatusNone) {
      set_error("split-K scratch too small for this launch (%zu B, %zu counters) and the stream is "
                "capturing: call fb2_pregrow_scratch(device, max_m) before capture", pb, nc);
      return FB2_ERR_ARG;
    }
    int rc = grow_scratch(device, pb, nc);
    if (rc) return rc;
    if ((rc = grow_xperm(device, xb))) return rc;
    sc = current_scratch(device, &sms);
  }
  if (w.perm && dk_gather_ok(w, rm, tbl, epilogue) && ((uintptr_t)x % 16) == 0) {
    // act-order, decode batch: the kernel copies the rows of x into shared memory and gathers there
  } else if (w.perm) {
    // act-order: A'[m, j] = A[m, perm[j]] once per GEMM (reference: gather_columns.cu:15 via
    // marlin_gemm_with_perm, quant.rs:434-509); every CTA then loads its slice with 16-byte
    // vectors instead of repeating the gather per weight tile.
    __half *xp = g_scratch[device].xperm;
    int rc = launch_permute_cols(x, w.perm, xp, (int)x_rows, w.k, stream);
    if (rc) return rc;
    x = xp;
    w.perm = nullptr;
  }
  return launch_tc_gemm(w, rm, tbl, x, x_rows, y, epilogue, sc, sms, stream, push, dk_flags, out_map);
}

// Registry of live weights per device so that fb2_pregrow_scratch can size the pool for
// "every linear/stack created so far" (capabilities.rs:212 semantics).
struct RegistryItem {
  const void *owner;
  int device;
  WeightRef w;
};
struct Registry {
  std::mutex mu;
  std::vector<RegistryItem> items;
};
static Registry g_registry;
static void registry_add(const void *owner, int device, const WeightRef &w) {
  std::lock_guard<std::mutex> lock(g_registry.mu);
  g_registry.items.push_back(RegistryItem{owner, device, w});
}
static void registry_remove(const void *owner) {
  std::lock_guard<std::mutex> lock(g_registry.mu);
  auto &v = g_registry.items;
  v.erase(std::remove_if(v.begin(), v.end(), [&](const RegistryItem &it) { return it.owner == owner; }),
          v.end());
}


extern "C" {

const char *fb2_version(void) { return "fb2-w4a16 0.1.0 (sm_100a)"; }
const char *fb2_last_error(void) { return t_error; }
uint64_t fb2_launch_count(void) { return g_launches.load(); }
void fb2_reset_launch_count(void) { g_launches.store(0); }

uint64_t fb2_gemm_algorithmic_bytes(size_t k, size_t n, size_t m, size_t group_size, int act_order) {
  if (group_size == 0) return 0;
  // int4 weights + f16 scales + packed zeros (+ g_idx/perm) + A + C
  return (uint64_t)k * n / 2 + (uint64_t)(k / group_size) * n * 2 +
         (uint64_t)(k / group_size) * (n / 8) * 4 + (act_order ? (uint64_t)k * 4 : 0) +
         (uint64_t)m * k * 2 + (uint64_t)m * n * 2;
}
uint64_t fb2_gemm_flops(size_t k, size_t n, size_t m) { return 2ull * m * n * k; }

// ------------------------------------------------------------------------ dense linear
int fb2_gptq_linear_create(const int32_t *qweight, const float *scales, const int32_t *qzeros,
                           const int32_t *g_idx, const float *bias, uint32_t bits,
                           size_t group_size, size_t k, size_t n, int device, fb2_linear_t **out) {
  if (!out) { set_error("out is NULL"); return FB2_ERR_ARG; }
  *out = nullptr;
  if (!qweight || !scales || !qzeros) { set_error("qweight/scales/qzeros must not be NULL"); return FB2_ERR_ARG; }
  int rc = validate_common(bits, group_size, k, n, "GPTQ linear");
  if (rc) return rc;
  Ingest ing;
  if ((rc = analyse_g_idx(g_idx, group_size, k, n, &ing))) return rc;
  if ((rc = probe_device(device))) return rc;

  fb2_linear *lin = new fb2_linear();
  lin->device = device;
  lin->act_order = ing.act_order;
  WeightRef &w = lin->w;
  w.k = (int)k; w.n = (int)n; w.ld = (int)n; w.col0 = 0;
  w.group_size = (int)group_size; w.num_experts = 1;
  RawDev raw;
  auto fail = [&](int code) {
    free_raw(&raw);
    free_all(lin->allocs);
    delete lin;
    return code;
  };
  if ((rc = upload_raw(qweight, scales, qzeros, k, n, group_size, &raw))) return fail(rc);
  if (ing.tc) {
    int32_t *d_perm = nullptr;
    if (ing.act_order) {
      if ((rc = upload(&d_perm, ing.perm.data(), k))) return fail(rc);
      lin->allocs.push_back(d_perm);
    }
    uint8_t *tiles = nullptr;
    if (cudaMalloc((void **)&tiles, tile_bytes(k, n)) != cudaSuccess) {
      set_error("cudaMalloc(%zu) for weight tiles failed", tile_bytes(k, n));
      return fail(FB2_ERR_CUDA);
    }
    lin->allocs.push_back(tiles);
    if ((rc = repack_into(raw, d_perm, tiles, k, n))) return fail(rc);
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("repack kernel failed"); return fail(FB2_ERR_CUDA); }
    w.tiles = tiles;
    w.perm = d_perm;
    free_raw(&raw);
  } else {
    // generic path keeps the raw tensors (scales re-rounded to f16, quant.rs:717)
    const size_t ns = (k / group_size) * n;
    __half *sc16 = nullptr;
    if (cudaMalloc((void **)&sc16, ns * sizeof(__half)) != cudaSuccess) { set_error("cudaMalloc scales failed"); return fail(FB2_ERR_CUDA); }
    lin->allocs.push_back(sc16);
    f32_to_f16_kernel<<<(unsigned)((ns + 255) / 256), 256>>>(raw.scales_f32, sc16, ns);
    count_launch();
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("scale conversion failed"); return fail(FB2_ERR_CUDA); }
    w.qweight = raw.qweight; w.qzeros = raw.qzeros; w.scales = sc16;
    lin->allocs.push_back(raw.qweight);
    lin->allocs.push_back(raw.qzeros);
    cudaFree(raw.scales_f32);
    raw = RawDev{};
    if (g_idx && ing.act_order) {
      int32_t *d_g = nullptr;
      if ((rc = upload(&d_g, g_idx, k))) return fail(rc);
      lin->allocs.push_back(d_g);
      w.g_idx = d_g;
    }
  }
  if (bias) {
    float *d_b = nullptr;
    if ((rc = upload(&d_b, bias, n))) return fail(rc);
    lin->allocs.push_back(d_b);
    w.bias = d_b;
  }
  if ((rc = reserve_scratch_for(w, device, 64))) return fail(rc);
  registry_add(lin, device, w);
  g_handle_gen.fetch_add(1);
  *out = lin;
  return FB2_OK;
}

void fb2_gptq_linear_destroy(fb2_linear_t *lin) {
  if (!lin) return;
  cudaSetDevice(lin->device);
  registry_remove(lin);
  g_handle_gen.fetch_add(1);
  free_all(lin->allocs);
  if (lin->h_x) cudaFreeHost(lin->h_x);
  if (lin->h_y) cudaFreeHost(lin->h_y);
  if (lin->d_x) cudaFree(lin->d_x);
  if (lin->d_y) cudaFree(lin->d_y);
  if (lin->host_stream) cudaStreamDestroy(lin->host_stream);
  delete lin;
}

size_t fb2_gptq_linear_in_features(const fb2_linear_t *lin) { return lin ? (size_t)lin->w.k : 0; }
size_t fb2_gptq_linear_out_features(const fb2_linear_t *lin) { return lin ? (size_t)lin->w.n : 0; }
int fb2_gptq_linear_is_act_order(const fb2_linear_t *lin) { return lin && lin->act_order ? 1 : 0; }
const char *fb2_gptq_linear_kernel_name(const fb2_linear_t *lin, size_t m) {
  return lin ? tc_kernel_name(lin->w, m) : "none";
}

static WeightRef view_of(const fb2_linear *lin) {
  // Expert views address their expert by offsetting the base pointers, so that kernels see
  // a one-expert weight.
  WeightRef w = lin->w;
  if (lin->expert != 0) {
    const size_t groups = (size_t)w.k / w.group_size;
    if (w.tiles) w.tiles += (size_t)lin->expert * (w.ld / kTileN) * (w.k / kGroup) * kStageBytes;
    if (w.qweight) {
      w.qweight += (size_t)lin->expert * (w.k / 8) * w.ld;
      w.scales += (size_t)lin->expert * groups * w.ld;
      w.qzeros += (size_t)lin->expert * groups * (w.ld / 8);
    }
  }
  w.num_experts = 1;
  return w;
}

}  // extern "C"

namespace fb2 {
// Row-parallel forward whose epilogue pushes the output tiles to every rank (parallel.cu).
// *ctas = CTAs that will signal.  UNSUPPORTED when the layer is not on the tensor-core path.
int forward_linear_push(const fb2_linear_t *lin, cudaStream_t stream, const __half *x, __half *y_unused, size_t m,
                        const PushTargets *push, uint32_t *ctas) {
  if (!lin || !x || !push) { set_error("NULL argument"); return FB2_ERR_ARG; }
  FB2_CUDA_TRY(cudaSetDevice(lin->device));
  WeightRef w = view_of(lin);
  if (!tc_supported(w)) { set_error("fused all-reduce needs the tiled weight format"); return FB2_ERR_UNSUPPORTED; }
  RowMap rm{};
  rm.mode = ROWS_DENSE;
  rm.m = (int)m;
  int sms = 0;
  current_scratch(lin->device, &sms);
  *ctas = (uint32_t)tc_grid_size(w, rm, 0, sms);
  return run_tc(w, rm, nullptr, x, m, y_unused, 0, lin->device, stream, push);
}
}  // namespace fb2

extern "C" {

int fb2_gptq_linear_forward(const fb2_linear_t *lin, fb2_stream_t stream, const fb2_half *x,
                            fb2_half *y, size_t m) {
  if (!lin) { set_error("linear handle is NULL"); return FB2_ERR_ARG; }
  if (m == 0) return FB2_OK;
  if (!x || !y) { set_error("x / y must not be NULL"); return FB2_ERR_ARG; }
  if (m > (size_t)INT32_MAX / 2) { set_error("m too large"); return FB2_ERR_SHAPE; }
  FB2_CUDA_TRY(cudaSetDevice(lin->device));
  WeightRef w = view_of(lin);
  RowMap rm{};
  rm.mode = ROWS_DENSE;
  rm.m = (int)m;
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_supported(w)) return run_tc(w, rm, nullptr, (const __half *)x, m, (__half *)y, 0, lin->device, st);
  return launch_generic_gemm(w, rm, (const __half *)x, (__half *)y, 0, st);
}

static std::mutex &host_call_mutex(int device);   // per device, shared with the batch call (below)

// x already in this layer's act-order column order (x_sorted[:, j] = x[:, perm[j]]): the producer folded the gather
// into its own store (fb2_fused_add_rms_norm with out_perm = fb2_gptq_linear_act_order_perm).
int fb2_gptq_linear_forward_sorted(const fb2_linear_t *lin, fb2_stream_t stream, const fb2_half *x_sorted,
                                   fb2_half *y, size_t m) {
  if (!lin) { set_error("linear handle is NULL"); return FB2_ERR_ARG; }
  if (m == 0) return FB2_OK;
  if (!x_sorted || !y) { set_error("x / y must not be NULL"); return FB2_ERR_ARG; }
  if (m > (size_t)INT32_MAX / 2) { set_error("m too large"); return FB2_ERR_SHAPE; }
  FB2_CUDA_TRY(cudaSetDevice(lin->device));
  WeightRef w = view_of(lin);
  if (!tc_supported(w)) {
    if (w.g_idx) { set_error("forward_sorted: this act-order layer is served by the generic kernel, which reads x in disk order"); return FB2_ERR_UNSUPPORTED; }
    RowMap rmg{}; rmg.mode = ROWS_DENSE; rmg.m = (int)m;
    return launch_generic_gemm(w, rmg, (const __half *)x_sorted, (__half *)y, 0, (cudaStream_t)stream);
  }
  w.perm = nullptr;
  RowMap rm{};
  rm.mode = ROWS_DENSE;
  rm.m = (int)m;
  return run_tc(w, rm, nullptr, (const __half *)x_sorted, m, (__half *)y, 0, lin->device, (cudaStream_t)stream);
}

int fb2_gptq_linear_act_order_perm(const fb2_linear_t *lin, const int32_t **perm_dev) {
  if (!lin || !perm_dev) { set_error("NULL argument"); return FB2_ERR_ARG; }
  *perm_dev = tc_supported(lin->w) ? lin->w.perm : nullptr;
  return FB2_OK;
}

// gate | up linear (n = 2 * inter, gate columns first) with SiLU(gate) * up in the GEMM epilogue: y [m, n / 2].
// (reference: the pair Linear::forward + fused_silu_mul_split, llama_family.rs:3468-3509)
int fb2_gptq_linear_forward_swiglu(const fb2_linear_t *lin, fb2_stream_t stream, const fb2_half *x, fb2_half *y,
                                   size_t m, int x_is_sorted) {
  if (!lin) { set_error("linear handle is NULL"); return FB2_ERR_ARG; }
  if (m == 0) return FB2_OK;
  if (!x || !y) { set_error("x / y must not be NULL"); return FB2_ERR_ARG; }
  if (m > (size_t)INT32_MAX / 2) { set_error("m too large"); return FB2_ERR_SHAPE; }
  FB2_CUDA_TRY(cudaSetDevice(lin->device));
  WeightRef w = view_of(lin);
  if (w.n % 2) { set_error("SwiGLU needs an even number of output features"); return FB2_ERR_SHAPE; }
  if (w.bias) { set_error("SwiGLU epilogue: a gate|up layer with bias is not supported"); return FB2_ERR_UNSUPPORTED; }
  RowMap rm{};
  rm.mode = ROWS_DENSE;
  rm.m = (int)m;
  cudaStream_t st = (cudaStream_t)stream;
  if (tc_supported(w) && (w.n / 2) % kTileN == 0) {
    if (x_is_sorted) w.perm = nullptr;
    return run_tc(w, rm, nullptr, (const __half *)x, m, (__half *)y, 1, lin->device, st);
  }
  if (x_is_sorted && w.g_idx) { set_error("forward_swiglu: sorted input needs the tiled weight format"); return FB2_ERR_UNSUPPORTED; }
  return launch_generic_gemm(w, rm, (const __half *)x, (__half *)y, 1, st);
}

int fb2_gptq_linear_forward_host(fb2_linear_t *lin, const fb2_half *x_host, fb2_half *y_host,
                                 size_t m) {
  if (!lin) { set_error("linear handle is NULL"); return FB2_ERR_ARG; }
  if (m == 0) return FB2_OK;
  if (!x_host || !y_host) { set_error("x / y must not be NULL"); return FB2_ERR_ARG; }
  // All host-buffer calls of one device share its split-K scratch (indexed by CTA and tile, not by launch): they are
  // serialised by the per-device mutex that fb2_gptq_linear_forward_host_batch takes as well.
  std::lock_guard<std::mutex> dev_lock(host_call_mutex(lin->device));
  std::lock_guard<std::mutex> lock(lin->host_mu);
  FB2_CUDA_TRY(cudaSetDevice(lin->device));
  const size_t k = lin->w.k, n = lin->w.n;
  if (m > lin->stage_rows) {
    size_t rows = std::max<size_t>(m, 64);
    if (lin->h_x) cudaFreeHost(lin->h_x);
    if (lin->h_y) cudaFreeHost(lin->h_y);
    if (lin->d_x) cudaFree(lin->d_x);
    if (lin->d_y) cudaFree(lin->d_y);
    lin->h_x = lin->h_y = lin->d_x = lin->d_y = nullptr;
    lin->stage_rows = 0;
    FB2_CUDA_TRY(cudaMallocHost(&lin->h_x, rows * k * 2));
    FB2_CUDA_TRY(cudaMallocHost(&lin->h_y, rows * n * 2));
    FB2_CUDA_TRY(cudaMalloc(&lin->d_x, rows * k * 2));
    FB2_CUDA_TRY(cudaMalloc(&lin->d_y, rows * n * 2));
    if (!lin->host_stream) FB2_CUDA_TRY(cudaStreamCreateWithFlags(&lin->host_stream, cudaStreamNonBlocking));
    int rc = reserve_scratch_for(lin->w, lin->device, rows);
    if (rc) return rc;
    lin->stage_rows = rows;
  }
  memcpy(lin->h_x, x_host, m * k * 2);
  FB2_CUDA_TRY(cudaMemcpyAsync(lin->d_x, lin->h_x, m * k * 2, cudaMemcpyHostToDevice, lin->host_stream));
  int rc = fb2_gptq_linear_forward(lin, lin->host_stream, (const fb2_half *)lin->d_x,
                                   (fb2_half *)lin->d_y, m);
  if (rc) return rc;
  FB2_CUDA_TRY(cudaMemcpyAsync(lin->h_y, lin->d_y, m * n * 2, cudaMemcpyDeviceToHost, lin->host_stream));
  FB2_CUDA_TRY(cudaStreamSynchronize(lin->host_stream));
  memcpy(y_host, lin->h_y, m * n * 2);
  return FB2_OK;
}

// One pinned + one device arena per device for the host-buffer batch call.
// The GEMM launches of one chunk, captured once per (linears, m) and replayed: a repeated batch call (a decode loop) pays one
// graph launch per chunk instead of per-GEMM launch, tensor-map and attribute calls, and gets the kernels' programmatic
// dependent launch edges.  Keyed by the handle list, m, the arena base and the handle generation.
struct BatchGraphs {
  std::vector<const fb2_linear_t *> lins;
  std::vector<size_t> ms;
  uint64_t gen = 0;
  const uint8_t *base = nullptr;
  bool no_graph = false;                       // capture failed once: stay on the eager path
  std::vector<cudaGraphExec_t> chunk;          // empty until the second call with this key
  std::vector<int> chunk_launches;             // kernels in each chunk's graph (fb2_launch_count bookkeeping)
  uint64_t last_use = 0;
};
struct HostArena {
  std::mutex mu;
  uint8_t *h = nullptr, *d = nullptr;
  size_t bytes = 0;
  cudaStream_t stream = nullptr, copy_in = nullptr, copy_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_done;
  std::vector<BatchGraphs> graphs;
  uint64_t tick = 0;
};
static void drop_graphs(BatchGraphs &g) {
  for (cudaGraphExec_t e : g.chunk)
    if (e) cudaGraphExecDestroy(e);
  g.chunk.clear();
}
static HostArena g_arena[kMaxDevices];
static std::mutex &host_call_mutex(int device) { return g_arena[device < 0 || device >= kMaxDevices ? 0 : device].mu; }

// Entry i has ms[i] rows (entries with zero rows are skipped).
int fb2_gptq_linear_forward_host_batch_m(fb2_linear_t *const *lins, size_t n,
                                         const fb2_half *const *x_host, fb2_half *const *y_host,
                                         const size_t *ms) {
  if (n == 0) return FB2_OK;
  if (!lins || !x_host || !y_host || !ms) { set_error("NULL argument"); return FB2_ERR_ARG; }
  size_t max_m = 0;
  for (size_t i = 0; i < n; i++) {
    if (ms[i] > (size_t)INT32_MAX / 2) { set_error("m too large"); return FB2_ERR_SHAPE; }
    max_m = std::max(max_m, ms[i]);
  }
  if (max_m == 0) return FB2_OK;
  const int device = lins[0] ? lins[0]->device : -1;
  size_t in_bytes = 0, out_bytes = 0;
  for (size_t i = 0; i < n; i++) {
    if (!lins[i] || !x_host[i] || !y_host[i]) { set_error("NULL entry %zu", i); return FB2_ERR_ARG; }
    if (lins[i]->device != device) { set_error("all linears of a batch call must share a device"); return FB2_ERR_ARG; }
    in_bytes += align256(ms[i] * lins[i]->w.k * 2);
    out_bytes += align256(ms[i] * lins[i]->w.n * 2);
  }
  FB2_CUDA_TRY(cudaSetDevice(device));
  HostArena &A = g_arena[device];
  std::lock_guard<std::mutex> lock(A.mu);
  const size_t need = in_bytes + out_bytes;
  if (need > A.bytes) {
    if (A.h) cudaFreeHost(A.h);
    if (A.d) cudaFree(A.d);
    A.h = A.d = nullptr;
    A.bytes = 0;
    for (BatchGraphs &g : A.graphs) drop_graphs(g);   // they hold the old arena's addresses
    A.graphs.clear();
    FB2_CUDA_TRY(cudaMallocHost((void **)&A.h, need));
    FB2_CUDA_TRY(cudaMalloc((void **)&A.d, need));
    A.bytes = need;
    for (size_t i = 0; i < n; i++) {
      int rc = reserve_scratch_for(view_of(lins[i]), device, std::max<size_t>(max_m, 64));
      if (rc) return rc;
    }
  }
  if (!A.stream) {
    FB2_CUDA_TRY(cudaStreamCreateWithFlags(&A.stream, cudaStreamNonBlocking));
    FB2_CUDA_TRY(cudaStreamCreateWithFlags(&A.copy_in, cudaStreamNonBlocking));
    FB2_CUDA_TRY(cudaStreamCreateWithFlags(&A.copy_out, cudaStreamNonBlocking));
  }
  // Page-locked caller buffers are DMA'd directly (no staging pass through the arena's pinned half);
  // pageable ones are staged.  Both directions are decided per call from the first entry.
  auto is_pinned = [](const void *ptr) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
  };
  bool pin_in = true, pin_out = true;
  for (size_t i = 0; i < n && (pin_in || pin_out); i++) {
    pin_in = pin_in && is_pinned(x_host[i]);
    pin_out = pin_out && is_pinned(y_host[i]);
  }
  // Three streams, chunks of kChunk linears (8: measured best of 4 / 8 / 16 on the 768-GEMM decode step): the H2D copy
  // of chunk c+1 and the D2H copy of chunk c-1 overlap the GEMMs of chunk c (two copy engines + compute).
  static const size_t kChunk = [] { const char *e = getenv("FB2_HOST_BATCH_CHUNK"); const int v = e ? atoi(e) : 0; return (size_t)(v > 0 ? v : 8); }();
  const size_t chunks = (n + kChunk - 1) / kChunk;
  if (A.ev_in.size() < chunks) {
    const size_t old = A.ev_in.size();
    A.ev_in.resize(chunks);
    A.ev_done.resize(chunks);
    for (size_t c = old; c < chunks; c++) {
      FB2_CUDA_TRY(cudaEventCreateWithFlags(&A.ev_in[c], cudaEventDisableTiming));
      FB2_CUDA_TRY(cudaEventCreateWithFlags(&A.ev_done[c], cudaEventDisableTiming));
    }
  }
  std::vector<size_t> xoff(n), yoff(n);
  {
    size_t xo = 0, yo = in_bytes;
    for (size_t i = 0; i < n; i++) {
      xoff[i] = xo; yoff[i] = yo;
      xo += align256(ms[i] * lins[i]->w.k * 2);
      yo += align256(ms[i] * lins[i]->w.n * 2);
    }
  }
  // cached launch graphs of this (linears, m)
  static const bool use_graphs = [] { const char *e = getenv("FB2_HOST_BATCH_GRAPHS"); return !(e && e[0] == '0'); }();
  BatchGraphs *G = nullptr;
  if (use_graphs) {
    const uint64_t gen = g_handle_gen.load();
    for (BatchGraphs &g : A.graphs)
      if (g.lins.size() == n && std::equal(g.lins.begin(), g.lins.end(), lins) && std::equal(g.ms.begin(), g.ms.end(), ms)) { G = &g; break; }
    if (G && (G->gen != gen || G->base != A.d)) {     // handles came or went: what the graphs point at may be gone
      drop_graphs(*G);
      G->gen = gen;
      G->base = A.d;
      G->no_graph = false;
      G = nullptr;                                    // this call runs eagerly; the next one captures
    } else if (!G) {
      if (A.graphs.size() >= 16) {                    // keep the 16 most recently used keys
        auto lru = std::min_element(A.graphs.begin(), A.graphs.end(),
                                    [](const BatchGraphs &a, const BatchGraphs &b) { return a.last_use < b.last_use; });
        drop_graphs(*lru);
        A.graphs.erase(lru);
      }
      BatchGraphs g;
      g.lins.assign(lins, lins + n);
      g.ms.assign(ms, ms + n);
      g.gen = gen;
      g.base = A.d;
      g.last_use = ++A.tick;
      A.graphs.push_back(std::move(g));               // first sighting: eager (it also grows the scratch pools)
    } else {
      G->last_use = ++A.tick;
      if (G->no_graph) G = nullptr;
    }
  }
  if (G && G->chunk.empty()) {
    // second call with this key: capture the GEMMs chunk by chunk
    G->chunk.assign(chunks, nullptr);
    G->chunk_launches.assign(chunks, 0);
    for (size_t c = 0; c < chunks && !G->no_graph; c++) {
      const size_t i0 = c * kChunk, i1 = std::min(n, i0 + kChunk);
      const uint64_t l0 = g_launches.load();
      if (cudaStreamBeginCapture(A.stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); G->no_graph = true; break; }
      int rc = FB2_OK;
      for (size_t i = i0; i < i1 && rc == FB2_OK; i++)
        rc = fb2_gptq_linear_forward(lins[i], A.stream, (const fb2_half *)(A.d + xoff[i]), (fb2_half *)(A.d + yoff[i]), ms[i]);
      cudaGraph_t graph = nullptr;
      const cudaError_t ce = cudaStreamEndCapture(A.stream, &graph);
      if (rc != FB2_OK || ce != cudaSuccess || !graph ||
          cudaGraphInstantiate(&G->chunk[c], graph, 0) != cudaSuccess) {
        cudaGetLastError();
        G->no_graph = true;
      }
      if (graph) cudaGraphDestroy(graph);
      const uint64_t nl = g_launches.load() - l0;     // counted while capturing: nothing ran yet
      g_launches.fetch_sub(nl);
      G->chunk_launches[c] = (int)nl;
    }
    if (G->no_graph) {
      drop_graphs(*G);
      G = nullptr;
    }
  }
  for (size_t c = 0; c < chunks; c++) {
    const size_t i0 = c * kChunk, i1 = std::min(n, i0 + kChunk);
    for (size_t i = i0; i < i1; i++) {
      const size_t xb = ms[i] * lins[i]->w.k * 2;
      if (xb == 0) continue;
      const void *src = x_host[i];
      if (!pin_in) { memcpy(A.h + xoff[i], x_host[i], xb); src = A.h + xoff[i]; }
      FB2_CUDA_TRY(cudaMemcpyAsync(A.d + xoff[i], src, xb, cudaMemcpyHostToDevice, A.copy_in));
    }
    FB2_CUDA_TRY(cudaEventRecord(A.ev_in[c], A.copy_in));
    FB2_CUDA_TRY(cudaStreamWaitEvent(A.stream, A.ev_in[c], 0));
    if (G) {
      FB2_CUDA_TRY(cudaGraphLaunch(G->chunk[c], A.stream));
      count_launch(G->chunk_launches[c]);
    } else {
      for (size_t i = i0; i < i1; i++) {
        int rc = fb2_gptq_linear_forward(lins[i], A.stream, (const fb2_half *)(A.d + xoff[i]),
                                         (fb2_half *)(A.d + yoff[i]), ms[i]);
        if (rc) return rc;
      }
    }
    FB2_CUDA_TRY(cudaEventRecord(A.ev_done[c], A.stream));
    FB2_CUDA_TRY(cudaStreamWaitEvent(A.copy_out, A.ev_done[c], 0));
    for (size_t i = i0; i < i1; i++) {
      const size_t yb = ms[i] * lins[i]->w.n * 2;
      if (yb == 0) continue;
      void *dst = pin_out ? (void *)y_host[i] : (void *)(A.h + yoff[i]);
      FB2_CUDA_TRY(cudaMemcpyAsync(dst, A.d + yoff[i], yb, cudaMemcpyDeviceToHost, A.copy_out));
    }
  }
  FB2_CUDA_TRY(cudaStreamSynchronize(A.copy_out));
  if (!pin_out)
    for (size_t i = 0; i < n; i++) memcpy(y_host[i], A.h + yoff[i], ms[i] * lins[i]->w.n * 2);
  return FB2_OK;
}

int fb2_gptq_linear_forward_host_batch(fb2_linear_t *const *lins, size_t n,
                                       const fb2_half *const *x_host, fb2_half *const *y_host,
                                       size_t m) {
  if (n == 0 || m == 0) return FB2_OK;
  std::vector<size_t> ms(n, m);
  return fb2_gptq_linear_forward_host_batch_m(lins, n, x_host, y_host, ms.data());
}

int fb2_gptq_linear_dequant_f32(const fb2_linear_t *lin, fb2_stream_t stream, float *w_nk) {
  if (!lin || !w_nk) { set_error("NULL argument"); return FB2_ERR_ARG; }
  FB2_CUDA_TRY(cudaSetDevice(lin->device));
  return launch_dequant(view_of(lin), 0, w_nk, nullptr, (cudaStream_t)stream);
}
int fb2_gptq_linear_dequant_f16(const fb2_linear_t *lin, fb2_stream_t stream, fb2_half *w_nk) {
  if (!lin || !w_nk) { set_error("NULL argument"); return FB2_ERR_ARG; }
  FB2_CUDA_TRY(cudaSetDevice(lin->device));
  return launch_dequant(view_of(lin), 0, nullptr, (__half *)w_nk, (cudaStream_t)stream);
}

int fb2_pregrow_scratch(int device, size_t max_m) {
  int rc = probe_device(device);
  if (rc) return rc;
  std::vector<WeightRef> ws;
  {
    std::lock_guard<std::mutex> lock(g_registry.mu);
    for (auto &it : g_registry.items)
      if (it.device == device) ws.push_back(it.w);
  }
  size_t pb = 0, nc = 0;
  for (auto &w : ws) {
    size_t a, b;
    worst_case_requirements(w, max_m, g_scratch[device].sm_count, &a, &b);
    pb = std::max(pb, a);
    nc = std::max(nc, b);
  }
  {
    DeviceScratch &s = g_scratch[device];
    std::lock_guard<std::mutex> lock(s.mu);
    pb = std::max(pb, s.want_partial_bytes);
    nc = std::max(nc, s.want_counters);
  }
  return grow_scratch(device, pb, nc);
}

// ------------------------------------------------------------------------------ stacks
int fb2_gptq_stack_create(const int32_t *const *qweights, const float *const *scales,
                          const int32_t *const *qzeros, const int32_t *g_idx, uint32_t bits,
                          size_t group_size, size_t k, size_t n_per_expert, size_t num_experts,
                          int device, fb2_stack_t **out) {
  if (!out) { set_error("out is NULL"); return FB2_ERR_ARG; }
  *out = nullptr;
  if (!qweights || !scales || !qzeros || num_experts == 0) {
    set_error("stacked GPTQ: empty expert list");            // quant.rs:781 "load_gptq_stacked: empty"
    return FB2_ERR_MODEL;
  }
  for (size_t e = 0; e < num_experts; e++)
    if (!qweights[e] || !scales[e] || !qzeros[e]) { set_error("expert %zu has a NULL tensor", e); return FB2_ERR_ARG; }
  int rc = validate_common(bits, group_size, k, n_per_expert, "stacked GPTQ");
  if (rc) return rc;
  Ingest ing;
  if ((rc = analyse_g_idx(g_idx, group_size, k, n_per_expert, &ing))) return rc;
  if ((rc = probe_device(device))) return rc;

  fb2_stack *st = new fb2_stack();
  st->device = device;
  st->act_order = ing.act_order;
  WeightRef &w = st->w;
  w.k = (int)k; w.n = (int)n_per_expert; w.ld = (int)n_per_expert; w.col0 = 0;
  w.group_size = (int)group_size; w.num_experts = (int)num_experts;
  RawDev raw;
  auto fail = [&](int code) {
    free_raw(&raw);
    free_all(st->allocs);
    delete st;
    return code;
  };
  const size_t groups = k / group_size;
  if (ing.tc) {
    int32_t *d_perm = nullptr;
    if (ing.act_order) {
      if ((rc = upload(&d_perm, ing.perm.data(), k))) return fail(rc);
      st->allocs.push_back(d_perm);
      std::vector<int32_t> inv(k);
      for (size_t i = 0; i < k; i++) inv[ing.perm[i]] = (int32_t)i;
      int32_t *d_inv = nullptr;
      if ((rc = upload(&d_inv, inv.data(), k))) return fail(rc);
      st->allocs.push_back(d_inv);
      w.inv_perm = d_inv;
    }
    const size_t per_expert = tile_bytes(k, n_per_expert);
    uint8_t *tiles = nullptr;
    if (cudaMalloc((void **)&tiles, per_expert * num_experts) != cudaSuccess) {
      set_error("cudaMalloc(%zu) for stacked weight tiles failed", per_expert * num_experts);
      return fail(FB2_ERR_CUDA);
    }
    st->allocs.push_back(tiles);
    // one staging set, re-used per expert (cudaMemcpy from pageable memory is synchronous)
    if (cudaMalloc((void **)&raw.qweight, (k / 8) * n_per_expert * 4) != cudaSuccess ||
        cudaMalloc((void **)&raw.scales_f32, groups * n_per_expert * 4) != cudaSuccess ||
        cudaMalloc((void **)&raw.qzeros, groups * (n_per_expert / 8) * 4) != cudaSuccess) {
      set_error("cudaMalloc for stacked staging failed");
      return fail(FB2_ERR_CUDA);
    }
    for (size_t e = 0; e < num_experts; e++) {
      if (cudaMemcpy(raw.qweight, qweights[e], (k / 8) * n_per_expert * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(raw.scales_f32, scales[e], groups * n_per_expert * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
          cudaMemcpy(raw.qzeros, qzeros[e], groups * (n_per_expert / 8) * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("H2D copy of expert %zu failed", e);
        return fail(FB2_ERR_CUDA);
      }
      if ((rc = repack_into(raw, d_perm, tiles + e * per_expert, k, n_per_expert))) return fail(rc);
    }
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("stacked repack failed"); return fail(FB2_ERR_CUDA); }
    free_raw(&raw);
    w.tiles = tiles;
    w.perm = d_perm;
  } else {
    const size_t qw_e = (k / 8) * n_per_expert, sc_e = groups * n_per_expert, qz_e = groups * (n_per_expert / 8);
    int32_t *d_qw = nullptr, *d_qz = nullptr;
    __half *d_sc = nullptr;
    float *d_tmp = nullptr;
    if (cudaMalloc((void **)&d_qw, qw_e * num_experts * 4) != cudaSuccess) { set_error("cudaMalloc failed"); return fail(FB2_ERR_CUDA); }
    st->allocs.push_back(d_qw);
    if (cudaMalloc((void **)&d_qz, qz_e * num_experts * 4) != cudaSuccess) { set_error("cudaMalloc failed"); return fail(FB2_ERR_CUDA); }
    st->allocs.push_back(d_qz);
    if (cudaMalloc((void **)&d_sc, sc_e * num_experts * 2) != cudaSuccess) { set_error("cudaMalloc failed"); return fail(FB2_ERR_CUDA); }
    st->allocs.push_back(d_sc);
    if (cudaMalloc((void **)&d_tmp, sc_e * 4) != cudaSuccess) { set_error("cudaMalloc failed"); return fail(FB2_ERR_CUDA); }
    for (size_t e = 0; e < num_experts; e++) {
      bool ok = cudaMemcpy(d_qw + e * qw_e, qweights[e], qw_e * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
                cudaMemcpy(d_qz + e * qz_e, qzeros[e], qz_e * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
                cudaMemcpy(d_tmp, scales[e], sc_e * 4, cudaMemcpyHostToDevice) == cudaSuccess;
      if (!ok) { cudaFree(d_tmp); set_error("H2D copy of expert %zu failed", e); return fail(FB2_ERR_CUDA); }
      f32_to_f16_kernel<<<(unsigned)((sc_e + 255) / 256), 256>>>(d_tmp, d_sc + e * sc_e, sc_e);
      count_launch();
    }
    cudaError_t se = cudaDeviceSynchronize();
    cudaFree(d_tmp);
    if (se != cudaSuccess) { set_error("scale conversion failed"); return fail(FB2_ERR_CUDA); }
    w.qweight = d_qw; w.qzeros = d_qz; w.scales = d_sc;
    if (g_idx && ing.act_order) {
      int32_t *d_g = nullptr;
      if ((rc = upload(&d_g, g_idx, k))) return fail(rc);
      st->allocs.push_back(d_g);
      w.g_idx = d_g;
    }
  }
  if ((rc = reserve_scratch_for(w, device, 64))) return fail(rc);
  registry_add(st, device, w);
  *out = st;
  return FB2_OK;
}

void fb2_gptq_stack_destroy(fb2_stack_t *st) {
  if (!st) return;
  cudaSetDevice(st->device);
  registry_remove(st);
  free_all(st->allocs);
  delete st;
}

size_t fb2_gptq_stack_n_per_expert(const fb2_stack_t *st) { return st ? (size_t)st->w.n : 0; }
size_t fb2_gptq_stack_k(const fb2_stack_t *st) { return st ? (size_t)st->w.k : 0; }
size_t fb2_gptq_stack_num_experts(const fb2_stack_t *st) { return st ? (size_t)st->w.num_experts : 0; }
int fb2_gptq_stack_zero_workspace(const fb2_stack_t *st, fb2_stream_t stream) {
  (void)stream;
  if (!st) { set_error("stack handle is NULL"); return FB2_ERR_ARG; }
  return FB2_OK;   // no lock workspace in this design
}

int fb2_gptq_stack_make_expert_linear(const fb2_stack_t *st, size_t expert_offset, size_t expert_n,
                                      const float *bias, fb2_linear_t **out) {
  if (!out) { set_error("out is NULL"); return FB2_ERR_ARG; }
  *out = nullptr;
  if (!st) { set_error("stack handle is NULL"); return FB2_ERR_ARG; }
  const size_t npe = st->w.n, total = npe * st->w.num_experts;
  if (expert_n == 0 || expert_offset + expert_n > total) {
    // cpu_marlin_stack.rs:100-105 "make_expert_linear OOB"
    set_error("make_expert_linear OOB: offset %zu + n %zu > stacked_n %zu", expert_offset, expert_n, total);
    return FB2_ERR_MODEL;
  }
  const size_t e = expert_offset / npe, col0 = expert_offset % npe;
  if (col0 + expert_n > npe) {
    set_error("make_expert_linear: slice [%zu, +%zu) crosses an expert boundary", expert_offset, expert_n);
    return FB2_ERR_UNSUPPORTED;
  }
  const size_t align = st->w.tiles ? kTileN : 8;
  if (col0 % align || expert_n % align) {
    set_error("make_expert_linear: offset/n must be multiples of %zu", align);
    return FB2_ERR_SHAPE;
  }
  FB2_CUDA_TRY(cudaSetDevice(st->device));
  fb2_linear *lin = new fb2_linear();
  lin->device = st->device;
  lin->act_order = st->act_order;
  lin->expert = (int)e;
  lin->w = st->w;
  lin->w.col0 = (int)col0;
  lin->w.n = (int)expert_n;
  lin->w.bias = nullptr;
  if (bias) {
    float *d_b = nullptr;
    int rc = upload(&d_b, bias, expert_n);
    if (rc) { delete lin; return rc; }
    lin->allocs.push_back(d_b);
    lin->w.bias = d_b;
  }
  int rc = reserve_scratch_for(view_of(lin), st->device, 64);
  if (rc) { free_all(lin->allocs); delete lin; return rc; }
  *out = lin;
  return FB2_OK;
}

int fb2_gptq_stack_dequant_f32(const fb2_stack_t *st, fb2_stream_t stream, float *w_enk) {
  if (!st || !w_enk) { set_error("NULL argument"); return FB2_ERR_ARG; }
  FB2_CUDA_TRY(cudaSetDevice(st->device));
  for (int e = 0; e < st->w.num_experts; e++) {
    int rc = launch_dequant(st->w, e, w_enk + (size_t)e * st->w.n * st->w.k, nullptr,
                            (cudaStream_t)stream);
    if (rc) return rc;
  }
  return FB2_OK;
}

int fb2_gptq_stack_gemm_batched(const fb2_stack_t *st, fb2_stream_t stream, const fb2_half *x,
                                const fb2_dispatch_t *dispatches, size_t nd, fb2_half *y, size_t k) {
  if (!st) { set_error("stack handle is NULL"); return FB2_ERR_ARG; }
  if (nd == 0) return FB2_OK;
  if (!x || !y || !dispatches) { set_error("NULL argument"); return FB2_ERR_ARG; }
  if (k != (size_t)st->w.k) {
    set_error("gemm_phase_batched: k %zu != stack k %d", k, st->w.k);
    return FB2_ERR_SHAPE;
  }
  FB2_CUDA_TRY(cudaSetDevice(st->device));
  int64_t max_m = 0;
  size_t x_rows = 0;
  for (size_t i = 0; i < nd; i++) {
    const fb2_dispatch_t &d = dispatches[i];
    if (d.in_row >= 0 && d.m > 0) x_rows = std::max<size_t>(x_rows, (size_t)(d.in_row + d.m));
    if (d.expert < 0 || d.expert >= st->w.num_experts || d.in_row < 0 || d.out_row < 0 || d.m < 0 ||
        d.m > INT32_MAX / 4 || d.in_row > INT32_MAX / 2 || d.out_row > INT32_MAX / 2) {
      set_error("gemm_phase_batched: dispatch %zu (expert %lld, in %lld, out %lld, m %lld) invalid",
                i, (long long)d.expert, (long long)d.in_row, (long long)d.out_row, (long long)d.m);
      return FB2_ERR_ARG;
    }
    max_m = std::max(max_m, d.m);
  }
  if (max_m == 0) return FB2_OK;
  const bool tc = tc_supported(st->w);
  const int nb = tc ? (max_m <= 16 ? 16 : max_m <= 32 ? 32 : 64) : 8;
  // split every dispatch into row blocks of <= nb rows; launch in chunks of the table size
  std::vector<DispatchEntry> blocks;
  for (size_t i = 0; i < nd; i++) {
    const fb2_dispatch_t &d = dispatches[i];
    for (int64_t r = 0; r < d.m; r += nb)
      blocks.push_back(DispatchEntry{(int)d.expert, (int)(d.in_row + r), (int)(d.out_row + r),
                                     (int)std::min<int64_t>(nb, d.m - r)});
  }
  for (size_t b0 = 0; b0 < blocks.size(); b0 += kMaxDispatchBlocks) {
    const size_t cnt = std::min<size_t>(kMaxDispatchBlocks, blocks.size() - b0);
    DispatchTable tbl;
    memcpy(tbl.e, blocks.data() + b0, cnt * sizeof(DispatchEntry));
    RowMap rm{};
    rm.mode = ROWS_DISPATCH;
    rm.num_dispatch = (int)cnt;
    rm.block = nb;
    int rc = tc ? run_tc(st->w, rm, &tbl, (const __half *)x, x_rows, (__half *)y, 0, st->device,
                         (cudaStream_t)stream)
                : launch_generic_gemm_dispatch(st->w, rm, tbl, (const __half *)x, (__half *)y,
                                               (cudaStream_t)stream);
    if (rc) return rc;
  }
  return FB2_OK;
}

static int routed_gemm(const fb2_stack_t *st, fb2_stream_t stream, const fb2_half *x,
                       const int32_t *sorted_token_ids, const int32_t *expert_ids,
                       const int32_t *num_tokens_post_padded, fb2_half *y, size_t prob_m,
                       size_t block, size_t top_k, size_t sorted_capacity, int epilogue, int dk_flags = 0,
                       const int32_t *out_map = nullptr, bool x_is_sorted = false) {
  if (!st) { set_error("stack handle is NULL"); return FB2_ERR_ARG; }
  if (prob_m == 0) return FB2_OK;
  if (!x || !y || !sorted_token_ids || !expert_ids || !num_tokens_post_padded) {
    set_error("NULL argument");
    return FB2_ERR_ARG;
  }
  if (top_k == 0) { set_error("top_k must be > 0"); return FB2_ERR_ARG; }
  if (block != 8 && block != 16 && block != 32 && block != 64) {
    set_error("moe_block_size %zu not in {8,16,32,64}", block);
    return FB2_ERR_UNSUPPORTED;
  }
  if (epilogue && st->w.n % 2) { set_error("SiLU epilogue needs an even n_per_expert"); return FB2_ERR_SHAPE; }
  FB2_CUDA_TRY(cudaSetDevice(st->device));
  const size_t pairs = prob_m * top_k;
  RowMap rm{};
  rm.mode = ROWS_ROUTED;
  rm.sorted_ids = sorted_token_ids;
  rm.block_experts = expert_ids;
  rm.num_post_padded = num_tokens_post_padded;
  rm.block = (int)block;
  rm.top_k = (int)top_k;
  rm.num_valid = (int)pairs;
  int bound = routed_block_bound(pairs, st->w.num_experts, block);
  if (sorted_capacity) bound = std::min<int>(bound, (int)(sorted_capacity / block));
  rm.max_blocks = bound;
  if (bound <= 0) return FB2_OK;
  const bool tc = tc_supported(st->w) && (!epilogue || (st->w.n / 2) % kTileN == 0);
  if (tc) {
    WeightRef w = st->w;
    if (x_is_sorted) w.perm = nullptr;         // the caller delivers the activations in this stack's sorted-k order
    return run_tc(w, rm, nullptr, (const __half *)x, prob_m, (__half *)y, epilogue, st->device, (cudaStream_t)stream,
                  nullptr, dk_flags, out_map);
  }
  if (x_is_sorted || out_map) { set_error("internal: fused act-order FFN needs the tiled weight format"); return FB2_ERR_ARG; }
  return launch_generic_gemm(st->w, rm, (const __half *)x, (__half *)y, epilogue, (cudaStream_t)stream);
}

int fb2_gptq_stack_gemm_routed(const fb2_stack_t *st, fb2_stream_t stream, const fb2_half *x,
                               const int32_t *sorted_token_ids, const int32_t *expert_ids,
                               const int32_t *num_tokens_post_padded, fb2_half *y, size_t prob_m,
                               size_t block, size_t top_k, size_t sorted_capacity) {
  return routed_gemm(st, stream, x, sorted_token_ids, expert_ids, num_tokens_post_padded, y, prob_m,
                     block, top_k, sorted_capacity, 0);
}

int fb2_gptq_stack_gemm_routed_silu(const fb2_stack_t *st, fb2_stream_t stream, const fb2_half *x,
                                    const int32_t *sorted_token_ids, const int32_t *expert_ids,
                                    const int32_t *num_tokens_post_padded, fb2_half *act,
                                    size_t prob_m, size_t block, size_t top_k,
                                    size_t sorted_capacity) {
  return routed_gemm(st, stream, x, sorted_token_ids, expert_ids, num_tokens_post_padded, act,
                     prob_m, block, top_k, sorted_capacity, 1);
}

// ----------------------------------------------------------------- MoE routing / post-ops
int fb2_moe_route_topk_softmax(fb2_stream_t stream, const fb2_half *logits, int32_t *out_ids,
                               float *out_weights, size_t batch, size_t num_experts, size_t top_k,
                               int norm_topk_prob) {
  if (top_k == 0) { set_error("top_k must be > 0"); return FB2_ERR_ARG; }                 // router.rs:124
  if (top_k > num_experts) {                                                              // router.rs:125-130
    set_error("top_k %zu exceeds num_experts %zu", top_k, num_experts);
    return FB2_ERR_ARG;
  }
  if (batch == 0) return FB2_OK;
  if (!logits || !out_ids || !out_weights) { set_error("NULL argument"); return FB2_ERR_ARG; }
  if (num_experts > 12000) { set_error("num_experts %zu too large for the router kernel", num_experts); return FB2_ERR_UNSUPPORTED; }
  return launch_route_topk_softmax((const __half *)logits, out_ids, out_weights, (int)batch,
                                   (int)num_experts, (int)top_k, norm_topk_prob, (cudaStream_t)stream);
}

int fb2_moe_build_pairs_by_token(fb2_stream_t stream, const int32_t *expert_ids,
                                 int32_t *pairs_by_token, int32_t *packed_token_idx,
                                 int32_t *expert_offsets, size_t batch_x_topk, size_t num_experts,
                                 size_t top_k) {
  if (!expert_ids || !pairs_by_token || !packed_token_idx || !expert_offsets) { set_error("NULL argument"); return FB2_ERR_ARG; }
  if (top_k == 0 || num_experts == 0) { set_error("top_k / num_experts must be > 0"); return FB2_ERR_ARG; }
  return launch_build_pairs(expert_ids, pairs_by_token, packed_token_idx, expert_offsets,
                            (int)batch_x_topk, (int)num_experts, (int)top_k, (cudaStream_t)stream);
}

int fb2_moe_align_block_size(fb2_stream_t stream, const int32_t *expert_ids_per_pair,
                             int32_t *sorted_token_ids, int32_t *block_ids,
                             int32_t *total_tokens_post_pad, size_t batch_x_topk,
                             size_t num_experts, size_t block_size, size_t sorted_max_size,
                             int pair_ids) {
  if (!expert_ids_per_pair || !sorted_token_ids || !block_ids || !total_tokens_post_pad) { set_error("NULL argument"); return FB2_ERR_ARG; }
  if (block_size == 0 || num_experts == 0) { set_error("block_size / num_experts must be > 0"); return FB2_ERR_ARG; }
  return launch_align_block_size(expert_ids_per_pair, sorted_token_ids, block_ids,
                                 total_tokens_post_pad, (int)batch_x_topk, (int)num_experts,
                                 (int)block_size, (int)sorted_max_size, pair_ids,
                                 (cudaStream_t)stream);
}

int fb2_moe_combine(fb2_stream_t stream, const fb2_half *packed_down, const int32_t *pairs_by_token,
                    const float *pair_weights, fb2_half *out, size_t batch, size_t hidden,
                    size_t top_k) {
  if (batch == 0) return FB2_OK;
  if (!packed_down || !pairs_by_token || !pair_weights || !out) { set_error("NULL argument"); return FB2_ERR_ARG; }
  if (hidden % 2) { set_error("hidden must be even"); return FB2_ERR_SHAPE; }
  return launch_moe_combine((const __half *)packed_down, pairs_by_token, pair_weights, (__half *)out,
                            (int)batch, (int)hidden, (int)top_k, (cudaStream_t)stream);
}

int fb2_moe_weighted_sum_batched(fb2_stream_t stream, const fb2_half *slots, const float *weights,
                                 fb2_half *out, size_t batch, size_t top_k, size_t hidden) {
  if (batch == 0) return FB2_OK;
  if (!slots || !weights || !out) { set_error("NULL argument"); return FB2_ERR_ARG; }
  if (hidden % 2) { set_error("hidden must be even"); return FB2_ERR_SHAPE; }
  return launch_moe_combine((const __half *)slots, nullptr, weights, (__half *)out, (int)batch,
                            (int)hidden, (int)top_k, (cudaStream_t)stream);
}

int fb2_fused_silu_mul_split(fb2_stream_t stream, const fb2_half *gate_up, fb2_half *out,
                             size_t tokens, size_t im) {
  if (tokens * im == 0) return FB2_OK;
  if (!gate_up || !out) { set_error("NULL argument"); return FB2_ERR_ARG; }
  return launch_silu_mul_split((const __half *)gate_up, (__half *)out, tokens, im, 0, (cudaStream_t)stream);
}
int fb2_fused_gelu_tanh_mul_split(fb2_stream_t stream, const fb2_half *gate_up, fb2_half *out,
                                  size_t tokens, size_t im) {
  if (tokens * im == 0) return FB2_OK;
  if (!gate_up || !out) { set_error("NULL argument"); return FB2_ERR_ARG; }
  return launch_silu_mul_split((const __half *)gate_up, (__half *)out, tokens, im, 1, (cudaStream_t)stream);
}

int fb2_fused_add_rms_norm(fb2_stream_t stream, const fb2_half *input, fb2_half *residual, const fb2_half *weight,
                           fb2_half *output, size_t tokens, size_t hidden, float eps, const int32_t *out_perm) {
  if (tokens * hidden == 0) return FB2_OK;
  if (!input || !weight || !output) { set_error("NULL argument"); return FB2_ERR_ARG; }
  if (hidden > 64 * 1024) { set_error("hidden %zu too large for the one-CTA-per-token norm", hidden); return FB2_ERR_UNSUPPORTED; }
  if (input == output && out_perm) { set_error("a permuted output cannot alias the input"); return FB2_ERR_ARG; }
  return launch_add_rms_norm((const __half *)input, (__half *)residual, (const __half *)weight, (__half *)output, out_perm,
                             (int)tokens, (int)hidden, eps, (cudaStream_t)stream);
}

int fb2_argmax_rows_f16(fb2_stream_t stream, const fb2_half *logits, size_t m, size_t n, int32_t *out_idx) {
  if (m == 0) return FB2_OK;
  if (!logits || !out_idx || n == 0) { set_error("NULL / empty argument"); return FB2_ERR_ARG; }
  return launch_argmax_rows((const __half *)logits, (int)m, (int)n, out_idx, (cudaStream_t)stream);
}

// --------------------------------------------------------------------- fused MoE FFN
// Workspace layout (all 256-byte aligned):
//   ids [pairs] i32 | weights [pairs] f32 | sorted [cap] i32 | block_ids [cap/block + 1] i32 |
//   total [1] i32 | act [pairs, I] f16 | down [pairs, H] f16
constexpr size_t kFfnBlock = 16;
struct FfnLayout {
  size_t ids, wts, sorted, block_ids, total, act, down, xperm, bytes, cap;
};
static FfnLayout ffn_layout(size_t batch, size_t top_k, size_t num_experts, size_t inter, size_t hidden) {
  FfnLayout L;
  const size_t pairs = batch * top_k;
  L.cap = pairs + num_experts * kFfnBlock;           // dispatch.rs:1865 sorted_max_size
  size_t off = 0;
  L.ids = off; off = align256(off + pairs * 4);
  L.wts = off; off = align256(off + pairs * 4);
  L.sorted = off; off = align256(off + L.cap * 4);
  L.block_ids = off; off = align256(off + (L.cap / kFfnBlock + 1) * 4);
  L.total = off; off = align256(off + 4);
  L.act = off; off = align256(off + pairs * inter * 2);
  L.down = off; off = align256(off + pairs * hidden * 2);
  L.xperm = off; off = align256(off + batch * hidden * 2);      // act-order gate|up stack: x in its sorted-k order
  L.bytes = off;
  return L;
}

size_t fb2_moe_ffn_workspace_bytes(const fb2_stack_t *gate_up, const fb2_stack_t *down,
                                   size_t max_batch, size_t top_k) {
  if (!gate_up || !down) return 0;
  return ffn_layout(max_batch, top_k, gate_up->w.num_experts, gate_up->w.n / 2, down->w.n).bytes;
}

int fb2_moe_ffn_forward(const fb2_stack_t *gate_up, const fb2_stack_t *down, fb2_stream_t stream,
                        const fb2_half *x, const fb2_half *router_logits, fb2_half *out,
                        size_t batch, size_t top_k, int norm_topk_prob, void *workspace,
                        size_t workspace_bytes) {
  if (!gate_up || !down) { set_error("stack handle is NULL"); return FB2_ERR_ARG; }
  if (batch == 0) return FB2_OK;
  if (!x || !router_logits || !out || !workspace) { set_error("NULL argument"); return FB2_ERR_ARG; }
  const size_t E = gate_up->w.num_experts, inter = gate_up->w.n / 2, hidden = down->w.n;
  if ((size_t)down->w.num_experts != E || (size_t)down->w.k != inter || (size_t)gate_up->w.k != hidden ||
      gate_up->w.n % 2 || gate_up->device != down->device) {
    set_error("moe_ffn: gate_up [E=%zu, K=%d, N=%d] and down [E=%d, K=%d, N=%d] do not form an FFN",
              E, gate_up->w.k, gate_up->w.n, down->w.num_experts, down->w.k, down->w.n);
    return FB2_ERR_SHAPE;
  }
  const FfnLayout L = ffn_layout(batch, top_k, E, inter, hidden);
  if (workspace_bytes < L.bytes) {
    set_error("moe_ffn: workspace %zu < required %zu bytes", workspace_bytes, L.bytes);
    return FB2_ERR_ARG;
  }
  uint8_t *ws = (uint8_t *)workspace;
  int32_t *ids = (int32_t *)(ws + L.ids);
  float *wts = (float *)(ws + L.wts);
  int32_t *sorted = (int32_t *)(ws + L.sorted);
  int32_t *block_ids = (int32_t *)(ws + L.block_ids);
  int32_t *total = (int32_t *)(ws + L.total);
  fb2_half *act = (fb2_half *)(ws + L.act);
  fb2_half *dn = (fb2_half *)(ws + L.down);
  const size_t pairs = batch * top_k;
  // Decode batches (on average <= 2 pairs per expert): blocks of 8 rows, served by the decode kernel.  The pair of
  // grouped GEMMs is then chained: the gate|up kernel releases its dependents only after its own grid dependency
  // has resolved (the routing arrays are complete by then), so the down kernel may plan and stream its weights
  // without waiting for the gate|up kernel to finish.
  const size_t block = pairs <= 2 * E ? 8 : kFfnBlock;
  bool chained = false;
  {
    RowMap rm{};
    rm.mode = ROWS_ROUTED;
    rm.block = (int)block;
    chained = tc_supported(gate_up->w) && tc_supported(down->w) && (gate_up->w.n / 2) % kTileN == 0 &&
              dk_routed_ok(gate_up->w, rm, 1) && dk_routed_ok(down->w, rm, 0);
  }
  int rc;
  // Act-order stacks in the chained form need no gather launch of their own: the routing launch also gathers x into
  // the gate|up stack's sorted-k order (CTAs 1..batch), and the gate|up epilogue stores act[:, c] at the column the
  // down stack's order wants (its inverse permutation), so the down GEMM reads act as it lies.
  const bool small_route = route_align_small_supported((int)batch, (int)E, (int)top_k);
  const bool fuse_x = chained && small_route && gate_up->w.perm != nullptr;
  const bool fuse_act = chained && down->w.perm != nullptr && down->w.inv_perm != nullptr;
  const fb2_half *x_gu = x;
  if (small_route) {
    // 1+2 for decode batches: routing and block alignment in one launch
    FB2_CUDA_TRY(cudaSetDevice(gate_up->device));
    __half *xp = fuse_x ? (__half *)(ws + L.xperm) : nullptr;
    if ((rc = launch_route_align_small((const __half *)router_logits, ids, wts, sorted, block_ids, total, (int)batch,
                                       (int)E, (int)top_k, norm_topk_prob, (int)block, (int)L.cap,
                                       (cudaStream_t)stream, fuse_x ? (const __half *)x : nullptr,
                                       fuse_x ? gate_up->w.perm : nullptr, xp, (int)hidden))) return rc;
    if (fuse_x) x_gu = (const fb2_half *)xp;
  } else {
    // 1. route (dispatch.rs:1645)
    if ((rc = fb2_moe_route_topk_softmax(stream, router_logits, ids, wts, batch, E, top_k, norm_topk_prob))) return rc;
    // 2. align with pair ids: A row = id / top_k (gather fused), C row = id (dispatch.rs:1998-2009)
    if ((rc = fb2_moe_align_block_size(stream, ids, sorted, block_ids, total, pairs, E, block, L.cap, 1))) return rc;
  }
  // 3+4. gate-up grouped GEMM with SwiGLU epilogue -> act [pairs, I]
  if ((rc = routed_gemm(gate_up, stream, x_gu, sorted, block_ids, total, act, batch, block, top_k, L.cap, 1,
                        chained ? DK_FLAG_LATE_TRIGGER : 0, fuse_act ? down->w.inv_perm : nullptr, fuse_x))) return rc;
  // 5. down grouped GEMM: rows are pair ids on both sides (prob_m = pairs, top_k = 1)
  if ((rc = routed_gemm(down, stream, act, sorted, block_ids, total, dn, pairs, block, 1, L.cap, 0,
                        chained ? DK_FLAG_META_EARLY : 0, nullptr, fuse_act))) return rc;
  // 6. combine: dense [batch, top_k, hidden] weighted sum (dispatch.rs:2147-2159)
  return fb2_moe_weighted_sum_batched(stream, dn, wts, out, batch, top_k, hidden);
}

}  // extern "C"
