// parallel.cu -- tensor-parallel sharding of GPTQ tensors (host side, C++) and the NCCL
// communicator behind the C ABI (BackendCollective, crates/ferrum-kernels/src/backend/
// capabilities.rs:85-107; NcclRank, crates/ferrum-kernels/src/backend/cuda/nccl_comm.rs:12-105).
//
// Shard definitions follow crates/ferrum-engine/src/parallel/tensor_parallel.rs:171-195 and
// crates/ferrum-models/tests/tp_sharding_test.rs:7-170:
//   column-parallel: split N; every fused segment (q|k|v, gate|up) is split separately and the
//       rank's pieces are re-fused;
//   row-parallel: split K on group boundaries (act-order: in g_idx-SORTED order, so every rank
//       owns whole groups and only its own scales / zeros); partial [M, N] outputs are summed by
//       one all-reduce.
// NCCL is bound at run time (dlopen of the libnccl.so.2 that the host process already carries,
// e.g. the one bundled with PyTorch): the library has no link-time NCCL dependency and every
// fb2_comm_* call fails with FB2_ERR_NCCL when it cannot be found.
#include <dlfcn.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace fb2 {
namespace {

// ------------------------------------------------------------------------ NCCL binding
// Minimal declarations of the NCCL 2.x C API (stable since 2.0; ncclUniqueId is 128 bytes).
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm *NcclComm;
enum { kNcclSuccess = 0, kNcclHalf = 6 /* ncclFloat16 */, kNcclSum = 0 };

struct Nccl {
  void *lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool ok = false;
};

Nccl &nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {getenv("FB2_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      if (!nm || !*nm) continue;
      n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (n.lib) break;
    }
    if (!n.lib) return;
#define FB2_SYM(field, sym) *(void **)(&n.field) = dlsym(n.lib, sym)
    FB2_SYM(GetUniqueId, "ncclGetUniqueId");
    FB2_SYM(CommInitRank, "ncclCommInitRank");
    FB2_SYM(CommDestroy, "ncclCommDestroy");
    FB2_SYM(AllReduce, "ncclAllReduce");
    FB2_SYM(AllGather, "ncclAllGather");
    FB2_SYM(Broadcast, "ncclBroadcast");
    FB2_SYM(Send, "ncclSend");
    FB2_SYM(Recv, "ncclRecv");
    FB2_SYM(GroupStart, "ncclGroupStart");
    FB2_SYM(GroupEnd, "ncclGroupEnd");
    FB2_SYM(GetErrorString, "ncclGetErrorString");
#undef FB2_SYM
    n.ok = n.GetUniqueId && n.CommInitRank && n.CommDestroy && n.AllReduce && n.AllGather && n.Broadcast &&
           n.Send && n.Recv && n.GroupStart && n.GroupEnd;
  });
  return n;
}

int nccl_ready() {
  if (!nccl().ok) {
    set_error("NCCL is not available (libnccl.so.2 could not be opened; set FB2_NCCL_LIB)");
    return FB2_ERR_NCCL;
  }
  return FB2_OK;
}

#define FB2_NCCL_TRY(expr)                                                                    \
  do {                                                                                        \
    int _r = (expr);                                                                          \
    if (_r != kNcclSuccess) {                                                                 \
      fb2::set_error("%s failed: %s", #expr, nccl().GetErrorString ? nccl().GetErrorString(_r) : "?"); \
      return FB2_ERR_NCCL;                                                                    \
    }                                                                                         \
  } while (0)

// ------------------------------------------------------------------------ shard geometry
struct ShardGeom {
  size_t k_local = 0, n_local = 0;
  std::vector<size_t> cols;   // column-parallel: source columns, in shard order
};

int shard_geometry(size_t k, size_t n, size_t group, int split, int rank, int world, const size_t *segments,
                   int n_seg, ShardGeom *g) {
  if (world < 1 || rank < 0 || rank >= world) { set_error("tp_rank %d outside [0, %d)", rank, world); return FB2_ERR_ARG; }
  if (group == 0 || k % group || k % 8 || n % 8) { set_error("GPTQ shapes: K %% group, K %% 8, N %% 8 must be 0"); return FB2_ERR_SHAPE; }
  if (split == FB2_SPLIT_COLUMN) {
    std::vector<size_t> seg;
    if (segments && n_seg > 0) seg.assign(segments, segments + n_seg); else seg.push_back(n);
    size_t sum = 0;
    for (size_t w : seg) sum += w;
    if (sum != n) { set_error("column split: segment widths sum to %zu, N is %zu", sum, n); return FB2_ERR_SHAPE; }
    size_t off = 0;
    for (size_t w : seg) {
      if (w % (size_t)world || (w / world) % 8) {
        // qzeros packs 8 columns per word: a piece must start and end on a word boundary
        set_error("column split: segment %zu is not divisible into %d pieces of a multiple of 8 columns", w, world);
        return FB2_ERR_SHAPE;
      }
      const size_t s = w / world;
      if ((off + rank * s) % 8) { set_error("column split: piece does not start on a qzeros word"); return FB2_ERR_SHAPE; }
      for (size_t c = 0; c < s; c++) g->cols.push_back(off + rank * s + c);
      off += w;
    }
    g->k_local = k;
    g->n_local = g->cols.size();
    return FB2_OK;
  }
  if (split == FB2_SPLIT_ROW) {
    const size_t groups = k / group;
    if (groups % (size_t)world) { set_error("row split: %zu groups not divisible by tp=%d", groups, world); return FB2_ERR_SHAPE; }
    g->k_local = groups / world * group;
    g->n_local = n;
    return FB2_OK;
  }
  set_error("split must be FB2_SPLIT_COLUMN or FB2_SPLIT_ROW");
  return FB2_ERR_ARG;
}

inline uint32_t nibble(const int32_t *qw, size_t n, size_t row, size_t col) {
  return ((uint32_t)qw[(row >> 3) * n + col] >> (4 * (row & 7))) & 0xFu;
}

}  // namespace
}  // namespace fb2

using namespace fb2;

struct fb2_comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1, device = 0;
  // one-shot all-reduce over peer memory (fb2_comm_enable_p2p)
  bool p2p = false;
  size_t slot_elems = 0;
  uint8_t *ws = nullptr;                 // this rank's buffer: [data][flags][epoch][done]
  uint8_t *peer_ws[kMaxRanks] = {};      // IPC mappings (peer_ws[rank] == ws)
  size_t flags_off = 0, epoch_off = 0, done_off = 0;
};

extern "C" {

int fb2_gptq_shard_dims(size_t k, size_t n, size_t group_size, int split, int tp_rank, int tp_size,
                        const size_t *segments, int n_seg, size_t *k_local, size_t *n_local) {
  ShardGeom g;
  int rc = shard_geometry(k, n, group_size, split, tp_rank, tp_size, segments, n_seg, &g);
  if (rc) return rc;
  if (k_local) *k_local = g.k_local;
  if (n_local) *n_local = g.n_local;
  return FB2_OK;
}

int fb2_gptq_shard_host(const int32_t *qweight, const float *scales, const int32_t *qzeros, const int32_t *g_idx,
                        const float *bias, size_t group_size, size_t k, size_t n, int split, int tp_rank,
                        int tp_size, const size_t *segments, int n_seg, int32_t *qweight_out, float *scales_out,
                        int32_t *qzeros_out, float *bias_out, int32_t *x_cols_out) {
  if (!qweight || !scales || !qzeros || !qweight_out || !scales_out || !qzeros_out) {
    set_error("shard: NULL tensor pointer");
    return FB2_ERR_ARG;
  }
  ShardGeom g;
  int rc = shard_geometry(k, n, group_size, split, tp_rank, tp_size, segments, n_seg, &g);
  if (rc) return rc;
  const size_t groups = k / group_size;
  if (split == FB2_SPLIT_COLUMN) {
    const size_t nl = g.n_local;
    for (size_t r = 0; r < k / 8; r++)
      for (size_t c = 0; c < nl; c++) qweight_out[r * nl + c] = qweight[r * n + g.cols[c]];
    for (size_t gr = 0; gr < groups; gr++) {
      for (size_t c = 0; c < nl; c++) scales_out[gr * nl + c] = scales[gr * n + g.cols[c]];
      for (size_t c = 0; c < nl / 8; c++) qzeros_out[gr * (nl / 8) + c] = qzeros[gr * (n / 8) + g.cols[c * 8] / 8];
    }
    if (bias && bias_out)
      for (size_t c = 0; c < nl; c++) bias_out[c] = bias[g.cols[c]];
    if (x_cols_out)
      for (size_t i = 0; i < k; i++) x_cols_out[i] = (int32_t)i;   // consumes the full activation
    return FB2_OK;
  }
  // row-parallel
  const size_t gl = groups / (size_t)tp_size, g_lo = (size_t)tp_rank * gl, kl = g.k_local;
  bool trivial = g_idx == nullptr;
  if (g_idx) {
    trivial = true;
    for (size_t i = 0; i < k; i++) {
      if (g_idx[i] < 0 || (size_t)g_idx[i] >= groups) { set_error("g_idx[%zu] = %d outside [0, %zu)", i, g_idx[i], groups); return FB2_ERR_MODEL; }
      if ((size_t)g_idx[i] != i / group_size) trivial = false;
    }
  }
  std::vector<size_t> rows(kl);
  if (trivial) {
    for (size_t i = 0; i < kl; i++) rows[i] = g_lo * group_size + i;
  } else {
    // act-order groups must be balanced: exactly group_size rows each (native_safetensors.rs:1327-1372)
    std::vector<size_t> hist(groups, 0);
    for (size_t i = 0; i < k; i++) hist[(size_t)g_idx[i]]++;
    for (size_t gr = 0; gr < groups; gr++)
      if (hist[gr] != group_size) {
        set_error("row split: act-order group %zu has %zu rows, expected %zu", gr, hist[gr], group_size);
        return FB2_ERR_UNSUPPORTED;
      }
    // stable argsort of g_idx (quant.rs:693), then this rank's chunk of whole groups
    std::vector<size_t> perm(k);
    std::iota(perm.begin(), perm.end(), (size_t)0);
    std::stable_sort(perm.begin(), perm.end(), [&](size_t a, size_t b) { return g_idx[a] < g_idx[b]; });
    for (size_t i = 0; i < kl; i++) {
      rows[i] = perm[g_lo * group_size + i];
      if ((size_t)g_idx[rows[i]] != g_lo + i / group_size) {
        set_error("row split: act-order groups must be balanced (exactly %zu rows each)", group_size);
        return FB2_ERR_UNSUPPORTED;
      }
    }
  }
  for (size_t r = 0; r < kl / 8; r++)
    for (size_t c = 0; c < n; c++) {
      uint32_t word = 0;
      for (int i = 0; i < 8; i++) word |= nibble(qweight, n, rows[r * 8 + i], c) << (4 * i);
      qweight_out[r * n + c] = (int32_t)word;
    }
  memcpy(scales_out, scales + g_lo * n, gl * n * sizeof(float));
  memcpy(qzeros_out, qzeros + g_lo * (n / 8), gl * (n / 8) * sizeof(int32_t));
  if (bias_out) {
    // the bias is added once: by rank 0's partial
    for (size_t c = 0; c < n; c++) bias_out[c] = (bias && tp_rank == 0) ? bias[c] : 0.0f;
  }
  if (x_cols_out)
    for (size_t i = 0; i < kl; i++) x_cols_out[i] = (int32_t)rows[i];
  return FB2_OK;
}

int fb2_gptq_linear_create_sharded(const int32_t *qweight, const float *scales, const int32_t *qzeros,
                                   const int32_t *g_idx, const float *bias, uint32_t bits, size_t group_size,
                                   size_t k, size_t n, int device, int tp_rank, int tp_size, int split,
                                   const size_t *segments, int n_seg, int32_t *x_cols_out, fb2_linear_t **out) {
  if (!out) { set_error("out is NULL"); return FB2_ERR_ARG; }
  *out = nullptr;
  ShardGeom g;
  int rc = shard_geometry(k, n, group_size, split, tp_rank, tp_size, segments, n_seg, &g);
  if (rc) return rc;
  const size_t kl = g.k_local, nl = g.n_local, gl = kl / group_size;
  std::vector<int32_t> qw(kl / 8 * nl), qz(gl * (nl / 8));
  std::vector<float> sc(gl * nl), bs(bias ? nl : 0);
  rc = fb2_gptq_shard_host(qweight, scales, qzeros, g_idx, bias, group_size, k, n, split, tp_rank, tp_size, segments,
                           n_seg, qw.data(), sc.data(), qz.data(), bias ? bs.data() : nullptr, x_cols_out);
  if (rc) return rc;
  // column shards keep the layer's g_idx (all K rows are local); row shards are stored in sorted
  // order already, so their groups are sequential
  const int32_t *gi = split == FB2_SPLIT_COLUMN ? g_idx : nullptr;
  const bool has_bias = bias && (split == FB2_SPLIT_COLUMN || tp_rank == 0);
  return fb2_gptq_linear_create(qw.data(), sc.data(), qz.data(), gi, has_bias ? bs.data() : nullptr, bits, group_size,
                                kl, nl, device, out);
}

// ---------------------------------------------------------------------------- communicator
int fb2_comm_get_unique_id(void *id_out) {
  if (!id_out) { set_error("id_out is NULL"); return FB2_ERR_ARG; }
  int rc = nccl_ready();
  if (rc) return rc;
  NcclUniqueId id;
  FB2_NCCL_TRY(nccl().GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return FB2_OK;
}

int fb2_comm_init_rank(const void *unique_id, int rank, int world_size, int device, fb2_comm_t **out) {
  if (!out || !unique_id) { set_error("unique_id / out is NULL"); return FB2_ERR_ARG; }
  *out = nullptr;
  if (world_size < 1 || rank < 0 || rank >= world_size) { set_error("rank %d outside [0, %d)", rank, world_size); return FB2_ERR_ARG; }
  int rc = nccl_ready();
  if (rc) return rc;
  FB2_CUDA_TRY(cudaSetDevice(device));
  NcclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  fb2_comm *c = new fb2_comm();
  c->rank = rank; c->world = world_size; c->device = device;
  int r = nccl().CommInitRank(&c->comm, world_size, id, rank);
  if (r != kNcclSuccess) {
    set_error("ncclCommInitRank failed: %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
    delete c;
    return FB2_ERR_NCCL;
  }
  *out = c;
  return FB2_OK;
}

void fb2_comm_destroy(fb2_comm_t *comm) {
  if (!comm) return;
  if (comm->p2p) {
    cudaSetDevice(comm->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < comm->world; r++)
      if (r != comm->rank && comm->peer_ws[r]) cudaIpcCloseMemHandle(comm->peer_ws[r]);
    if (comm->ws) cudaFree(comm->ws);
  }
  if (comm->comm && nccl().ok) nccl().CommDestroy(comm->comm);
  delete comm;
}

int fb2_comm_world_size(const fb2_comm_t *comm) { return comm ? comm->world : 1; }
int fb2_comm_rank(const fb2_comm_t *comm) { return comm ? comm->rank : 0; }

int fb2_all_reduce_f16(fb2_comm_t *comm, fb2_stream_t stream, fb2_half *buf, size_t len) {
  if (!comm) { set_error("comm is NULL"); return FB2_ERR_ARG; }
  if (len == 0 || comm->world == 1) return FB2_OK;
  if (!buf) { set_error("buf is NULL"); return FB2_ERR_ARG; }
  FB2_NCCL_TRY(nccl().AllReduce(buf, buf, len, kNcclHalf, kNcclSum, comm->comm, (cudaStream_t)stream));
  return FB2_OK;
}

int fb2_all_gather_f16(fb2_comm_t *comm, fb2_stream_t stream, const fb2_half *send, fb2_half *recv,
                       size_t len_per_rank) {
  if (!comm) { set_error("comm is NULL"); return FB2_ERR_ARG; }
  if (len_per_rank == 0) return FB2_OK;
  if (!send || !recv) { set_error("send / recv is NULL"); return FB2_ERR_ARG; }
  if (comm->world == 1) {
    if (send != recv) FB2_CUDA_TRY(cudaMemcpyAsync(recv, send, len_per_rank * 2, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return FB2_OK;
  }
  FB2_NCCL_TRY(nccl().AllGather(send, recv, len_per_rank, kNcclHalf, comm->comm, (cudaStream_t)stream));
  return FB2_OK;
}

int fb2_broadcast_f16(fb2_comm_t *comm, fb2_stream_t stream, fb2_half *buf, size_t len, int root) {
  if (!comm) { set_error("comm is NULL"); return FB2_ERR_ARG; }
  if (len == 0 || comm->world == 1) return FB2_OK;
  if (!buf || root < 0 || root >= comm->world) { set_error("buf is NULL or root outside the communicator"); return FB2_ERR_ARG; }
  FB2_NCCL_TRY(nccl().Broadcast(buf, buf, len, kNcclHalf, root, comm->comm, (cudaStream_t)stream));
  return FB2_OK;
}

int fb2_all_to_all_f16(fb2_comm_t *comm, fb2_stream_t stream, const fb2_half *send, const size_t *send_offsets,
                       const size_t *send_counts, fb2_half *recv, const size_t *recv_offsets,
                       const size_t *recv_counts) {
  if (!comm) { set_error("comm is NULL"); return FB2_ERR_ARG; }
  if (!send || !recv || !send_offsets || !send_counts || !recv_offsets || !recv_counts) {
    set_error("all_to_all: NULL argument");
    return FB2_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (comm->world == 1) {
    const size_t cnt = std::min(send_counts[0], recv_counts[0]);
    if (cnt) FB2_CUDA_TRY(cudaMemcpyAsync(recv + recv_offsets[0], send + send_offsets[0], cnt * 2, cudaMemcpyDeviceToDevice, st));
    return FB2_OK;
  }
  FB2_NCCL_TRY(nccl().GroupStart());
  for (int peer = 0; peer < comm->world; peer++) {
    if (send_counts[peer])
      FB2_NCCL_TRY(nccl().Send(send + send_offsets[peer], send_counts[peer], kNcclHalf, peer, comm->comm, st));
    if (recv_counts[peer])
      FB2_NCCL_TRY(nccl().Recv(recv + recv_offsets[peer], recv_counts[peer], kNcclHalf, peer, comm->comm, st));
  }
  FB2_NCCL_TRY(nccl().GroupEnd());
  return FB2_OK;
}

// ----------------------------------------------------------------- expert-parallel MoE FFN
namespace {
struct EpLayout {
  size_t ids, wts, slot, send_meta, recv_meta, sorted, block_ids, total, send_x, recv_x, act, down_out, ret, bytes;
  size_t cap, rows, sorted_cap;
};
EpLayout ep_layout(size_t batch_loc, size_t top_k, size_t world, size_t e_loc, size_t hidden, size_t inter) {
  EpLayout L{};
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  L.cap = batch_loc * top_k;
  L.rows = world * L.cap;
  L.sorted_cap = L.rows + e_loc * 16;
  size_t o = 0;
  L.ids = o; o += al(L.cap * 4);
  L.wts = o; o += al(L.cap * 4);
  L.slot = o; o += al(L.cap * 4);
  L.send_meta = o; o += al(L.rows * 4);
  L.recv_meta = o; o += al(L.rows * 4);
  L.sorted = o; o += al(L.sorted_cap * 4);
  L.block_ids = o; o += al((L.sorted_cap / 16 + 1) * 4);
  L.total = o; o += 256;
  L.send_x = o; o += al(L.rows * hidden * 2);
  L.recv_x = o; o += al(L.rows * hidden * 2);
  L.act = o; o += al(L.rows * inter * 2);
  L.down_out = o; o += al(L.rows * hidden * 2);
  L.ret = o; o += al(L.rows * hidden * 2);
  L.bytes = o;
  return L;
}
}  // namespace

size_t fb2_moe_ep_workspace_bytes(const fb2_stack_t *gate_up, const fb2_stack_t *down, size_t max_batch_local,
                                  size_t top_k, int world_size) {
  if (!gate_up || !down || world_size < 1) return 0;
  return ep_layout(max_batch_local, top_k, (size_t)world_size, fb2_gptq_stack_num_experts(gate_up),
                   fb2_gptq_stack_k(gate_up), fb2_gptq_stack_n_per_expert(gate_up) / 2).bytes;
}

int fb2_moe_ep_ffn_forward(const fb2_stack_t *gate_up, const fb2_stack_t *down, fb2_comm_t *comm, fb2_stream_t stream,
                           const fb2_half *x, const fb2_half *router_logits, fb2_half *out, size_t batch_local,
                           size_t top_k, int norm_topk_prob, void *workspace, size_t workspace_bytes) {
  if (!gate_up || !down || !comm || !workspace) { set_error("NULL argument (stacks, comm, workspace)"); return FB2_ERR_ARG; }
  if (batch_local == 0) return FB2_OK;
  if (!x || !router_logits || !out) { set_error("x / router_logits / out is NULL"); return FB2_ERR_ARG; }
  const size_t world = (size_t)comm->world, e_loc = fb2_gptq_stack_num_experts(gate_up);
  const size_t hidden = fb2_gptq_stack_k(gate_up), inter = fb2_gptq_stack_n_per_expert(gate_up) / 2;
  if (fb2_gptq_stack_num_experts(down) != e_loc || fb2_gptq_stack_k(down) != inter || fb2_gptq_stack_n_per_expert(down) != hidden) {
    set_error("expert-parallel FFN: gate_up [%zu -> 2*%zu] and down stacks do not match", hidden, inter);
    return FB2_ERR_SHAPE;
  }
  const size_t E = e_loc * world;
  if (!ep_route_plan_supported((int)batch_local, (int)E, (int)top_k, (int)world) || hidden % 8) {
    set_error("expert-parallel FFN: needs batch_local * top_k <= 1024, <= 256 experts, hidden %% 8 == 0");
    return FB2_ERR_UNSUPPORTED;
  }
  const EpLayout L = ep_layout(batch_local, top_k, world, e_loc, hidden, inter);
  if (workspace_bytes < L.bytes) { set_error("workspace too small: %zu < %zu bytes", workspace_bytes, L.bytes); return FB2_ERR_ARG; }
  FB2_CUDA_TRY(cudaSetDevice(comm->device));
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t *ws = (uint8_t *)workspace;
  int32_t *ids = (int32_t *)(ws + L.ids), *slot = (int32_t *)(ws + L.slot), *send_meta = (int32_t *)(ws + L.send_meta);
  int32_t *recv_meta = (int32_t *)(ws + L.recv_meta), *sorted = (int32_t *)(ws + L.sorted);
  int32_t *block_ids = (int32_t *)(ws + L.block_ids), *total = (int32_t *)(ws + L.total);
  float *wts = (float *)(ws + L.wts);
  __half *send_x = (__half *)(ws + L.send_x), *recv_x = (__half *)(ws + L.recv_x), *act = (__half *)(ws + L.act);
  __half *down_out = (__half *)(ws + L.down_out), *ret = (__half *)(ws + L.ret);
  int rc;
  // 1. route this rank's tokens + dispatch plan (one launch), 2. rows into the send buffer
  if ((rc = launch_ep_route_plan((const __half *)router_logits, ids, wts, slot, send_meta, (int)batch_local, (int)E,
                                 (int)top_k, norm_topk_prob, (int)world, (int)L.cap, st))) return rc;
  if ((rc = launch_ep_scatter_rows((const __half *)x, slot, send_x, (int)L.cap, (int)hidden, (int)top_k, st))) return rc;
  // 3. dispatch: rows and their expert ids to the expert owners -- ONE grouped all-to-all
  const size_t xb = L.cap * hidden, mb = L.cap * 2;              // elements (halves) per peer
  if (world > 1) {
    rc = nccl_ready();
    if (rc) return rc;
    FB2_NCCL_TRY(nccl().GroupStart());
    for (size_t peer = 0; peer < world; peer++) {
      FB2_NCCL_TRY(nccl().Send(send_x + peer * xb, xb, kNcclHalf, (int)peer, comm->comm, st));
      FB2_NCCL_TRY(nccl().Recv(recv_x + peer * xb, xb, kNcclHalf, (int)peer, comm->comm, st));
      FB2_NCCL_TRY(nccl().Send((const __half *)send_meta + peer * mb, mb, kNcclHalf, (int)peer, comm->comm, st));
      FB2_NCCL_TRY(nccl().Recv((__half *)recv_meta + peer * mb, mb, kNcclHalf, (int)peer, comm->comm, st));
    }
    FB2_NCCL_TRY(nccl().GroupEnd());
  } else {
    recv_x = send_x;
    recv_meta = send_meta;
  }
  // 4. local grouped GEMMs over the received rows (row id == position in the receive buffer, top_k = 1)
  if ((rc = fb2_moe_align_block_size(stream, recv_meta, sorted, block_ids, total, L.rows, e_loc, 16, L.sorted_cap, 1))) return rc;
  if ((rc = fb2_gptq_stack_gemm_routed_silu(gate_up, stream, (const fb2_half *)recv_x, sorted, block_ids, total,
                                            (fb2_half *)act, L.rows, 16, 1, L.sorted_cap))) return rc;
  if ((rc = fb2_gptq_stack_gemm_routed(down, stream, (const fb2_half *)act, sorted, block_ids, total,
                                       (fb2_half *)down_out, L.rows, 16, 1, L.sorted_cap))) return rc;
  // 5. return to the token owners, 6. weighted combine over this rank's tokens
  if (world > 1) {
    FB2_NCCL_TRY(nccl().GroupStart());
    for (size_t peer = 0; peer < world; peer++) {
      FB2_NCCL_TRY(nccl().Send(down_out + peer * xb, xb, kNcclHalf, (int)peer, comm->comm, st));
      FB2_NCCL_TRY(nccl().Recv(ret + peer * xb, xb, kNcclHalf, (int)peer, comm->comm, st));
    }
    FB2_NCCL_TRY(nccl().GroupEnd());
  } else {
    ret = down_out;
  }
  return fb2_moe_combine(stream, (const fb2_half *)ret, slot, wts, out, batch_local, hidden, top_k);
}

int fb2_comm_enable_p2p(fb2_comm_t *comm, fb2_stream_t stream, size_t max_elems) {
  if (!comm) { set_error("comm is NULL"); return FB2_ERR_ARG; }
  if (comm->p2p) { set_error("peer memory is already enabled on this communicator"); return FB2_ERR_ARG; }
  if (comm->world > kMaxRanks) { set_error("peer-memory all-reduce supports up to %d ranks", kMaxRanks); return FB2_ERR_UNSUPPORTED; }
  if (max_elems == 0) { set_error("max_elems is 0"); return FB2_ERR_ARG; }
  if (comm->world == 1) return FB2_OK;
  int rc = nccl_ready();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  FB2_CUDA_TRY(cudaSetDevice(comm->device));
  const size_t slot = (max_elems + 7) / 8 * 8;
  const size_t data_bytes = 2 * (size_t)comm->world * slot * sizeof(__half);
  comm->flags_off = (data_bytes + 127) / 128 * 128;
  comm->epoch_off = comm->flags_off + 2 * (size_t)comm->world * sizeof(uint32_t);
  comm->done_off = comm->epoch_off + sizeof(uint32_t);
  const size_t total = comm->done_off + sizeof(uint32_t);
  FB2_CUDA_TRY(cudaMalloc((void **)&comm->ws, total));
  FB2_CUDA_TRY(cudaMemset(comm->ws, 0, total));
  // exchange the IPC handles through the communicator itself (64 bytes per rank)
  cudaIpcMemHandle_t mine;
  FB2_CUDA_TRY(cudaIpcGetMemHandle(&mine, comm->ws));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  uint8_t *d_h = nullptr;
  FB2_CUDA_TRY(cudaMalloc((void **)&d_h, 64 * (size_t)(comm->world + 1)));
  FB2_CUDA_TRY(cudaMemcpyAsync(d_h, &mine, 64, cudaMemcpyHostToDevice, st));
  FB2_NCCL_TRY(nccl().AllGather(d_h, d_h + 64, 32 /* halves */, kNcclHalf, comm->comm, st));
  std::vector<cudaIpcMemHandle_t> all(comm->world);
  FB2_CUDA_TRY(cudaMemcpyAsync(all.data(), d_h + 64, 64 * (size_t)comm->world, cudaMemcpyDeviceToHost, st));
  FB2_CUDA_TRY(cudaStreamSynchronize(st));
  FB2_CUDA_TRY(cudaFree(d_h));
  for (int r = 0; r < comm->world; r++) {
    if (r == comm->rank) { comm->peer_ws[r] = comm->ws; continue; }
    void *ptr = nullptr;
    FB2_CUDA_TRY(cudaIpcOpenMemHandle(&ptr, all[r], cudaIpcMemLazyEnablePeerAccess));
    comm->peer_ws[r] = (uint8_t *)ptr;
  }
  // nobody may push before every rank has zeroed its buffer
  FB2_NCCL_TRY(nccl().AllReduce(comm->ws + comm->done_off, comm->ws + comm->done_off, 2, kNcclHalf, kNcclSum, comm->comm, st));
  FB2_CUDA_TRY(cudaStreamSynchronize(st));
  comm->slot_elems = slot;
  comm->p2p = true;
  return FB2_OK;
}

int fb2_gptq_linear_forward_allreduce(const fb2_linear_t *lin, fb2_comm_t *comm, fb2_stream_t stream,
                                      const fb2_half *x, fb2_half *y, size_t m) {
  if (!comm) { set_error("comm is NULL"); return FB2_ERR_ARG; }
  const size_t elems = m * fb2_gptq_linear_out_features(lin);
  if (comm->p2p && comm->world > 1 && elems <= comm->slot_elems && elems % 8 == 0 && m > 0 && m <= 64) {
    // fused path: the GEMM epilogue pushes its tiles into every rank's receive slot over NVLink, a
    // small reduce kernel (PDL-chained) sums the world slots -- no NCCL launch on the critical path
    PushTargets pt{};
    for (int r = 0; r < comm->world; r++) {
      pt.data[r] = (__half *)comm->peer_ws[r];
      pt.flags[r] = (uint32_t *)(comm->peer_ws[r] + comm->flags_off);
    }
    pt.epoch = (const uint32_t *)(comm->ws + comm->epoch_off);
    pt.slot_elems = comm->slot_elems;
    pt.world = comm->world;
    pt.rank = comm->rank;
    uint32_t ctas = 0;
    int rc = forward_linear_push(lin, (cudaStream_t)stream, (const __half *)x, (__half *)y, m, &pt, &ctas);
    if (rc == FB2_OK) {
      PushExpect ex{};
      for (int r = 0; r < comm->world; r++) ex.ctas[r] = ctas;   // equal shards: every rank launches the same grid
      P2pLocal loc{(__half *)comm->ws, (uint32_t *)(comm->ws + comm->flags_off), (uint32_t *)(comm->ws + comm->epoch_off),
                   (uint32_t *)(comm->ws + comm->done_off), comm->slot_elems, comm->world};
      return launch_p2p_reduce(loc, ex, (__half *)y, elems, (cudaStream_t)stream);
    }
    if (rc != FB2_ERR_UNSUPPORTED) return rc;
  }
  const size_t n_out = fb2_gptq_linear_out_features(lin);
  int rc = fb2_gptq_linear_forward(lin, stream, x, y, m);
  if (rc) return rc;
  return fb2_all_reduce_f16(comm, stream, y, m * n_out);
}

}  // extern "C"
