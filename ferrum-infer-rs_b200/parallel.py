"""Tensor-parallel sharding of GPTQ tensors and the collectives around the sharded GEMMs.

Shard definitions follow crates/ferrum-engine/src/parallel/tensor_parallel.rs:171-195 and
crates/ferrum-models/tests/tp_sharding_test.rs:7-170:
  column-parallel (qkv, gate_up): split N; every fused segment (q, k, v / gate, up) is split
      separately and the rank's pieces are re-fused, so SiLU(gate)*up and the head layout stay local;
  row-parallel (o, down): split K on 128-row group boundaries; each rank produces a partial
      [M, N] that is summed by one all-reduce (crates/ferrum-kernels/src/backend/cuda/nccl_comm.rs:40).
The pure-numpy shard functions run on CPU (tests use gloo, world_size 2); torch.distributed is
only the plumbing for the collective.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

COLUMN_PARALLEL = ("qkv_proj", "gate_up_proj", "q_proj", "k_proj", "v_proj", "gate_proj", "up_proj")
ROW_PARALLEL = ("o_proj", "down_proj")


def _unpack_rows(qweight: np.ndarray, k: int, n: int) -> np.ndarray:
    qw = np.ascontiguousarray(qweight, dtype=np.int32).reshape(k // 8, n).view(np.uint32)
    sh = (np.arange(8, dtype=np.uint32) * 4)[None, :, None]
    return ((qw[:, None, :] >> sh) & 0xF).reshape(k, n).astype(np.uint8)


def _pack_rows(nib: np.ndarray) -> np.ndarray:
    k, n = nib.shape
    v = nib.reshape(k // 8, 8, n).astype(np.uint32)
    out = np.zeros((k // 8, n), dtype=np.uint32)
    for bi in range(8):
        out |= v[:, bi, :] << np.uint32(bi * 4)
    return out.view(np.int32)


def column_shard(qweight, scales, qzeros, k: int, n: int, group: int, rank: int, world: int,
                 segments: Optional[Sequence[int]] = None):
    """Split N.  `segments` = widths of the fused parts along N (sum == n); each part is split into
    `world` equal pieces and the rank's pieces are concatenated.  Piece widths must be multiples of 8
    (qzeros packs 8 columns per word)."""
    segments = list(segments) if segments else [n]
    assert sum(segments) == n, (segments, n)
    qw = np.asarray(qweight, dtype=np.int32).reshape(k // 8, n)
    sc = np.asarray(scales, dtype=np.float32).reshape(k // group, n)
    qz = np.asarray(qzeros, dtype=np.int32).reshape(k // group, n // 8)
    cols, off = [], 0
    for w in segments:
        assert w % world == 0 and (w // world) % 8 == 0, f"segment {w} not divisible into {world} x 8-aligned pieces"
        s = w // world
        cols.append(np.arange(off + rank * s, off + (rank + 1) * s))
        off += w
    cols = np.concatenate(cols)
    zcols = cols.reshape(-1, 8)[:, 0] // 8
    return (np.ascontiguousarray(qw[:, cols]), np.ascontiguousarray(sc[:, cols]),
            np.ascontiguousarray(qz[:, zcols]), len(cols))


def row_shard(qweight, scales, qzeros, g_idx, k: int, n: int, group: int, rank: int, world: int):
    """Split K on group boundaries.  Returns (qweight, scales, qzeros, g_idx, k_local, x_cols):
    x_cols are the columns of the (full) input activation this shard multiplies, in the order the
    shard expects them.

    Sequential groups: rows [rank*K/R, (rank+1)*K/R), x_cols the same contiguous range.
    Act-order: the split is taken in g_idx-SORTED order, so that every rank owns whole groups and
    only its own scales/zeros; x_cols = perm[rank chunk].  The producer of the activation (a
    column-parallel layer, whose column-to-rank assignment is free) must emit exactly those columns
    -- see permuted_column_order()."""
    groups = k // group
    assert groups % world == 0, f"{groups} groups not divisible by tp={world}"
    gl = groups // world
    sc = np.asarray(scales, dtype=np.float32).reshape(groups, n)
    qz = np.asarray(qzeros, dtype=np.int32).reshape(groups, n // 8)
    g_lo, g_hi = rank * gl, (rank + 1) * gl
    trivial = g_idx is None or np.array_equal(np.asarray(g_idx), np.arange(k) // group)
    if trivial:
        qw = np.asarray(qweight, dtype=np.int32).reshape(k // 8, n)
        rows = np.arange(g_lo * group, g_hi * group)
        return (np.ascontiguousarray(qw[g_lo * group // 8: g_hi * group // 8]), np.ascontiguousarray(sc[g_lo:g_hi]),
                np.ascontiguousarray(qz[g_lo:g_hi]), None, gl * group, rows)
    gi = np.asarray(g_idx, dtype=np.int32)
    perm = np.argsort(gi, kind="stable")
    sel = perm[g_lo * group: g_hi * group]
    assert np.array_equal(gi[sel], np.repeat(np.arange(g_lo, g_hi), group)), "act-order groups must be balanced"
    nib = _unpack_rows(qweight, k, n)[sel]
    return (_pack_rows(nib), np.ascontiguousarray(sc[g_lo:g_hi]), np.ascontiguousarray(qz[g_lo:g_hi]),
            None, gl * group, sel)


def shard_gptq(name: str, qweight, scales, qzeros, g_idx, k: int, n: int, group: int, rank: int, world: int,
               qkv_segments: Optional[Sequence[int]] = None):
    """Shard one named projection -> (qweight, scales, qzeros, g_idx, k_local, n_local)."""
    if name in COLUMN_PARALLEL:
        if name == "qkv_proj":
            seg = list(qkv_segments) if qkv_segments else _default_qkv_segments(n)
        elif name == "gate_up_proj":
            seg = [n // 2, n // 2]
        else:
            seg = [n]
        qw, sc, qz, nl = column_shard(qweight, scales, qzeros, k, n, group, rank, world, seg)
        return qw, sc, qz, g_idx, k, nl
    if name in ROW_PARALLEL:
        qw, sc, qz, gi, kl, _ = row_shard(qweight, scales, qzeros, g_idx, k, n, group, rank, world)
        return qw, sc, qz, gi, kl, n
    raise ValueError(f"unknown projection {name}")


def _default_qkv_segments(n: int) -> List[int]:
    # Llama-3.1-8B 4096+1024+1024, Qwen3-30B-A3B 4096+512+512, Gemma3-27B 4096+2048+2048
    table = {6144: [4096, 1024, 1024], 5120: [4096, 512, 512], 8192: [4096, 2048, 2048]}
    return table.get(n, [n])


def all_reduce_(t, comm=None):
    """Sum all-reduce of a partial [M, N] f16 output, in place (nccl_comm.rs:40-48): through the
    library's own communicator when one is given, else torch.distributed."""
    if comm is not None:
        comm.all_reduce_(t)
        return t
    import torch.distributed as dist
    dist.all_reduce(t)
    return t


# ------------------------------------------------------------------------------------------
# The C-ABI side of the same logic (csrc/parallel.cu): host sharding in C++, NCCL communicator.
# ------------------------------------------------------------------------------------------
SPLIT_COLUMN, SPLIT_ROW = 0, 1


def shard_host(qweight, scales, qzeros, g_idx, bias, group: int, k: int, n: int, split: int, rank: int,
               world: int, segments: Optional[Sequence[int]] = None):
    """fb2_gptq_shard_host: the library's C++ statement of column_shard() / row_shard().
    Returns (qweight, scales, qzeros, bias, x_cols, k_local, n_local).  Pure host code."""
    import ctypes as C
    from ._ffi import lib, check
    qw = np.ascontiguousarray(qweight, dtype=np.int32)
    sc = np.ascontiguousarray(scales, dtype=np.float32)
    qz = np.ascontiguousarray(qzeros, dtype=np.int32)
    gi = None if g_idx is None else np.ascontiguousarray(g_idx, dtype=np.int32)
    bs = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
    seg = None if not segments else (C.c_size_t * len(segments))(*[int(v) for v in segments])
    kl, nl = C.c_size_t(0), C.c_size_t(0)
    check(lib().fb2_gptq_shard_dims(k, n, group, split, rank, world, seg, len(segments) if segments else 0,
                                    C.byref(kl), C.byref(nl)))
    kl, nl = kl.value, nl.value
    qw_o = np.empty((kl // 8, nl), dtype=np.int32)
    sc_o = np.empty((kl // group, nl), dtype=np.float32)
    qz_o = np.empty((kl // group, nl // 8), dtype=np.int32)
    bs_o = np.empty(nl, dtype=np.float32) if bs is not None else None
    xc_o = np.empty(kl, dtype=np.int32)
    ptr = lambda a: None if a is None else a.ctypes.data
    check(lib().fb2_gptq_shard_host(ptr(qw), ptr(sc), ptr(qz), ptr(gi), ptr(bs), group, k, n, split, rank, world,
                                    seg, len(segments) if segments else 0, ptr(qw_o), ptr(sc_o), ptr(qz_o),
                                    ptr(bs_o), ptr(xc_o)))
    return qw_o, sc_o, qz_o, bs_o, xc_o, kl, nl


class Comm:
    """fb2_comm_t: the library's NCCL communicator (BackendCollective, capabilities.rs:85-107).
    `Comm.from_torch_distributed()` distributes rank 0's unique id over an already initialised
    torch.distributed group (any backend) -- the id exchange is the only thing torch does here."""

    def __init__(self, unique_id: bytes, rank: int, world: int, device: int):
        import ctypes as C
        from ._ffi import lib, check
        assert len(unique_id) == 128
        self._h = C.c_void_p()
        buf = C.create_string_buffer(unique_id, 128)
        check(lib().fb2_comm_init_rank(buf, rank, world, device, C.byref(self._h)))
        self.rank, self.world, self.device = rank, world, device

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from ._ffi import lib, check
        buf = C.create_string_buffer(128)
        check(lib().fb2_comm_get_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int):
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        ids = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        torch.cuda.set_device(device)
        return cls(ids[0], rank, world, device)

    def _stream(self):
        import torch
        return torch.cuda.current_stream().cuda_stream

    def all_reduce_(self, t):
        from ._ffi import lib, check
        assert t.is_contiguous() and t.dtype.itemsize == 2
        check(lib().fb2_all_reduce_f16(self._h, self._stream(), t.data_ptr(), t.numel()))
        return t

    def all_gather(self, t, out):
        from ._ffi import lib, check
        assert t.is_contiguous() and out.is_contiguous() and out.numel() == t.numel() * self.world
        check(lib().fb2_all_gather_f16(self._h, self._stream(), t.data_ptr(), out.data_ptr(), t.numel()))
        return out

    def broadcast_(self, t, root=0):
        from ._ffi import lib, check
        check(lib().fb2_broadcast_f16(self._h, self._stream(), t.data_ptr(), t.numel(), root))
        return t

    def all_to_all_single(self, recv, send):
        """Equal split of `send` / `recv` (f16, contiguous) over the ranks."""
        import ctypes as C
        from ._ffi import lib, check
        per = send.numel() // self.world
        arr = (C.c_size_t * self.world)
        offs, cnts = arr(*[p * per for p in range(self.world)]), arr(*[per] * self.world)
        check(lib().fb2_all_to_all_f16(self._h, self._stream(), send.data_ptr(), offs, cnts, recv.data_ptr(), offs, cnts))
        return recv

    def enable_p2p(self, max_elems: int):
        """fb2_comm_enable_p2p: afterwards forward_allreduce() fuses the all-reduce of outputs of up to
        max_elems halves into the GEMM epilogue (peer-memory pushes + one reduce kernel)."""
        from ._ffi import lib, check
        check(lib().fb2_comm_enable_p2p(self._h, self._stream(), int(max_elems)))

    def forward_allreduce(self, linear, x, out, m):
        """fb2_gptq_linear_forward_allreduce on a row-parallel GptqLinear shard."""
        from ._ffi import lib, check
        check(lib().fb2_gptq_linear_forward_allreduce(linear._h, self._h, self._stream(), x.data_ptr(), out.data_ptr(), m))
        return out

    def close(self):
        from ._ffi import lib
        if self._h:
            lib().fb2_comm_destroy(self._h)
            self._h = None


# ------------------------------------------------------------------------------------------
# Tensor-parallel linear: a sharded GptqLinear + the collective that completes it.
# ------------------------------------------------------------------------------------------
class TensorParallelLinear:
    """One projection of a tensor-parallel layer (tensor_parallel.rs:171-195).

    column-parallel: forward(x_full [M, K]) -> local [M, N/R] slice (no exchange: it feeds the local
    row-parallel shard).  row-parallel: forward(x_local [M, K/R]) -> partial [M, N], then one sum
    all-reduce (nccl_comm.rs:40); `x_cols` tells which columns of the full activation this rank
    consumes (for act-order layers that is perm[rank chunk], see row_shard())."""

    def __init__(self, name, qweight, scales, qzeros, g_idx, bias, bits, group, k, n, rank, world, device=0,
                 qkv_segments=None):
        from .gptq import GptqLinear
        self.name, self.rank, self.world = name, rank, world
        self.row_parallel = name in ROW_PARALLEL
        self.x_cols = None
        if world == 1:
            qw, sc, qz, gi, kl, nl = qweight, scales, qzeros, g_idx, k, n
        elif self.row_parallel:
            qw, sc, qz, gi, kl, self.x_cols = row_shard(qweight, scales, qzeros, g_idx, k, n, group, rank, world)
            nl = n
            if bias is not None and rank != 0:
                bias = None                          # bias is added once, by rank 0's partial
        else:
            qw, sc, qz, gi, kl, nl = shard_gptq(name, qweight, scales, qzeros, g_idx, k, n, group, rank, world,
                                                qkv_segments)
            if bias is not None:
                seg = (list(qkv_segments) if qkv_segments else _default_qkv_segments(n)) if name == "qkv_proj" \
                    else ([n // 2, n // 2] if name == "gate_up_proj" else [n])
                cols, off = [], 0
                for w in seg:
                    s = w // world
                    cols.append(np.arange(off + rank * s, off + (rank + 1) * s))
                    off += w
                bias = np.asarray(bias, dtype=np.float32)[np.concatenate(cols)]
        self.k_local, self.n_local = kl, nl
        self.linear = GptqLinear.from_raw(qw, sc, qz, gi, bias, bits, group, kl, nl, device=device)

    def forward(self, x, out=None, m=None, comm=None):
        if self.row_parallel and self.world > 1 and comm is not None and out is not None:
            # library path (fb2_gptq_linear_forward_allreduce): fused push all-reduce for decode-sized
            # outputs, GEMM + ncclAllReduce through the library communicator otherwise
            return comm.forward_allreduce(self.linear, x, out, m if m is not None else x.shape[0])
        y = self.linear.forward(x, out, m)
        if self.row_parallel and self.world > 1:
            all_reduce_(y, comm)
        return y

    __call__ = forward


# ------------------------------------------------------------------------------------------
# Expert parallelism: experts partitioned contiguously over ranks, tokens batch-sharded, one
# all-to-all dispatch and one all-to-all return per MoE layer.  The reference has
# no EP (is_ep is hard-wired false, backend/cuda/quant.rs:1179); parity is defined against the
# single-GPU MoE FFN.
# ------------------------------------------------------------------------------------------
def ep_plan(ids: np.ndarray, num_experts: int, world: int, capacity: int):
    """Host (numpy) statement of the dispatch plan, used by the CPU tests: for every (token, k)
    pair the destination rank, its slot in the fixed-capacity send buffer [world, capacity] and
    the expert id local to the destination."""
    ids = np.asarray(ids, dtype=np.int64).reshape(-1)
    e_loc = num_experts // world
    dest = ids // e_loc
    local = ids - dest * e_loc
    slot = np.zeros_like(ids)
    counts = np.zeros(world, dtype=np.int64)
    for p, d in enumerate(dest):                      # (token, k) order == MoeBucketPlan order
        slot[p] = counts[d]
        counts[d] += 1
    assert counts.max(initial=0) <= capacity
    return dest, slot, local, counts


class ExpertParallelFfn:
    """MoE FFN with the experts of rank r = [r*E/R, (r+1)*E/R).  forward() takes this rank's tokens
    ([B_loc, H], router logits [B_loc, E]) and returns their outputs.  Fixed-capacity exchange
    buffers (capacity = B_loc * top_k pairs per peer) keep the whole layer free of host syncs."""

    def __init__(self, gate_up, down, top_k, norm_topk_prob, max_batch, num_experts, rank, world, comm=None):
        import torch
        self.comm = comm        # fb2 Comm: the f16 exchanges go through fb2_all_to_all_f16 instead of torch
        self.native = False     # set by enable_native(): run the layer through fb2_moe_ep_ffn_forward
        from . import moe
        self.moe, self.torch = moe, torch
        self.gate_up, self.down = gate_up, down
        self.top_k, self.norm, self.max_batch = top_k, norm_topk_prob, max_batch
        self.E, self.rank, self.world = num_experts, rank, world
        self.e_loc = num_experts // world
        assert gate_up.num_experts() == self.e_loc
        self.hidden, self.inter = gate_up.k(), gate_up.n_per_expert() // 2
        self.cap = max_batch * top_k
        dev = f"cuda:{gate_up.device}"
        rows = world * self.cap
        f16, i32 = torch.float16, torch.int32
        self.send_x = torch.zeros((rows, self.hidden), dtype=f16, device=dev)
        self.recv_x = torch.zeros((rows, self.hidden), dtype=f16, device=dev)
        self.send_meta = torch.full((rows,), -1, dtype=i32, device=dev)
        self.recv_meta = torch.full((rows,), -1, dtype=i32, device=dev)
        self.act = torch.empty((rows, self.inter), dtype=f16, device=dev)
        self.down_out = torch.zeros((rows, self.hidden), dtype=f16, device=dev)
        self.ret = torch.zeros((rows, self.hidden), dtype=f16, device=dev)
        self.block = 16

    def forward(self, x, router_logits, out=None):
        torch, moe = self.torch, self.moe
        if self.comm is not None and self.native:
            # fb2_moe_ep_ffn_forward: the whole layer inside the library (8 launches incl. 2 all-to-alls)
            from ._ffi import lib, check
            B = x.shape[0]
            if out is None:
                out = torch.empty((B, self.hidden), dtype=torch.float16, device=x.device)
            check(lib().fb2_moe_ep_ffn_forward(self.gate_up._h, self.down._h, self.comm._h,
                                               torch.cuda.current_stream().cuda_stream, x.data_ptr(),
                                               router_logits.data_ptr(), out.data_ptr(), B, self.top_k,
                                               1 if self.norm else 0, self.workspace.data_ptr(), self.workspace.numel()))
            return out
        import torch.distributed as dist
        B = x.shape[0]
        K, R, C = self.top_k, self.world, self.cap
        if out is None:
            out = torch.empty((B, self.hidden), dtype=torch.float16, device=x.device)
        ids, w = moe.route_topk_softmax(router_logits, K, self.norm)            # [B,K]
        flat = ids.reshape(-1)
        dest = torch.div(flat, self.e_loc, rounding_mode="floor").to(torch.int32)
        # counting sort of the pairs by destination rank (deterministic, (token,k) order inside a rank)
        pbt, pti, offs = moe.moe_build_pairs_by_token(dest, R, K)
        slot = dest.long() * C + (pbt.long() - offs[:-1].long()[dest.long()])   # position in [R, C]
        self.send_meta.fill_(-1)
        self.send_meta[slot] = (flat - dest * self.e_loc)
        tok = torch.arange(B * K, device=x.device) // K
        self.send_x[slot] = x[tok]
        if R > 1:
            if self.comm is not None:
                self.comm.all_to_all_single(self.recv_x, self.send_x)
                # int32 metadata travels as pairs of halves through the same f16 collective
                self.comm.all_to_all_single(self.recv_meta.view(torch.float16), self.send_meta.view(torch.float16))
            else:
                dist.all_to_all_single(self.recv_x, self.send_x)
                dist.all_to_all_single(self.recv_meta, self.send_meta)
            rx, rmeta = self.recv_x, self.recv_meta
        else:
            rx, rmeta = self.send_x, self.send_meta
        # local grouped GEMMs over the received rows (row id == position in the receive buffer)
        rows = R * C
        s, b, t = moe.moe_align_block_size(rmeta, self.e_loc, self.block, None, pair_ids=True)
        self.gate_up.gemm_phase_vllm(rx, s, b, t, self.act, rows, self.block, 1, silu=True)
        self.down.gemm_phase_vllm(self.act, s, b, t, self.down_out, rows, self.block, 1)
        if R > 1:
            if self.comm is not None:
                self.comm.all_to_all_single(self.ret, self.down_out)
            else:
                dist.all_to_all_single(self.ret, self.down_out)
            ret = self.ret
        else:
            ret = self.down_out
        moe.moe_combine(ret, slot.to(torch.int32).contiguous(), w.reshape(-1).contiguous(), out, B, self.hidden, K)
        return out

    def enable_native(self):
        """Run the layer through fb2_moe_ep_ffn_forward (needs `comm`)."""
        from ._ffi import lib
        assert self.comm is not None
        n = int(lib().fb2_moe_ep_workspace_bytes(self.gate_up._h, self.down._h, self.max_batch, self.top_k, self.world))
        self.workspace = self.torch.empty(n, dtype=self.torch.uint8, device=f"cuda:{self.gate_up.device}")
        self.native = True
        return self

    __call__ = forward

