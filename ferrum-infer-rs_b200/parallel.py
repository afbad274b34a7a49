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


def all_reduce_(t):
    """Sum all-reduce of a partial [M, N] f16 output, in place (nccl_comm.rs:40-48)."""
    import torch.distributed as dist
    dist.all_reduce(t)
    return t
