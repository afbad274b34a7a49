"""Safetensors GPTQ checkpoint loader: host-side mirror of
crates/ferrum-quantization/src/native_safetensors.rs (NativeSafetensorsLoader) over the C ABI
(`fb2_ckpt_*`, csrc/checkpoint.cu).  The parsing, fusion and validation live in the C++ library;
this file only marshals."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from ._ffi import GptqInfoC, QuantConfigC, check, lib
from .gptq import GptqExpertStack, GptqLinear, QuantConfig

_METHODS = {0: "none", 1: "gptq", 2: "awq", 3: "gguf"}
_DTYPES = {-1: "other", 0: "F32", 1: "F16", 2: "BF16", 3: "I32"}
SPLIT_COLUMN, SPLIT_ROW = 0, 1


def _names(proj_names: Sequence[str]):
    return (C.c_char_p * len(proj_names))(*[p.encode() for p in proj_names])


class NativeSafetensorsLoader:
    """native_safetensors.rs:130-365, 588-760 (the GPTQ side of `impl WeightLoader`)."""

    def __init__(self, handle: int, model_dir: str):
        self._h, self.model_dir = handle, model_dir

    @classmethod
    def open(cls, model_dir: str) -> "NativeSafetensorsLoader":
        h = C.c_void_p()
        check(lib().fb2_ckpt_open(str(model_dir).encode(), C.byref(h)))
        return cls(h.value, str(model_dir))

    # ---- introspection
    def tensor_names(self):
        n = lib().fb2_ckpt_num_tensors(self._h)
        return [lib().fb2_ckpt_tensor_name(self._h, i).decode() for i in range(n)]

    def has_tensor(self, name: str) -> bool:
        return bool(lib().fb2_ckpt_has_tensor(self._h, name.encode()))

    def tensor_info(self, name: str) -> Tuple[str, Tuple[int, ...], int]:
        dt, rank, nbytes = C.c_int(), C.c_size_t(), C.c_size_t()
        shape = (C.c_size_t * 8)()
        check(lib().fb2_ckpt_tensor_info(self._h, name.encode(), C.byref(dt), C.byref(rank), shape, C.byref(nbytes)))
        return _DTYPES[dt.value], tuple(int(shape[i]) for i in range(rank.value)), int(nbytes.value)

    def quant_config(self) -> Optional[QuantConfig]:
        """quant_config_ref (:367)."""
        q = QuantConfigC()
        check(lib().fb2_ckpt_quant_config(self._h, C.byref(q)))
        if not q.present:
            return None
        return QuantConfig(method=_METHODS[q.method], bits=int(q.bits), group_size=int(q.group_size),
                           desc_act=bool(q.desc_act), sym=bool(q.sym))

    # ---- typed reads
    def read_f32(self, name: str) -> np.ndarray:
        _, shape, _ = self.tensor_info(name)
        out = np.empty(shape, dtype=np.float32)
        check(lib().fb2_ckpt_read_f32(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size))
        return out

    def read_i32(self, name: str) -> np.ndarray:
        _, shape, _ = self.tensor_info(name)
        out = np.empty(shape, dtype=np.int32)
        check(lib().fb2_ckpt_read_i32(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size))
        return out

    # ---- GPTQ linears
    def gptq_info(self, name: str) -> dict:
        info = GptqInfoC()
        check(lib().fb2_ckpt_gptq_info(self._h, name.encode(), C.byref(info)))
        return {"k": int(info.k), "n": int(info.n), "has_g_idx": bool(info.has_g_idx), "has_bias": bool(info.has_bias),
                "desc_act": bool(info.desc_act), "part_n": [int(info.part_n[i]) for i in range(info.n_parts)]}

    def read_gptq_raw(self, name: str):
        """read_gptq_raw (:342-365), extended to the fused names load_linear resolves:
        (qweight, scales, qzeros, g_idx | None, bias | None, k, n)."""
        info = self.gptq_info(name)
        k, n = info["k"], info["n"]
        g = k // self.quant_config().group_size
        qw = np.empty((k // 8, n), np.int32)
        sc = np.empty((g, n), np.float32)
        qz = np.empty((g, n // 8), np.int32)
        gi = np.empty((k,), np.int32) if info["has_g_idx"] else None
        bs = np.empty((n,), np.float32) if info["has_bias"] else None
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(lib().fb2_ckpt_read_gptq(self._h, name.encode(), p(qw), p(sc), p(qz), p(gi), p(bs)))
        return qw, sc, qz, gi, bs, k, n

    def load_linear(self, name: str, device: int = 0) -> GptqLinear:
        """WeightLoader::load_linear for a GPTQ layer (:597-634, :763-880, :887-1102)."""
        h = C.c_void_p()
        check(lib().fb2_ckpt_load_gptq_linear(self._h, name.encode(), device, C.byref(h)))
        return GptqLinear(h.value, device, {"name": name})

    def load_linear_sharded(self, name: str, tp_rank: int, tp_size: int, split: int, device: int = 0):
        """This rank's tensor-parallel shard; returns (linear, x_cols | None)."""
        info = self.gptq_info(name)
        x_cols = np.empty((info["k"],), np.int32) if split == SPLIT_ROW else None
        h = C.c_void_p()
        check(lib().fb2_ckpt_load_gptq_linear_sharded(
            self._h, name.encode(), device, tp_rank, tp_size, split,
            None if x_cols is None else x_cols.ctypes.data_as(C.c_void_p), C.byref(h)))
        lin = GptqLinear(h.value, device, {"name": name, "tp": (tp_rank, tp_size, split)})
        if x_cols is not None:
            x_cols = x_cols[:lin.in_features()]
        return lin, x_cols

    # ---- stacked experts
    def stacked_gptq_info(self, expert_prefix_fmt: str, proj_names: Sequence[str], expert: int = 0):
        k, n, g = C.c_size_t(), C.c_size_t(), C.c_int()
        check(lib().fb2_ckpt_stacked_gptq_info(self._h, expert_prefix_fmt.encode(), _names(proj_names), len(proj_names),
                                               expert, C.byref(k), C.byref(n), C.byref(g)))
        return int(k.value), int(n.value), bool(g.value)

    def read_stacked_gptq_expert(self, expert_prefix_fmt: str, proj_names: Sequence[str], expert: int):
        k, n, has_g = self.stacked_gptq_info(expert_prefix_fmt, proj_names, expert)
        g = k // self.quant_config().group_size
        qw = np.empty((k // 8, n), np.int32)
        sc = np.empty((g, n), np.float32)
        qz = np.empty((g, n // 8), np.int32)
        gi = np.empty((k,), np.int32) if has_g else None
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)  # noqa: E731
        check(lib().fb2_ckpt_read_stacked_gptq_expert(self._h, expert_prefix_fmt.encode(), _names(proj_names),
                                                      len(proj_names), expert, p(qw), p(sc), p(qz), p(gi)))
        return qw, sc, qz, gi

    def load_stacked_gptq_experts(self, expert_prefix_fmt: str, num_experts: int, proj_names: Sequence[str],
                                  device: int = 0, expert_begin: int = 0):
        """load_stacked_gptq_experts (:388-585): returns (stack, n_per_expert, k)."""
        h, n, k = C.c_void_p(), C.c_size_t(), C.c_size_t()
        check(lib().fb2_ckpt_load_stacked_gptq_experts(self._h, expert_prefix_fmt.encode(), expert_begin, num_experts,
                                                       _names(proj_names), len(proj_names), device, C.byref(h),
                                                       C.byref(n), C.byref(k)))
        return GptqExpertStack(h.value, device), int(n.value), int(k.value)

    def close(self) -> None:
        if self._h is not None:
            lib().fb2_ckpt_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
