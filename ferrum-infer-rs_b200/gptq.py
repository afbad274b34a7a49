"""GPTQ linear / stacked-expert operators (mirror of crates/ferrum-quantization/src/gptq.rs
and crates/ferrum-kernels/src/marlin_expert_stack.rs) over the C ABI."""
from __future__ import annotations

import ctypes as C
import json
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _ffi
from ._ffi import Dispatch, check, lib


# ------------------------------------------------------------------------------- config
@dataclass
class QuantConfig:
    """crates/ferrum-quantization/src/config.rs:26-45 (serde names: `quant_method` with
    alias `method`; missing fields default)."""
    method: str = "none"
    bits: int = 0
    group_size: int = 0
    desc_act: bool = False
    sym: bool = False

    @classmethod
    def from_json(cls, text: str) -> "QuantConfig":
        d = json.loads(text)
        method = d.get("quant_method", d.get("method", "none"))
        if method not in ("none", "gptq", "awq", "gguf"):
            raise ValueError(f"unknown variant `{method}`")
        return cls(method=method, bits=int(d.get("bits", 0)), group_size=int(d.get("group_size", 0)),
                   desc_act=bool(d.get("desc_act", False)), sym=bool(d.get("sym", False)))


def canonicalize_gptq_qzeros_for_sym(qzeros: np.ndarray, sym: bool) -> np.ndarray:
    """native_safetensors.rs:1242-1246: sym=true checkpoints get every zero nibble forced to
    7 (zero point 8)."""
    if not sym:
        return qzeros
    out = np.empty_like(np.ascontiguousarray(qzeros, dtype=np.int32))
    out[...] = np.int32(0x77777777)
    return out


# ------------------------------------------------------------------------------ helpers
def _np_i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _np_f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_dev_f16(t: torch.Tensor, name: str, numel: int) -> None:
    if not (t.is_cuda and t.dtype == torch.float16 and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous CUDA float16 tensor")
    if t.numel() < numel:
        raise ValueError(f"{name} has {t.numel()} elements, needs {numel}")


def gemm_algorithmic_bytes(k: int, n: int, m: int, group_size: int = 128, act_order: bool = False) -> int:
    return int(lib().fb2_gemm_algorithmic_bytes(k, n, m, group_size, 1 if act_order else 0))


def gemm_flops(k: int, n: int, m: int) -> int:
    return int(lib().fb2_gemm_flops(k, n, m))


def pregrow_scratch(device: int, max_m: int) -> None:
    """BackendQuantMarlin::pregrow_marlin_gather_scratch (capabilities.rs:212)."""
    check(lib().fb2_pregrow_scratch(device, max_m))


# ------------------------------------------------------------------------------- linear
class _LinearBase:
    _h: Optional[int] = None

    def in_features(self) -> int:
        return int(lib().fb2_gptq_linear_in_features(self._h))

    def out_features(self) -> int:
        return int(lib().fb2_gptq_linear_out_features(self._h))

    @property
    def is_act_order(self) -> bool:
        return bool(lib().fb2_gptq_linear_is_act_order(self._h))

    def kernel_name(self, m: int) -> str:
        return lib().fb2_gptq_linear_kernel_name(self._h, m).decode()

    def forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, m: Optional[int] = None) -> torch.Tensor:
        """Linear::forward (linear.rs:123): appends the GEMM to the current stream.
        x [m, K] f16, out [m, N] f16 (allocated when None)."""
        k, n = self.in_features(), self.out_features()
        if m is None:
            m = x.numel() // k
        if out is None:
            out = torch.empty((m, n), dtype=torch.float16, device=x.device)
        if m == 0:
            return out
        _check_dev_f16(x, "input", m * k)
        _check_dev_f16(out, "out", m * n)
        check(lib().fb2_gptq_linear_forward(self._h, _stream(), x.data_ptr(), out.data_ptr(), m))
        return out

    __call__ = forward

    def act_order_perm(self) -> Optional[torch.Tensor]:
        """The layer's act-order column order as a device int32 tensor view (sorted position -> input column), or None.
        Feed it to fused_add_rms_norm(out_perm=...) and call forward_sorted / forward_swiglu(x_is_sorted=True)."""
        p = C.c_void_p()
        check(lib().fb2_gptq_linear_act_order_perm(self._h, C.byref(p)))
        return p.value

    def forward_sorted(self, x_sorted: torch.Tensor, out: Optional[torch.Tensor] = None, m: Optional[int] = None) -> torch.Tensor:
        """fb2_gptq_linear_forward_sorted: x already in this layer's act-order column order."""
        k, n = self.in_features(), self.out_features()
        if m is None:
            m = x_sorted.numel() // k
        if out is None:
            out = torch.empty((m, n), dtype=torch.float16, device=x_sorted.device)
        if m == 0:
            return out
        _check_dev_f16(x_sorted, "input", m * k)
        _check_dev_f16(out, "out", m * n)
        check(lib().fb2_gptq_linear_forward_sorted(self._h, _stream(), x_sorted.data_ptr(), out.data_ptr(), m))
        return out

    def forward_swiglu(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, m: Optional[int] = None,
                       x_is_sorted: bool = False) -> torch.Tensor:
        """fb2_gptq_linear_forward_swiglu: this gate|up layer with SiLU(gate) * up in the epilogue -> [m, N / 2]."""
        k, n = self.in_features(), self.out_features()
        if m is None:
            m = x.numel() // k
        if out is None:
            out = torch.empty((m, n // 2), dtype=torch.float16, device=x.device)
        if m == 0:
            return out
        _check_dev_f16(x, "input", m * k)
        _check_dev_f16(out, "out", m * (n // 2))
        check(lib().fb2_gptq_linear_forward_swiglu(self._h, _stream(), x.data_ptr(), out.data_ptr(), m, 1 if x_is_sorted else 0))
        return out

    def forward_host(self, x: np.ndarray, m: Optional[int] = None) -> np.ndarray:
        """The plugin call with HOST buffers (f16 numpy in, f16 numpy out); synchronous."""
        k, n = self.in_features(), self.out_features()
        x = np.ascontiguousarray(x, dtype=np.float16)
        if m is None:
            m = x.size // k
        y = np.empty((m, n), dtype=np.float16)
        check(lib().fb2_gptq_linear_forward_host(self._h, _ptr(x), _ptr(y), m))
        return y

    def dequantize(self, dtype=torch.float32) -> torch.Tensor:
        """[N, K] weights in disk k order (cpu_dequant_gptq / dequantize_gptq_with_g_idx)."""
        n, k = self.out_features(), self.in_features()
        w = torch.empty((n, k), dtype=dtype, device=f"cuda:{self.device}")
        fn = lib().fb2_gptq_linear_dequant_f32 if dtype == torch.float32 else lib().fb2_gptq_linear_dequant_f16
        check(fn(self._h, _stream(), w.data_ptr()))
        return w

    def close(self) -> None:
        if self._h is not None:
            lib().fb2_gptq_linear_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GptqLinear(_LinearBase):
    """crates/ferrum-quantization/src/gptq.rs:42-147."""

    def __init__(self, handle: int, device: int, metadata=None):
        self._h, self.device, self._metadata = handle, device, metadata

    @classmethod
    def from_raw(cls, qweight, scales, qzeros, g_idx, bias, bits: int, group_size: int,
                 in_features: int, out_features: int, device: int = 0) -> "GptqLinear":
        return cls.from_raw_with_metadata(qweight, scales, qzeros, g_idx, bias, bits, group_size,
                                          in_features, out_features, None, device)

    @classmethod
    def from_raw_with_metadata(cls, qweight, scales, qzeros, g_idx, bias, bits: int, group_size: int,
                               in_features: int, out_features: int, metadata=None,
                               device: int = 0) -> "GptqLinear":
        k, n = in_features, out_features
        qw, sc, qz = _np_i32(qweight).reshape(-1), _np_f32(scales).reshape(-1), _np_i32(qzeros).reshape(-1)
        gi = _np_i32(g_idx).reshape(-1) if g_idx is not None else None
        bs = _np_f32(bias).reshape(-1) if bias is not None else None
        if group_size > 0 and k % 8 == 0 and n % 8 == 0 and k % group_size == 0:
            g = k // group_size
            if qw.size != (k // 8) * n or sc.size != g * n or qz.size != g * (n // 8):
                raise _ffi.Fb2Error(-6, f"GPTQ tensor lengths do not match K={k} N={n} group_size={group_size}")
            if gi is not None and gi.size != k:
                raise _ffi.Fb2Error(-6, f"g_idx length {gi.size} != K {k}")   # native_safetensors.rs:1308
            if bs is not None and bs.size != n:
                raise _ffi.Fb2Error(-6, f"bias length {bs.size} != N {n}")
        h = C.c_void_p()
        check(lib().fb2_gptq_linear_create(_ptr(qw), _ptr(sc), _ptr(qz), _ptr(gi), _ptr(bs), bits,
                                           group_size, k, n, device, C.byref(h)))
        return cls(h.value, device, metadata)

    def metadata(self):
        return self._metadata


def fused_add_rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float, residual: Optional[torch.Tensor] = None,
                       out: Optional[torch.Tensor] = None, out_perm: Optional[int] = None) -> torch.Tensor:
    """fb2_fused_add_rms_norm: residual += x (in place, when given), out = rms_norm(.) * weight, optionally stored in an
    act-order linear's column order (out_perm = GptqLinear.act_order_perm())."""
    tokens, hidden = x.shape
    if out is None:
        out = torch.empty_like(x)
    check(lib().fb2_fused_add_rms_norm(_stream(), x.data_ptr(), residual.data_ptr() if residual is not None else None,
                                       weight.data_ptr(), out.data_ptr(), tokens, hidden, float(eps), out_perm))
    return out


def argmax_rows(logits: torch.Tensor) -> torch.Tensor:
    """fb2_argmax_rows_f16: [m, n] f16 -> [m] int32."""
    m, n = logits.shape
    out = torch.empty((m,), dtype=torch.int32, device=logits.device)
    check(lib().fb2_argmax_rows_f16(_stream(), logits.data_ptr(), m, n, out.data_ptr()))
    return out


def forward_host_batch(linears: Sequence[_LinearBase], xs: Sequence[np.ndarray], m: int,
                       outs: Optional[Sequence[np.ndarray]] = None) -> Sequence[np.ndarray]:
    """fb2_gptq_linear_forward_host_batch: host f16 arrays in, host f16 arrays out, one sync."""
    n = len(linears)
    xs = [np.ascontiguousarray(x, dtype=np.float16) for x in xs]
    if outs is None:
        outs = [np.empty((m, lin.out_features()), dtype=np.float16) for lin in linears]
    hs = (C.c_void_p * n)(*[lin._h for lin in linears])
    xp = (C.c_void_p * n)(*[x.ctypes.data for x in xs])
    yp = (C.c_void_p * n)(*[y.ctypes.data for y in outs])
    check(lib().fb2_gptq_linear_forward_host_batch(hs, n, xp, yp, m))
    return outs


class HostBatch:
    """Pre-bound argument arrays for repeated forward_host_batch calls (keeps ctypes marshalling
    out of the timed region)."""

    def __init__(self, linears: Sequence[_LinearBase], xs: Sequence[np.ndarray], outs: Sequence[np.ndarray], m):
        """m: rows of every entry, or one row count per entry (fb2_gptq_linear_forward_host_batch_m)."""
        n = len(linears)
        self._keep = (list(linears), list(xs), list(outs))
        self.n, self.m = n, m
        self.ms = None if isinstance(m, int) else (C.c_size_t * n)(*[int(v) for v in m])
        self.hs = (C.c_void_p * n)(*[lin._h for lin in linears])
        self.xp = (C.c_void_p * n)(*[x.ctypes.data for x in xs])
        self.yp = (C.c_void_p * n)(*[y.ctypes.data for y in outs])
        self.h2d_bytes = sum(int(x.nbytes) for x in xs)
        self.d2h_bytes = sum(int(y.nbytes) for y in outs)

    def run(self) -> None:
        if self.ms is None:
            check(lib().fb2_gptq_linear_forward_host_batch(self.hs, self.n, self.xp, self.yp, self.m))
        else:
            check(lib().fb2_gptq_linear_forward_host_batch_m(self.hs, self.n, self.xp, self.yp, self.ms))


# -------------------------------------------------------------------------------- stack
class GptqExpertStack:
    """MarlinExpertStack<B> (marlin_expert_stack.rs:31-109) built by load_gptq_stacked
    (capabilities.rs:180)."""

    def __init__(self, handle: int, device: int):
        self._h, self.device = handle, device

    @classmethod
    def load_gptq_stacked(cls, qweights: Sequence, scales: Sequence, qzeros: Sequence, g_idx, bits: int,
                          group_size: int, k: int, n_per_expert: int, device: int = 0) -> "GptqExpertStack":
        e = len(qweights)
        if not (len(scales) == e and len(qzeros) == e):
            raise _ffi.Fb2Error(-6, "load_gptq_stacked: per-expert list lengths differ")
        qw = [_np_i32(a).reshape(-1) for a in qweights]
        sc = [_np_f32(a).reshape(-1) for a in scales]
        qz = [_np_i32(a).reshape(-1) for a in qzeros]
        gi = _np_i32(g_idx).reshape(-1) if g_idx is not None else None
        if group_size > 0 and k % 8 == 0 and n_per_expert % 8 == 0 and k % group_size == 0:
            g = k // group_size
            for i in range(e):
                if qw[i].size != (k // 8) * n_per_expert or sc[i].size != g * n_per_expert or \
                        qz[i].size != g * (n_per_expert // 8):
                    raise _ffi.Fb2Error(-6, f"expert {i}: tensor lengths do not match K={k} n={n_per_expert}")
        arr = lambda lst: (C.c_void_p * max(e, 1))(*[a.ctypes.data for a in lst])  # noqa: E731
        h = C.c_void_p()
        check(lib().fb2_gptq_stack_create(arr(qw), arr(sc), arr(qz), _ptr(gi), bits, group_size, k,
                                          n_per_expert, e, device, C.byref(h)))
        return cls(h.value, device)

    def n_per_expert(self) -> int:
        return int(lib().fb2_gptq_stack_n_per_expert(self._h))

    def k(self) -> int:
        return int(lib().fb2_gptq_stack_k(self._h))

    def num_experts(self) -> int:
        return int(lib().fb2_gptq_stack_num_experts(self._h))

    def requires_vllm_moe(self) -> bool:
        return False

    def zero_workspace(self) -> None:
        check(lib().fb2_gptq_stack_zero_workspace(self._h, _stream()))

    def gemm_phase_batched(self, x: torch.Tensor, dispatches: Sequence[Tuple[int, int, int, int]],
                           output: torch.Tensor, k: int) -> None:
        """marlin_expert_stack.rs:64-70."""
        nd = len(dispatches)
        if nd == 0:
            return
        arr = (Dispatch * nd)(*[Dispatch(*map(int, d)) for d in dispatches])
        in_rows = max(d[1] + d[3] for d in dispatches)
        out_rows = max(d[2] + d[3] for d in dispatches)
        _check_dev_f16(x, "input", in_rows * k)
        _check_dev_f16(output, "output", out_rows * self.n_per_expert())
        check(lib().fb2_gptq_stack_gemm_batched(self._h, _stream(), x.data_ptr(), arr, nd,
                                                output.data_ptr(), k))

    def gemm_phase_vllm(self, x: torch.Tensor, sorted_token_ids: torch.Tensor, expert_ids: torch.Tensor,
                        num_tokens_past_padded: torch.Tensor, output: torch.Tensor, prob_m: int,
                        moe_block_size: int, top_k: int, silu: bool = False) -> None:
        """marlin_expert_stack.rs:82-96.  The output need not be pre-zeroed here."""
        for t, nm in ((sorted_token_ids, "sorted_token_ids"), (expert_ids, "expert_ids"),
                      (num_tokens_past_padded, "num_tokens_past_padded")):
            if not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
                raise ValueError(f"{nm} must be a contiguous CUDA int32 tensor")
        n_out = self.n_per_expert() // 2 if silu else self.n_per_expert()
        _check_dev_f16(x, "input", prob_m * self.k())
        _check_dev_f16(output, "output", prob_m * top_k * n_out)
        fn = lib().fb2_gptq_stack_gemm_routed_silu if silu else lib().fb2_gptq_stack_gemm_routed
        check(fn(self._h, _stream(), x.data_ptr(), sorted_token_ids.data_ptr(), expert_ids.data_ptr(),
                 num_tokens_past_padded.data_ptr(), output.data_ptr(), prob_m, moe_block_size, top_k,
                 sorted_token_ids.numel()))

    def make_expert_linear(self, expert_offset: int, expert_n: int, bias=None) -> "StackedExpertLinear":
        return StackedExpertLinear(self, expert_offset, expert_n, bias)

    def dequantize(self) -> torch.Tensor:
        """[E, n_per_expert, K] f32."""
        w = torch.empty((self.num_experts(), self.n_per_expert(), self.k()), dtype=torch.float32,
                        device=f"cuda:{self.device}")
        check(lib().fb2_gptq_stack_dequant_f32(self._h, _stream(), w.data_ptr()))
        return w

    def close(self) -> None:
        if self._h is not None:
            lib().fb2_gptq_stack_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StackedExpertLinear(_LinearBase):
    """crates/ferrum-quantization/src/gptq.rs:154-198: a Linear view of columns
    [expert_offset, expert_offset + expert_n) of a shared stack (kept alive by the view)."""

    def __init__(self, stack: GptqExpertStack, expert_offset: int, expert_n: int, bias=None):
        self._stack = stack
        self.device = stack.device
        bs = _np_f32(bias).reshape(-1) if bias is not None else None
        if bs is not None and bs.size != expert_n:
            raise _ffi.Fb2Error(-6, f"bias length {bs.size} != expert_n {expert_n}")
        h = C.c_void_p()
        check(lib().fb2_gptq_stack_make_expert_linear(stack._h, expert_offset, expert_n, _ptr(bs),
                                                      C.byref(h)))
        self._h = h.value

    @classmethod
    def new(cls, stack, expert_offset, expert_n):
        return cls(stack, expert_offset, expert_n)

    @classmethod
    def new_with_bias(cls, stack, expert_offset, expert_n, bias):
        return cls(stack, expert_offset, expert_n, bias)
