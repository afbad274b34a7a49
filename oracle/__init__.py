"""CPU oracle for the GPTQ-Int4 W4A16 path -- TEST INFRASTRUCTURE ONLY.

ctypes front-end of ``oracle/liboracle.so`` (built from ``oracle/oracle.c``, a CPU
restatement of the reference's Rust CPU path; every C function cites the reference
file:line it follows) plus an independent numpy restatement of the unpack/dequant
(``np_dequant_gptq``) used to cross-check the C port the way the reference's own
``cpu_selfcheck`` test cross-checks two implementations
(crates/ferrum-quantization/tests/gptq_parity_test.rs:109-145).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product package
(``ferrum-infer-rs_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

c_i32p = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)
c_f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> str:
    """Compile oracle.c -> liboracle.so (gcc; a second or two)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_rnd_u32.restype = C.c_uint32
        _lib.oracle_rnd_f32.restype = C.c_float
        _lib.oracle_rnd_f32.argtypes = [C.POINTER(C.c_uint64), C.c_float, C.c_float]
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a, ty):
    return a.ctypes.data_as(ty) if a is not None else None


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def max_threads() -> int:
    return int(lib().oracle_max_threads())


# --------------------------------------------------------------------------- PRNG
class Lcg:
    """gptq_parity_test.rs:28-38 (pure python; for small draws and to pin the C LCG)."""

    MASK = (1 << 64) - 1

    def __init__(self, seed: int):
        self.state = seed & self.MASK

    def u32(self) -> int:
        self.state = (self.state * 6364136223846793005 + 1442695040888963407) & self.MASK
        return (self.state >> 33) & 0xFFFFFFFF

    def f32(self, lo: float, hi: float) -> np.float32:
        u = np.float32(self.u32() & 0x00FFFFFF) / np.float32(16777216.0)
        return np.float32(lo) + u * (np.float32(hi) - np.float32(lo))


def lcg_f32_array(seed: int, count: int, lo: float, hi: float) -> np.ndarray:
    """`count` successive rnd_f32 draws (C loop)."""
    out = np.empty(count, dtype=np.float32)
    st = C.c_uint64(seed)
    fn = lib().oracle_rnd_f32
    for i in range(count):
        out[i] = fn(C.byref(st), lo, hi)
    return out


class SyntheticGptq:
    """Mirror of the test fixture struct (gptq_parity_test.rs:40-48)."""

    def __init__(self, k, n, group_size, qweight, scales, qzeros):
        self.k, self.n, self.bits, self.group_size = k, n, 4, group_size
        self.qweight, self.scales, self.qzeros = qweight, scales, qzeros


def make_synthetic(k: int, n: int, group_size: int, seed: int, symmetric: bool = False,
                   f16_scales: bool = False) -> SyntheticGptq:
    """gptq_parity_test.rs:61-103.  f16_scales rounds scales through f16 (what an on-disk
    checkpoint holds, ferrum-quantization/src/gptq.rs:53) so GPU dequant can be bit-exact."""
    assert k % 8 == 0 and n % 8 == 0 and k % group_size == 0
    g = k // group_size
    qw = np.empty((k // 8, n), dtype=np.int32)
    sc = np.empty((g, n), dtype=np.float32)
    qz = np.empty((g, n // 8), dtype=np.int32)
    lib().oracle_make_synthetic(C.c_int64(k), C.c_int64(n), C.c_int64(group_size),
                                C.c_uint64(seed), C.c_int(1 if symmetric else 0),
                                _p(qw, c_i32p), _p(sc, c_f32p), _p(qz, c_i32p))
    if f16_scales:
        sc = sc.astype(np.float16).astype(np.float32)
    return SyntheticGptq(k, n, group_size, qw, sc, qz)


def make_desc_act_g_idx(k: int, group_size: int) -> np.ndarray:
    """gptq_parity_test.rs:50-59."""
    g = np.empty(k, dtype=np.int32)
    lib().oracle_make_desc_act_g_idx(C.c_int64(k), C.c_int64(group_size), _p(g, c_i32p))
    return g


def random_balanced_g_idx(k: int, group_size: int, seed: int) -> np.ndarray:
    """A seeded random balanced act-order look-alike (series iii)."""
    rng = np.random.default_rng(seed)
    g = np.repeat(np.arange(k // group_size, dtype=np.int32), group_size)
    rng.shuffle(g)
    return g


# ------------------------------------------------------------------------ dequant
def dequant_gptq(qweight, scales, qzeros, g_idx, bits, group_size, k, n) -> np.ndarray:
    """C port of cpu_dequant_gptq / dequantize_gptq_with_g_idx -> [n, k] f32."""
    qweight, scales, qzeros = _i32(qweight), _f32(scales), _i32(qzeros)
    gi = _i32(g_idx) if g_idx is not None else None
    w = np.empty((n, k), dtype=np.float32)
    rc = lib().oracle_dequant_gptq(_p(qweight, c_i32p), _p(scales, c_f32p), _p(qzeros, c_i32p),
                                   _p(gi, c_i32p), C.c_uint32(bits), C.c_int64(group_size),
                                   C.c_int64(k), C.c_int64(n), _p(w, c_f32p))
    if rc != 0:
        raise ValueError(f"CPU GPTQ: only bits=4 supported (got {bits})")
    return w


def np_dequant_gptq(qweight, scales, qzeros, g_idx, group_size, k, n) -> np.ndarray:
    """Independent vectorised numpy restatement of SURVEY.md A.1 -> [n, k] f32."""
    qw = np.asarray(qweight, dtype=np.int32).reshape(k // 8, n).view(np.uint32)
    shifts = (np.arange(8, dtype=np.uint32) * 4)[None, :, None]
    q = ((qw[:, None, :] >> shifts) & 0xF).reshape(k, n).astype(np.int32)      # [k, n]
    qz = np.asarray(qzeros, dtype=np.int32).reshape(-1, n // 8).view(np.uint32)
    zshift = (np.arange(8, dtype=np.uint32) * 4)[None, None, :]
    z = ((qz[:, :, None] >> zshift) & 0xF).reshape(qz.shape[0], n).astype(np.int32) + 1  # [G, n]
    grp = np.asarray(g_idx, dtype=np.int64) if g_idx is not None else np.arange(k) // group_size
    sc = np.asarray(scales, dtype=np.float32).reshape(-1, n)
    w_kn = (q - z[grp]).astype(np.float32) * sc[grp]
    return np.ascontiguousarray(w_kn.T)


# --------------------------------------------------------------------------- GEMM
def gemm(a, b_nk, m, n, k, threads: int = 1) -> np.ndarray:
    """CpuBackend::gemm (Linux branch): f64 accumulate, p ascending.  b is [n, k]."""
    a, b_nk = _f32(a), _f32(b_nk)
    assert a.size >= m * k and b_nk.size >= n * k
    out = np.empty((m, n), dtype=np.float32)
    lib().oracle_gemm(_p(a, c_f32p), _p(b_nk, c_f32p), _p(out, c_f32p), C.c_int64(m),
                      C.c_int64(n), C.c_int64(k), C.c_int(threads))
    return out


def linear_forward(x, w_nk, bias, m, threads: int = 1) -> np.ndarray:
    """CpuGptqLinear::forward: gemm (+ add_bias)."""
    n, k = w_nk.shape
    out = gemm(x, w_nk, m, n, k, threads)
    if bias is not None:
        b = _f32(bias)
        lib().oracle_add_bias(_p(out, c_f32p), _p(b, c_f32p), C.c_int64(m), C.c_int64(n))
    return out


def silu_mul_split(gate_up, tokens, im) -> np.ndarray:
    gu = _f32(gate_up)
    out = np.empty((tokens, im), dtype=np.float32)
    lib().oracle_silu_mul_split(_p(gu, c_f32p), _p(out, c_f32p), C.c_int64(tokens), C.c_int64(im))
    return out


def gelu_tanh_mul_split(gate_up, tokens, im) -> np.ndarray:
    gu = _f32(gate_up)
    out = np.empty((tokens, im), dtype=np.float32)
    lib().oracle_gelu_tanh_mul_split(_p(gu, c_f32p), _p(out, c_f32p), C.c_int64(tokens),
                                     C.c_int64(im))
    return out


# ---------------------------------------------------------------------------- MoE
def route(logits, batch, num_experts, top_k, norm_topk_prob):
    """route_into (router.rs:113-195).  Raises like the reference's assert!s."""
    lg = _f32(logits).reshape(-1)
    ids = np.zeros(max(batch * top_k, 0), dtype=np.uint32)
    wts = np.zeros(max(batch * top_k, 0), dtype=np.float32)
    rc = lib().oracle_route_into(_p(lg, c_f32p), C.c_int64(lg.size), C.c_int64(batch),
                                 C.c_int64(num_experts), C.c_int64(top_k),
                                 C.c_int(1 if norm_topk_prob else 0), _p(ids, c_u32p),
                                 _p(wts, c_f32p))
    if rc == -1:
        raise AssertionError("router logits shape mismatch")
    if rc == -2:
        raise AssertionError("top_k must be > 0")
    if rc == -3:
        raise AssertionError(f"top_k {top_k} exceeds num_experts {num_experts}")
    return ids, wts


def bucket_plan(expert_ids, batch, num_experts, top_k):
    """MoeBucketPlan::rebuild_into -> (expert_offsets[E+1], packed_token_idx, pairs_by_token)."""
    ids = np.ascontiguousarray(expert_ids, dtype=np.uint32)
    offs = np.zeros(num_experts + 1, dtype=np.int32)
    pti = np.zeros(batch * top_k, dtype=np.uint32)
    pbt = np.zeros(batch * top_k, dtype=np.int32)
    lib().oracle_bucket_plan(_p(ids, c_u32p), C.c_int64(batch), C.c_int64(num_experts),
                             C.c_int64(top_k), _p(offs, c_i32p), _p(pti, c_u32p), _p(pbt, c_i32p))
    return offs, pti, pbt


def moe_align_block_size(expert_ids_per_pair, num_experts, block_size, sorted_max_size,
                         pair_ids: bool):
    ids = _i32(expert_ids_per_pair)
    sorted_ids = np.zeros(sorted_max_size, dtype=np.int32)
    block_ids = np.full(sorted_max_size // block_size + 1, -1, dtype=np.int32)
    total = np.zeros(1, dtype=np.int32)
    lib().oracle_moe_align_block_size(_p(ids, c_i32p), C.c_int64(ids.size), C.c_int64(num_experts),
                                      C.c_int64(block_size), C.c_int64(sorted_max_size),
                                      C.c_int(1 if pair_ids else 0), _p(sorted_ids, c_i32p),
                                      _p(block_ids, c_i32p), _p(total, c_i32p))
    return sorted_ids, block_ids, int(total[0])


def moe_combine(packed_down, pairs_by_token, pair_weights, batch, hidden, top_k) -> np.ndarray:
    pd, pbt, pw = _f32(packed_down), _i32(pairs_by_token), _f32(pair_weights)
    out = np.empty((batch, hidden), dtype=np.float32)
    lib().oracle_moe_combine(_p(pd, c_f32p), _p(pbt, c_i32p), _p(pw, c_f32p), _p(out, c_f32p),
                             C.c_int64(batch), C.c_int64(hidden), C.c_int64(top_k))
    return out


def moe_forward_cpu(x, batch, hidden, inter, top_k, expert_ids, expert_weights, gate_up_w, down_w,
                    threads: int = 1) -> np.ndarray:
    """moe_forward_cpu (dispatch.rs:2208-2288).  gate_up_w [E,2I,H], down_w [E,H,I] f32."""
    x = _f32(x)
    if x.size != batch * hidden:
        raise ValueError(f"moe_forward_cpu: x len {x.size} doesn't match batch*hidden")
    ids = np.ascontiguousarray(expert_ids, dtype=np.uint32)
    if ids.size != batch * top_k:
        raise ValueError("moe_forward_cpu: router has wrong number of expert_ids")
    wts, gu, dn = _f32(expert_weights), _f32(gate_up_w), _f32(down_w)
    n_experts = gu.shape[0]
    out = np.empty((batch, hidden), dtype=np.float32)
    rc = lib().oracle_moe_forward_cpu(_p(x, c_f32p), C.c_int64(batch), C.c_int64(hidden),
                                      C.c_int64(inter), C.c_int64(top_k), _p(ids, c_u32p),
                                      _p(wts, c_f32p), C.c_int64(n_experts), _p(gu, c_f32p),
                                      _p(dn, c_f32p), _p(out, c_f32p), C.c_int(threads))
    if rc != 0:
        raise ValueError("moe_forward_cpu: router selected expert >= num_experts")
    return out


def validate_desc_act_g_idx(g_idx, k, group_size) -> int:
    g = _i32(g_idx)
    return int(lib().oracle_validate_desc_act_g_idx(_p(g, c_i32p), C.c_int64(k),
                                                    C.c_int64(group_size)))


# ------------------------------------------------------------------------ helpers
def f16_round(a) -> np.ndarray:
    """Round f32 -> f16 -> f32 (what the GPU sees; gptq_parity_test.rs:369-372)."""
    return np.asarray(a, dtype=np.float32).astype(np.float16).astype(np.float32)


def error_stats(got, ref):
    """max-abs, max-abs/max-ref (the reference's gate, gptq_parity_test.rs:385-404) and NMSE
    (ferrum-testkit/src/op_diff/mod.rs:24-25)."""
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1)
    d = np.abs(got - ref)
    max_abs = float(d.max()) if d.size else 0.0
    max_ref = max(float(np.abs(ref).max()) if ref.size else 0.0, 1e-6)
    nmse = float(((got - ref) ** 2).sum() / max((ref ** 2).sum(), 1e-30))
    return {"max_abs": max_abs, "rel": max_abs / max_ref, "nmse": nmse}
