"""The two fixtures of the reference's `crates/ferrum-kernels/tests/triton_w4a16_eq.rs` (asymmetric zero points that
include nibble 15 -> zero 16, f16 scales, no g_idx) restated in numpy, with that test's own CPU reference
(`cpu_reference_gemm`, :77-116: dequant in f32, f32 accumulate over k ascending, result rounded to f16).

SMALL = :18-21 (M=32, K=256, N=64, group 64, tolerance 5e-2 abs, :191-194);
REAL  = :24-27 (M=10, K=2048, N=2560, group 128: Qwen2.5-3B qkv_proj, tolerance 1.0 abs and no NaN / Inf, :340-361)."""
import numpy as np

SMALL = dict(m=32, k=256, n=64, g=64, a_amp=1.0, sc_base=0.01, sc_amp=0.005, tol=5e-2)
REAL = dict(m=10, k=2048, n=2560, g=128, a_amp=0.1, sc_base=0.005, sc_amp=0.002, tol=1.0)


def build(m, k, n, g, a_amp, sc_base, sc_amp, **_):
    """-> x [m,k] f32 (f16-representable), qweight [k/8,n] i32, scales [k/g,n] f32 (f16-representable),
    qzeros [k/g,n/8] i32   (build_gptq_tensors :30-74, build_real_gptq_tensors :197-233)"""
    f32 = np.float32
    i = np.arange(m * k, dtype=np.int64).astype(f32)
    x = (np.sin(i * f32(0.0017)).astype(f32) * f32(a_amp)).astype(np.float16).astype(f32).reshape(m, k)
    w_int = ((np.arange(k * n, dtype=np.int64) * 7 + 3) % 16).astype(np.uint32).reshape(k // 8, 8, n)
    qw = np.zeros((k // 8, n), dtype=np.uint32)
    for j in range(8):
        qw |= w_int[:, j, :] << np.uint32(4 * j)
    groups = k // g
    j = np.arange(groups * n, dtype=np.int64).astype(f32)
    scales = (f32(sc_base) + np.cos(j * f32(0.0007)).astype(f32) * f32(sc_amp)).astype(np.float16).astype(f32)
    z_int = ((np.arange(groups * n, dtype=np.int64) * 11 + 5) % 16).astype(np.uint32).reshape(groups, n // 8, 8)
    qz = np.zeros((groups, n // 8), dtype=np.uint32)
    for j8 in range(8):
        qz |= z_int[:, :, j8] << np.uint32(4 * j8)
    return x, qw.view(np.int32), scales.reshape(groups, n), qz.view(np.int32)


def dequant_kn(qw, scales, qz, k, n, g):
    """[k, n] f32, the test's formula: (nibble - (zero_nibble + 1)) as f32 * scale."""
    kk = np.arange(k)
    nib = (qw.view(np.uint32)[kk // 8, :] >> ((kk % 8) * 4).astype(np.uint32)[:, None]) & np.uint32(0xF)
    nn = np.arange(n)
    zero = ((qz.view(np.uint32)[:, nn // 8] >> ((nn % 8) * 4).astype(np.uint32)[None, :]) & np.uint32(0xF)).astype(np.int32) + 1
    grp = kk // g
    return (nib.astype(np.int32) - zero[grp, :]).astype(np.float32) * scales[grp, :]


def cpu_reference_gemm(x, deq_kn):
    """f32 accumulate, k ascending, rounded to f16 at the end (returned as f32 values)."""
    acc = np.zeros((x.shape[0], deq_kn.shape[1]), dtype=np.float32)
    for kk in range(x.shape[1]):
        acc += x[:, kk, None] * deq_kn[kk, None, :]
    return acc.astype(np.float16).astype(np.float32)
