"""Pins the oracle against the second CPU reference the reference repository holds for this path: the Triton
W4A16 equivalence test's fixtures and its f32-accumulate host GEMM (triton_w4a16_eq.rs:30-116, 197-290)."""
import numpy as np
import pytest

from triton_eq_fixture import REAL, SMALL, build, cpu_reference_gemm, dequant_kn


@pytest.mark.parametrize("fx", [SMALL, REAL], ids=["small", "qkv_shape"])
def test_oracle_matches_triton_eq_reference(oracle, fx):
    m, k, n, g = fx["m"], fx["k"], fx["n"], fx["g"]
    x, qw, scales, qz = build(**fx)
    nibbles = (qz.view(np.uint32)[..., None] >> (np.arange(8, dtype=np.uint32) * np.uint32(4))) & np.uint32(0xF)
    assert nibbles.max() == 15                                                     # the fixture reaches zero = 16
    deq = dequant_kn(qw, scales, qz, k, n, g)
    w_nk = oracle.dequant_gptq(qw, scales, qz, None, 4, g, k, n)
    assert w_nk.shape == (n, k)
    assert np.array_equal(w_nk.view(np.uint32), np.ascontiguousarray(deq.T).view(np.uint32))      # bit-exact
    ref32 = cpu_reference_gemm(x, deq)                                             # the reference test's own host GEMM
    got = oracle.linear_forward(x, w_nk, None, m, threads=4)                       # cpu.rs:481-492, f64 accumulate
    assert np.isfinite(got).all()
    max_abs = float(np.abs(got - ref32).max())
    assert max_abs < fx["tol"], max_abs                                            # the reference's gate
    # the two accumulation orders differ by f32 round-off plus the final f16 rounding of the test's reference
    assert max_abs < 2.0 ** -10 * max(1.0, float(np.abs(ref32).max())), max_abs
