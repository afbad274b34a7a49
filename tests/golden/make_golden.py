"""Writes tests/golden/gemm_small.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference (Rust) cannot be compiled in this image,
so these vectors pin the oracle port itself: dequantised weights and GEMM outputs of two small
fixtures built with the reference tests' generator (gptq_parity_test.rs:61-93) and seeds."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

out = {}
for tag, (k, n, g, m, seed) in {"seq": (256, 128, 128, 2, 0xDEADBEEF), "act": (512, 256, 128, 3, 0x5A170A7)}.items():
    syn = oracle.make_synthetic(k, n, g, seed, f16_scales=True)
    gi = oracle.make_desc_act_g_idx(k, g) if tag == "act" else None
    w = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, g, k, n)
    x = oracle.f16_round(oracle.lcg_f32_array(0xC0FFEE, m * k, -1.0, 1.0))
    out[f"{tag}_meta"] = np.array([k, n, g, m, seed], dtype=np.int64)
    out[f"{tag}_w"] = w
    out[f"{tag}_y"] = oracle.linear_forward(x, w, None, m)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gemm_small.npz"), **out)
print("wrote gemm_small.npz")
