"""Round-2 probes, to be run on a B200 with a probe build of the library:

    scripts/build_variant.sh rawnib "-DFB2_EXP_RAW_NIBBLES"      # 78 instructions per unit
    scripts/build_variant.sh nozero "-DFB2_EXP_NO_ZERO"          # 116 instructions per unit
    FB2_LIB=$PWD/ferrum-infer-rs_b200/csrc/build/variants/rawnib.so python scripts/rawnib_probe.py
    FB2_LIB=$PWD/ferrum-infer-rs_b200/csrc/build/variants/nozero.so python scripts/rawnib_probe.py

Answers two questions before the accumulator-side scale / zero epilogue is written ("Next (a)"):
 1. does tcgen05.mma kind::f16 honour fp16 SUBNORMAL A operands (q * 2^-24, q * 2^-20)?  The probe build writes the raw
    nibbles into the A operand, so y[n] must equal sum_k x[k] * q[n,k] * c(k), c = 2^-24 for k % 8 in {0,1,4,5} and
    2^-20 otherwise (dequant_word's register order).  All zeros = subnormals are flushed: use 0x6400 | q instead.
 2. how fast is the dequant loop at 78 instead of 146 instructions per unit?  (per-shape graph timing, M = 1)
The probe build is numerically wrong on purpose and is never the product."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

fb2 = importlib.import_module("ferrum-infer-rs_b200")
variant = os.environ.get("FB2_LIB", "")
assert "rawnib" in variant or "nozero" in variant, "run with FB2_LIB pointing at the rawnib or the nozero variant"

rng = np.random.default_rng(1)
k, n, m = 512, 256, 1
qw, sc, qz, _ = bench.synth_gptq(rng, k, n, act_order=False)
lin = fb2.GptqLinear.from_raw(qw, sc, qz, None, None, 4, 128, k, n)
x = rng.uniform(-1, 1, size=(m, k)).astype(np.float16)
y = lin.forward(torch.from_numpy(x).cuda()).float().cpu().numpy()
kk = np.arange(k)
q = ((qw.view(np.uint32)[kk // 8, :] >> ((kk % 8) * 4).astype(np.uint32)[:, None]) & 0xF).astype(np.float64)   # [k, n]
if "rawnib" in variant:
    c = np.where(np.isin(kk % 8, (0, 1, 4, 5)), 2.0 ** -24, 2.0 ** -20)
    want = (x.astype(np.float64) * c[None, :]) @ q
    err = np.abs(y - want).max() / np.abs(want).max()
    print(f"subnormal A operand: max|y|={np.abs(y).max():.3e} expected {np.abs(want).max():.3e} rel err {err:.2e} ->",
          "HONOURED" if err < 2e-3 else ("FLUSHED TO ZERO" if np.abs(y).max() == 0 else "MISMATCH (check the k order assumed here)"))
else:
    # nozero build: the A operand is f16_rn(q * s) (one fma, exact constants); the zero term is missing on purpose
    s_kn = np.repeat(np.asarray(sc, np.float32).reshape(k // 128, n), 128, axis=0)
    w = (q.astype(np.float32) * s_kn).astype(np.float16).astype(np.float64)
    want = x.astype(np.float64) @ w
    err = np.abs(y - want).max() / np.abs(want).max()
    print(f"q*s operand: rel err {err:.2e} ->", "AS DESIGNED" if err < 2e-3 else "MISMATCH")

out = {}
for (name, kk_, nn_) in bench.LLAMA["linears"]:
    lins = []
    for _ in range(8):
        a, b, c_, gi = bench.synth_gptq(rng, kk_, nn_, act_order=True)
        lins.append(fb2.GptqLinear.from_raw(a, b, c_, gi, None, 4, 128, kk_, nn_))
    xx = torch.randn(1, kk_, device="cuda").half()
    yy = torch.empty(1, nn_, device="cuda", dtype=torch.float16)

    def fn():
        for l in lins:
            l.forward(xx, yy, 1)
    g = bench.capture(fn)
    out[name] = round(bench.time_graph(g, 10, 3, 1) / 10 * 1e3 / len(lins), 2)
    del lins
print("us per GEMM at M=1 (probe build; product: qkv 8.83, o 7.83, gate_up 20.50, down 13.61):", out)
