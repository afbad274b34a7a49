"""Small driver for ncu: a handful of W4A16 forwards of one shape.
usage: python scripts/prof_gemm.py K N M [reps] [act_order]"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402  (synthetic weight generator)

fb2 = importlib.import_module("ferrum-infer-rs_b200")
k, n, m = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
act = bool(int(sys.argv[5])) if len(sys.argv) > 5 else True
rng = np.random.default_rng(0)
lins = []
for i in range(int(os.environ.get("PROF_LINEARS", "3"))):
    qw, sc, qz, gi = bench.synth_gptq(rng, k, n, act_order=act)
    lins.append(fb2.GptqLinear.from_raw(qw, sc, qz, gi, None, 4, 128, k, n))
fb2.pregrow_scratch(0, max(m, 64))
x = torch.randn(m, k, device="cuda").half()
y = torch.empty(m, n, device="cuda", dtype=torch.float16)
for r in range(reps):
    for lin in lins:
        lin.forward(x, y, m)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for r in range(20):
    for lin in lins:
        lin.forward(x, y, m)
e1.record()
torch.cuda.synchronize()
print(f"K={k} N={n} M={m} kernel={lins[0].kernel_name(m)} avg_us={e0.elapsed_time(e1) * 1e3 / (20 * len(lins)):.2f}")
