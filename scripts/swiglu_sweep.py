"""Graph-timed dense gate|up: plain forward vs the SwiGLU-epilogue forward (same weights): python scripts/swiglu_sweep.py [M ...]"""
import importlib, os, sys, math
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fb2 = importlib.import_module("ferrum-infer-rs_b200")
Ms = [int(a) for a in sys.argv[1:]] or [1, 8]
rng = np.random.default_rng(0)
k, n = 4096, 28672
lins = []
for _ in range(8):
    qw, sc, qz, gi = bench.synth_gptq(rng, k, n, act_order=False)
    lins.append(fb2.GptqLinear.from_raw(qw, sc, qz, gi, None, 4, 128, k, n))
fb2.pregrow_scratch(0, 64)
out = {}
for m in Ms:
    x = torch.randn(m, k, device="cuda").half()
    y = torch.empty(m, n, device="cuda", dtype=torch.float16)
    y2 = torch.empty(m, n // 2, device="cuda", dtype=torch.float16)
    def f1():
        for l in lins: l.forward(x, y, m)
    def f2():
        for l in lins: l.forward_swiglu(x, y2, m)
    for nm, fn in (("plain", f1), ("swiglu", f2)):
        g = bench.capture(fn)
        t = bench.time_graph(g, 10, 3, 1) / 10
        out[f"{nm} M={m}"] = round(t * 1e3 / len(lins), 2)
print(out)
