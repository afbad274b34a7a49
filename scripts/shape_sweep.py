"""Graph-timed per-shape sweep: python scripts/shape_sweep.py [act_order 0/1] [M ...]"""
import importlib, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fb2 = importlib.import_module("ferrum-infer-rs_b200")
act = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True
Ms = [int(a) for a in sys.argv[2:]] or [1]
import math
rng = np.random.default_rng(0)
out = {}
for (name, k, n) in bench.LLAMA["linears"]:
    lins = []
    # enough distinct weights to defeat the 126 MB L2
    layers = min(48, max(4, math.ceil(400e6 / bench.alg_bytes(k, n, 1, act))))
    for _ in range(layers):
        qw, sc, qz, gi = bench.synth_gptq(rng, k, n, act_order=act)
        lins.append(fb2.GptqLinear.from_raw(qw, sc, qz, gi, None, 4, 128, k, n))
    fb2.pregrow_scratch(0, 64)
    for m in Ms:
        x = torch.randn(m, k, device="cuda").half(); y = torch.empty(m, n, device="cuda", dtype=torch.float16)
        def fn():
            for l in lins: l.forward(x, y, m)
        g = bench.capture(fn)
        t = bench.time_graph(g, 10, 3, 1) / 10
        b = bench.alg_bytes(k, n, m, act) * layers
        out[f"{name} M={m}"] = (round(t * 1e3 / layers, 2), round(b / (t * 1e-3) / 1e9))
    del lins
import subprocess
try:
    print("clocks:", subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
except Exception as e:
    print("clocks: n/a", e)
print(f"act={act} pdl={os.environ.get('FB2_DEBUG_NO_PDL','0')!='1'}", out)
