"""Graph-timed Qwen3-30B-A3B MoE FFN layer sweep: python scripts/moe_sweep.py [layers] [B ...]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fb2 = importlib.import_module("ferrum-infer-rs_b200")
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 6
Bs = [int(a) for a in sys.argv[2:]] or [1, 8, 32]
H, I, E, K = 2048, 768, 128, 8
rng = np.random.default_rng(4242)
ffns = []
for _ in range(layers):
    gu = [bench.synth_gptq(rng, H, 2 * I, False)[:3] for _ in range(E)]
    dn = [bench.synth_gptq(rng, I, H, False)[:3] for _ in range(E)]
    gst = fb2.GptqExpertStack.load_gptq_stacked([g[0] for g in gu], [g[1] for g in gu], [g[2] for g in gu], None, 4, 128, H, 2 * I)
    dst = fb2.GptqExpertStack.load_gptq_stacked([d[0] for d in dn], [d[1] for d in dn], [d[2] for d in dn], None, 4, 128, I, H)
    ffns.append(fb2.moe.MoeFfn(gst, dst, K, True, max_batch=max(Bs)))
fb2.pregrow_scratch(0, max(Bs) * K)
out = {}
for B in Bs:
    x = torch.randn(B, H, device="cuda").clamp_(-1, 1).half()
    lg = (torch.rand(B, E, device="cuda") * 4 - 2).half()
    y = torch.empty(B, H, device="cuda", dtype=torch.float16)
    def fn():
        for f in ffns: f.forward(x, lg, y)
    g = bench.capture(fn)
    t = bench.time_graph(g, 10, 3, 1) / 10
    out[B] = round(t * 1e3 / layers, 2)
print("us per MoE layer:", out)
