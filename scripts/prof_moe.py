"""Small driver for ncu: a few Qwen3-30B-A3B MoE FFN forwards.  usage: python scripts/prof_moe.py B [reps]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fb2 = importlib.import_module("ferrum-infer-rs_b200")
B = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
H, I, E, K = 2048, 768, 128, 8
rng = np.random.default_rng(4242)
ffns = []
for _ in range(2):
    gu = [bench.synth_gptq(rng, H, 2 * I, False)[:3] for _ in range(E)]
    dn = [bench.synth_gptq(rng, I, H, False)[:3] for _ in range(E)]
    gst = fb2.GptqExpertStack.load_gptq_stacked([g[0] for g in gu], [g[1] for g in gu], [g[2] for g in gu], None, 4, 128, H, 2 * I)
    dst = fb2.GptqExpertStack.load_gptq_stacked([d[0] for d in dn], [d[1] for d in dn], [d[2] for d in dn], None, 4, 128, I, H)
    ffns.append(fb2.moe.MoeFfn(gst, dst, K, True, max_batch=B))
fb2.pregrow_scratch(0, B * K)
x = torch.randn(B, H, device="cuda").clamp_(-1, 1).half()
lg = (torch.rand(B, E, device="cuda") * 4 - 2).half()
y = torch.empty(B, H, device="cuda", dtype=torch.float16)
for _ in range(reps):
    for f in ffns: f.forward(x, lg, y)
torch.cuda.synchronize()
print("done", B)
