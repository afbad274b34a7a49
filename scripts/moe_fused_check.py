"""Fused single-launch MoE FFN vs the oracle (small config) -- run under `timeout`.
usage: moe_fused_check.py [batch] [hidden inter E top_k]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
fb2 = importlib.import_module("ferrum-infer-rs_b200")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
hidden, inter, E, top_k = (int(a) for a in sys.argv[2:6]) if len(sys.argv) > 5 else (512, 256, 8, 2)
g = 128
def stack(k, n, seed):
    parts = [oracle.make_synthetic(k, n, g, seed + e, f16_scales=True) for e in range(E)]
    st = fb2.GptqExpertStack.load_gptq_stacked([p.qweight for p in parts], [p.scales for p in parts], [p.qzeros for p in parts], None, 4, g, k, n)
    w = np.stack([oracle.dequant_gptq(p.qweight, p.scales, p.qzeros, None, 4, g, k, n) for p in parts])
    return st, w
gu, guw = stack(hidden, 2 * inter, 1); dn, dnw = stack(inter, hidden, 100)
x = oracle.f16_round(oracle.lcg_f32_array(1, batch * hidden, -1, 1)).reshape(batch, hidden)
lg = oracle.f16_round(oracle.lcg_f32_array(2, batch * E, -2, 2)).reshape(batch, E)
ffn = fb2.moe.MoeFfn(gu, dn, top_k, True, max_batch=batch)
xd = torch.from_numpy(x.astype(np.float16)).cuda(); ld = torch.from_numpy(lg.astype(np.float16)).cuda()
if os.environ.get("FB2_TRACE_DUMP"):
    dbg = torch.zeros(296 * 32, dtype=torch.int64).pin_memory()
    os.environ["FB2_DEBUG_BUF"] = str(dbg.data_ptr())
    import atexit
    def dump():
        d = dbg.numpy().reshape(296, 32)
        t0 = d[:, 20][d[:, 20] > 0].min()
        names = ["entry", "routed", "bar1", "body1", "bar2", "body2", "bar3", "end"]
        for c in [0, 1, 147, 148, 295]:
            print("cta", c, "T,G,u0,u1", d[c, 16:20].tolist(), {n: int(d[c, 20 + i] - t0) for i, n in enumerate(names)})
        print("max over ctas (ns since first entry):", {n: int((d[:, 20 + i] - t0).max()) for i, n in enumerate(names)})
    atexit.register(dump)
fb2.reset_launch_count()
outs = []
if os.environ.get('PRESYNC'): torch.cuda.synchronize()
for rep in range(3):
    out = ffn.forward(xd, ld)
    torch.cuda.synchronize()
    outs.append(out.clone())
print("launches per forward:", fb2.launch_count() / 3)
ids, w = oracle.route(lg, batch, E, top_k, True)
ref = oracle.moe_forward_cpu(x, batch, hidden, inter, top_k, ids, w, guw, dnw)
print("vs oracle:", oracle.error_stats(outs[0].float().cpu().numpy(), ref))
print("replay bit-identical:", all(torch.equal(outs[0], o) for o in outs[1:]))
