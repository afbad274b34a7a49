"""In-kernel timeline of the two grouped GEMMs of 4 Qwen3-30B-A3B MoE layers replayed from a CUDA graph (needs a
-DFB2_TRACE=1 build).  usage: python scripts/timeline_moe.py B"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fb2 = importlib.import_module("ferrum-infer-rs_b200")
B = int(sys.argv[1])
NL = 8
dbg = torch.zeros((296 + NL * 296) * 32, dtype=torch.int64, device="cuda")
os.environ["FB2_DEBUG_BUF"] = str(dbg.data_ptr())
H, I, E, K = 2048, 768, 128, 8
rng = np.random.default_rng(4242)
ffns = []
for _ in range(4):
    gu = [bench.synth_gptq(rng, H, 2 * I, False)[:3] for _ in range(E)]
    dn = [bench.synth_gptq(rng, I, H, False)[:3] for _ in range(E)]
    gst = fb2.GptqExpertStack.load_gptq_stacked([g[0] for g in gu], [g[1] for g in gu], [g[2] for g in gu], None, 4, 128, H, 2 * I)
    dst = fb2.GptqExpertStack.load_gptq_stacked([d[0] for d in dn], [d[1] for d in dn], [d[2] for d in dn], None, 4, 128, I, H)
    ffns.append(fb2.moe.MoeFfn(gst, dst, K, True, max_batch=B))
fb2.pregrow_scratch(0, B * K)
x = torch.randn(B, H, device="cuda").clamp_(-1, 1).half()
lg = (torch.rand(B, E, device="cuda") * 4 - 2).half()
y = torch.empty(B, H, device="cuda", dtype=torch.float16)
def fn():
    for f in ffns: f.forward(x, lg, y)
g = bench.capture(fn)
for _ in range(3): g.replay()
torch.cuda.synchronize()
dbg.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record()
torch.cuda.synchronize()
print(f"B={B}: {e0.elapsed_time(e1) * 1e3 / 4:.2f} us per MoE layer (with trace stores)")
d = dbg.cpu().numpy()[296 * 32:].reshape(NL, 296, 32)
t0 = d[0, :, 0][d[0, :, 0] > 0].min()
names = ["entry", "setup", "gdc", "data0", "loop_end", "acc", "epi_end", "exit", "units", "fix_end", "smid", "w_issue", "x_gdc", "x_land", "a_full"]
for L in range(NL):
    a = d[L]
    live = a[:, 0] > 0
    row = {}
    for i, nm in enumerate(names):
        v = a[live, i]
        v = v[v > 0] - t0
        row[nm] = (int(v.min()), int(np.median(v)), int(v.max())) if len(v) else None
    busy = a[live, 8] > 0
    print(f"launch {L} ({'gate_up' if L % 2 == 0 else 'down'}): ctas {int(live.sum())} working {int(busy.sum())} units/cta max {int(a[live, 8].max())}  " +
          "  ".join(f"{nm}={row[nm][0]}/{row[nm][1]}/{row[nm][2]}" for nm in names if row[nm] and nm not in ("units", "smid")))
# loop duration by CTA index / SM co-residency (launch 2 = a gate_up launch)
for L in (2, 3):
    a = d[L]
    live = (a[:, 0] > 0) & (a[:, 8] > 0)
    dur = (a[:, 4] - a[:, 3]) / 1e3
    idx = np.arange(296)
    lo, hi = live & (idx < 148), live & (idx >= 148)
    print(f"launch {L}: loop us (data0->loop_end) cta<148: median {np.median(dur[lo]):.1f} max {dur[lo].max():.1f} | cta>=148: median {np.median(dur[hi]) if hi.any() else 0:.1f} max {dur[hi].max() if hi.any() else 0:.1f}")
    print("   per-unit us by cta block of 37:", [round(float(np.median((dur / np.maximum(a[:, 8], 1))[live & (idx // 37 == b)])), 3) for b in range(8) if (live & (idx // 37 == b)).any()])
    sm = a[:, 10]
    pairs = {}
    for c in idx[live]:
        pairs.setdefault(int(sm[c]), []).append(c)
    both = [(v[0], v[1]) for v in pairs.values() if len(v) == 2]
    if both:
        d0 = np.array([dur[min(p)] for p in both]); d1 = np.array([dur[max(p)] for p in both])
        print(f"   SMs with two CTAs: {len(both)}; lower-index CTA median {np.median(d0):.1f}, higher-index CTA median {np.median(d1):.1f}, mean |diff| {np.mean(np.abs(d0-d1)):.1f}")
        ent0 = np.array([a[min(p), 0] for p in both]); ent1 = np.array([a[max(p), 0] for p in both])
        first = np.where(ent0 <= ent1, d0, d1); second = np.where(ent0 <= ent1, d1, d0)
        print(f"   first-entered CTA median {np.median(first):.1f}, second-entered {np.median(second):.1f}")
