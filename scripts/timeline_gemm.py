"""In-kernel timeline of 8 back-to-back W4A16 GEMM launches replayed from a CUDA graph (needs a
-DFB2_TRACE=1 build).  usage: python scripts/timeline_gemm.py K N M [act_order]"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fb2 = importlib.import_module("ferrum-infer-rs_b200")
k, n, m = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
act = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
NL = 8
dbg = torch.zeros((296 + NL * 296) * 32, dtype=torch.int64).pin_memory()
os.environ["FB2_DEBUG_BUF"] = str(dbg.data_ptr())
rng = np.random.default_rng(0)
lins = []
for i in range(NL):
    qw, sc, qz, gi = bench.synth_gptq(rng, k, n, act_order=act)
    lins.append(fb2.GptqLinear.from_raw(qw, sc, qz, gi, None, 4, 128, k, n))
fb2.pregrow_scratch(0, max(m, 64))
x = torch.randn(m, k, device="cuda").half()
y = torch.empty(m, n, device="cuda", dtype=torch.float16)
def fn():
    for lin in lins: lin.forward(x, y, m)
g = bench.capture(fn)          # eager pass = 8 launches (seq 0..7), capture = 8 more (seq 0..7 again)
for _ in range(3): g.replay()
torch.cuda.synchronize()
dbg.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record()
torch.cuda.synchronize()
print(f"K={k} N={n} M={m} act={act}: graph of {NL} GEMMs {e0.elapsed_time(e1) * 1e3 / NL:.2f} us per GEMM (with trace stores)")
d = dbg.numpy()[296 * 32:].reshape(NL, 296, 32)
t0 = d[0, :, 0][d[0, :, 0] > 0].min()
names = ["entry", "setup", "gdc", "data0", "loop_end", "acc", "epi_end", "exit", "units", "fix_end"]
for L in range(NL):
    a = d[L]
    live = a[:, 0] > 0
    row = {}
    for i, nm in enumerate(names):
        v = a[live, i]
        v = v[v > 0] - t0
        row[nm] = (int(v.min()), int(np.median(v)), int(v.max())) if len(v) else None
    print(f"launch {L}: ctas {int(live.sum())} units/cta {int(a[live, 8].max())}  " +
          "  ".join(f"{nm}={row[nm][0]}/{row[nm][1]}/{row[nm][2]}" for nm in names if row[nm] and nm != "units"))

if os.environ.get("TL_PER_CTA"):
    a = d[4]
    dur = (a[:, 4] - a[:, 2]) / 1e3          # gdc -> loop_end of the last segment, us
    sm = a[:, 10]
    order = np.argsort(sm, kind="stable")
    print("per SM (smid: cta ids -> loop us):")
    for smid in sorted(set(sm.tolist()))[:40]:
        ids = [int(c) for c in np.where(sm == smid)[0]]
        print(f"  sm {int(smid):3d}: " + "  ".join(f"cta {c:3d} units {int(a[c, 8])} entry {int(a[c,0]-t0)/1e3:.1f} loop {dur[c]:.1f}" for c in ids))
    print("loop us by cta id (first 296):", np.round(dur, 1).tolist())
