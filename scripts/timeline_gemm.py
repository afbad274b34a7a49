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
dbg = torch.zeros((296 + NL * 296) * 32, dtype=torch.int64, device="cuda")   # device memory: stamps to pinned host memory perturbed the epilogue by microseconds
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
d = dbg.cpu().numpy()[296 * 32:].reshape(NL, 296, 32)
t0 = d[0, :, 0][d[0, :, 0] > 0].min()
names = ["entry", "setup", "gdc", "data0", "loop_end", "acc", "epi_end", "exit", "units", "fix_end", "smid", "w_issue", "x_gdc", "x_land", "a_full", "rounds", "w_entry", "w_init", "w_seg"]
for L in range(NL):
    a = d[L]
    live = a[:, 0] > 0
    row = {}
    for i, nm in enumerate(names):
        v = a[live, i]
        v = v[v > 0] - t0
        row[nm] = (int(v.min()), int(np.median(v)), int(v.max())) if len(v) else None
    print(f"launch {L}: ctas {int(live.sum())} units/cta {int(a[live, 8].max())}  " +
          "  ".join(f"{nm}={row[nm][0]}/{row[nm][1]}/{row[nm][2]}" for nm in names if row[nm] and nm not in ("units", "smid", "rounds")))

if os.environ.get("TL_LATE"):
    a = d[4]
    live = a[:, 0] > 0
    base = a[live, 0].min()
    prev_exit = d[3][d[3][:, 0] > 0][:, 7].max()
    print(f"launch 4 relative to its first CTA entry (previous launch's last exit at {int(prev_exit - base)}):")
    order = np.argsort(a[:, 7])
    sel = [c for c in order if live[c]]
    show = sel[:4] + sel[len(sel) // 2 - 2: len(sel) // 2 + 2] + sel[-10:]
    for c in show:
        print(f"  cta {int(c):3d} sm {int(a[c,10]):3d} units {int(a[c,8]):2d}: " +
              " ".join(f"{nm}={int(a[c,i]-base) if a[c,i] > 0 else -1}" for i, nm in enumerate(names) if nm not in ("units", "smid", "rounds")) + f" rounds={int(a[c,15])}")
    c = sel[-1]
    print("the tile of the last CTA to exit:")
    for cc in range(max(0, int(c) - 1), min(296, int(c) + 10)):
        if live[cc]:
            print(f"  cta {cc:3d} units {int(a[cc,8]):2d}: entry={int(a[cc,0]-base)} data0={int(a[cc,3]-base)} x_land={int(a[cc,13]-base)} loop_end={int(a[cc,4]-base)} acc={int(a[cc,5]-base)} poll_end={int(a[cc,9]-base) if a[cc,9] > 0 else -1} epi_end={int(a[cc,6]-base)} rounds={int(a[cc,15])}")
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
