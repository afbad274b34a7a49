"""Phase timing of the expert-parallel MoE FFN (torchrun --nproc-per-node N scripts/ep_breakdown.py [B])."""
import importlib, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
fb2 = importlib.import_module("ferrum-infer-rs_b200")
par = importlib.import_module("ferrum-infer-rs_b200.parallel")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H, I, E, K = 2048, 768, 128, 8
rng = np.random.default_rng(7)
e_loc = E // world
gu = [bench.synth_gptq(rng, H, 2 * I, False)[:3] for _ in range(e_loc)]
dn = [bench.synth_gptq(rng, I, H, False)[:3] for _ in range(e_loc)]
gst = fb2.GptqExpertStack.load_gptq_stacked([g[0] for g in gu], [g[1] for g in gu], [g[2] for g in gu], None, 4, 128, H, 2 * I, device=local)
dst = fb2.GptqExpertStack.load_gptq_stacked([d[0] for d in dn], [d[1] for d in dn], [d[2] for d in dn], None, 4, 128, I, H, device=local)
comm = par.Comm.from_torch_distributed(local)
ep = par.ExpertParallelFfn(gst, dst, K, True, B // world, E, rank, world, comm=comm)
fb2.pregrow_scratch(local, B * K)
x = torch.randn(B // world, H, device="cuda").clamp_(-1, 1).half()
lg = (torch.rand(B // world, E, device="cuda") * 4 - 2).half()
moe = fb2.moe
def phases():
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
    Bl, R, C = x.shape[0], world, ep.cap
    ev[0].record()
    ids, w = moe.route_topk_softmax(lg, K, True)
    flat = ids.reshape(-1)
    dest = torch.div(flat, e_loc, rounding_mode="floor").to(torch.int32)
    pbt, pti, offs = moe.moe_build_pairs_by_token(dest, R, K)
    slot = dest.long() * C + (pbt.long() - offs[:-1].long()[dest.long()])
    ep.send_meta.fill_(-1); ep.send_meta[slot] = (flat - dest * e_loc)
    tok = torch.arange(Bl * K, device=x.device) // K
    ep.send_x[slot] = x[tok]
    ev[1].record()
    comm.all_to_all_single(ep.recv_x, ep.send_x)
    comm.all_to_all_single(ep.recv_meta.view(torch.float16), ep.send_meta.view(torch.float16))
    ev[2].record()
    rows = R * C
    s, b, t = moe.moe_align_block_size(ep.recv_meta, e_loc, ep.block, None, pair_ids=True)
    ev[3].record()
    ep.gate_up.gemm_phase_vllm(ep.recv_x, s, b, t, ep.act, rows, ep.block, 1, silu=True)
    ep.down.gemm_phase_vllm(ep.act, s, b, t, ep.down_out, rows, ep.block, 1)
    ev[4].record()
    comm.all_to_all_single(ep.ret, ep.down_out)
    ev[5].record()
    out = torch.empty((Bl, H), dtype=torch.float16, device=x.device)
    moe.moe_combine(ep.ret, slot.to(torch.int32).contiguous(), w.reshape(-1).contiguous(), out, Bl, H, K)
    ev[6].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(6)]
for _ in range(5): phases()
acc = np.zeros(6)
n = 30
for _ in range(n): acc += np.array(phases())
ep.enable_native()
def timed(fn, n=50):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
t_native = timed(lambda: ep.forward(x, lg))
ep.native = False
t_torch = timed(lambda: ep.forward(x, lg))
if rank == 0:
    print(f"EP layer world={world} B={B} eager: fb2_moe_ep_ffn_forward {t_native:.1f} us, torch-level path {t_torch:.1f} us")
    names = ["route+prepare (torch ops)", "dispatch all-to-all x2", "align", "gate_up+down GEMMs", "return all-to-all", "combine"]
    print(f"EP world={world} B={B} (eager, us):", {nm: round(v / n, 1) for nm, v in zip(names, acc)}, "total", round(acc.sum() / n, 1))
dist.barrier()
os._exit(0)
