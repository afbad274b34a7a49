"""Multi-GPU parity worker (launched by torchrun from tests/test_multi_gpu.py, one rank per GPU):
TP row/column-parallel GPTQ linears + all-reduce, and the expert-parallel MoE FFN with all-to-all,
each against the single-GPU result of the same library (and transitively the CPU oracle)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

fb2 = importlib.import_module("ferrum-infer-rs_b200")
par = importlib.import_module("ferrum-infer-rs_b200.parallel")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = f"cuda:{local}"

    # ---------------- dense TP: gate_up (column) -> SiLU*up -> down (row) + all-reduce, act-order
    hidden, inter, g, m = 1024, 2048, 128, 5
    gu = oracle.make_synthetic(hidden, 2 * inter, g, 0x61, f16_scales=True)
    dn = oracle.make_synthetic(inter, hidden, g, 0x62, f16_scales=True)
    gi_gu = oracle.random_balanced_g_idx(hidden, g, 1)
    gi_dn = oracle.random_balanced_g_idx(inter, g, 2)
    x = oracle.f16_round(oracle.lcg_f32_array(3, m * hidden, -1.0, 1.0)).reshape(m, hidden)
    xd = torch.from_numpy(x.astype(np.float16)).to(dev)
    # single-GPU result
    l_gu = fb2.GptqLinear.from_raw(gu.qweight, gu.scales, gu.qzeros, gi_gu, None, 4, g, hidden, 2 * inter, device=local)
    l_dn = fb2.GptqLinear.from_raw(dn.qweight, dn.scales, dn.qzeros, gi_dn, None, 4, g, inter, hidden, device=local)
    h_full = l_gu.forward(xd)
    act_full = torch.empty((m, inter), dtype=torch.float16, device=dev)
    fb2.moe.fused_silu_mul_split(h_full, act_full, m, inter)
    y_full = l_dn.forward(act_full)
    # TP: the row-parallel shard's input columns define the column-parallel shard's output order
    t_dn = par.TensorParallelLinear("down_proj", dn.qweight, dn.scales, dn.qzeros, gi_dn, None, 4, g, inter, hidden,
                                    rank, world, device=local)
    cols = np.asarray(t_dn.x_cols)                     # columns of act (== gate/up channel ids) this rank consumes
    sel = np.concatenate([cols, inter + cols])         # permuted column-parallel gate||up slice
    qw_c = np.ascontiguousarray(gu.qweight.reshape(hidden // 8, 2 * inter)[:, sel])
    sc_c = np.ascontiguousarray(gu.scales.reshape(hidden // g, 2 * inter)[:, sel])
    # qzeros pack 8 columns per word: unpack / select / repack
    z = gu.qzeros.reshape(hidden // g, 2 * inter // 8).view(np.uint32)
    zn = ((z[:, :, None] >> (np.arange(8, dtype=np.uint32) * 4)) & 0xF).reshape(hidden // g, 2 * inter)[:, sel]
    qz_c = (zn.reshape(hidden // g, -1, 8).astype(np.uint32) << (np.arange(8, dtype=np.uint32) * 4)).sum(
        axis=2, dtype=np.uint32).view(np.int32)
    t_gu = fb2.GptqLinear.from_raw(qw_c, sc_c, qz_c, gi_gu, None, 4, g, hidden, len(sel), device=local)
    h_loc = t_gu.forward(xd)
    act_loc = torch.empty((m, len(cols)), dtype=torch.float16, device=dev)
    fb2.moe.fused_silu_mul_split(h_loc, act_loc, m, len(cols))
    y_tp = t_dn.forward(act_loc)
    torch.cuda.synchronize()
    rel = (y_tp.float() - y_full.float()).abs().max().item() / y_full.float().abs().max().item()
    assert rel < 2e-3, f"TP vs single-GPU rel {rel}"
    # and against the oracle
    w_gu = oracle.dequant_gptq(gu.qweight, gu.scales, gu.qzeros, gi_gu, 4, g, hidden, 2 * inter)
    w_dn = oracle.dequant_gptq(dn.qweight, dn.scales, dn.qzeros, gi_dn, 4, g, inter, hidden)
    ref = oracle.gemm(oracle.f16_round(oracle.silu_mul_split(oracle.f16_round(oracle.gemm(x, w_gu, m, 2 * inter, hidden)), m, inter)),
                      w_dn, m, hidden, inter)
    st = oracle.error_stats(y_tp.float().cpu().numpy(), ref)
    assert st["rel"] < 1e-2, st

    # ---------------- MoE EP: all-to-all dispatch / return vs the single-GPU fused FFN
    H, I, E, K, B = 512, 256, 8, 2, 4 * world
    gus = [oracle.make_synthetic(H, 2 * I, g, 0x700 + e, f16_scales=True) for e in range(E)]
    dns = [oracle.make_synthetic(I, H, g, 0x800 + e, f16_scales=True) for e in range(E)]
    xs = oracle.f16_round(oracle.lcg_f32_array(9, B * H, -1.0, 1.0)).reshape(B, H)
    lg = oracle.f16_round(oracle.lcg_f32_array(10, B * E, -2.0, 2.0)).reshape(B, E)
    mk = lambda parts, k, n: fb2.GptqExpertStack.load_gptq_stacked([p.qweight for p in parts], [p.scales for p in parts],
                                                                   [p.qzeros for p in parts], None, 4, g, k, n, device=local)
    full = fb2.moe.MoeFfn(mk(gus, H, 2 * I), mk(dns, I, H), K, True, max_batch=B)
    y_ref = full.forward(torch.from_numpy(xs.astype(np.float16)).to(dev), torch.from_numpy(lg.astype(np.float16)).to(dev))
    e_loc = E // world
    ep = par.ExpertParallelFfn(mk(gus[rank * e_loc:(rank + 1) * e_loc], H, 2 * I), mk(dns[rank * e_loc:(rank + 1) * e_loc], I, H),
                               K, True, B // world, E, rank, world)
    b_loc = B // world
    sl = slice(rank * b_loc, (rank + 1) * b_loc)
    y_ep = ep.forward(torch.from_numpy(xs[sl].astype(np.float16)).to(dev), torch.from_numpy(lg[sl].astype(np.float16)).to(dev))
    torch.cuda.synchronize()
    d = (y_ep.float() - y_ref[sl].float()).abs().max().item() / max(y_ref.float().abs().max().item(), 1e-6)
    assert d < 2e-3, f"EP vs single-GPU rel {d}"
    # ---------------- the library's own communicator + C++ sharding (csrc/parallel.cu)
    comm = par.Comm.from_torch_distributed(local)
    assert (comm.rank, comm.world) == (rank, world)
    t = torch.full((1000,), float(rank + 1), dtype=torch.float16, device=dev)
    comm.all_reduce_(t)
    torch.cuda.synchronize()
    assert torch.all(t == world * (world + 1) / 2), "fb2_all_reduce_f16"
    piece = torch.full((8,), float(rank), dtype=torch.float16, device=dev)
    gathered = torch.empty(8 * world, dtype=torch.float16, device=dev)
    comm.all_gather(piece, gathered)
    send = torch.arange(world * 4, dtype=torch.float16, device=dev) + 100 * rank
    recv = torch.empty_like(send)
    comm.all_to_all_single(recv, send)
    bc = torch.full((16,), float(rank + 7), dtype=torch.float16, device=dev)
    comm.broadcast_(bc, root=0)
    torch.cuda.synchronize()
    assert gathered.view(world, 8)[:, 0].tolist() == [float(r) for r in range(world)], "fb2_all_gather_f16"
    assert recv.view(world, 4)[:, 0].tolist() == [100.0 * r + 4 * rank for r in range(world)], "fb2_all_to_all_f16"
    assert torch.all(bc == 7.0), "fb2_broadcast_f16"
    # row-parallel down_proj created by fb2_gptq_linear_create_sharded, completed by
    # fb2_gptq_linear_forward_allreduce: must equal the torch.distributed path bit for bit
    import ctypes as C
    ffi = importlib.import_module("ferrum-infer-rs_b200._ffi")
    h = C.c_void_p()
    xcols = np.empty(inter // world, dtype=np.int32)
    keep = [np.ascontiguousarray(dn.qweight, dtype=np.int32), np.ascontiguousarray(dn.scales, dtype=np.float32),
            np.ascontiguousarray(dn.qzeros, dtype=np.int32), np.ascontiguousarray(gi_dn, dtype=np.int32)]
    ffi.check(ffi.lib().fb2_gptq_linear_create_sharded(keep[0].ctypes.data, keep[1].ctypes.data, keep[2].ctypes.data,
                                                        keep[3].ctypes.data, None, 4, g, inter, hidden, local, rank,
                                                        world, par.SPLIT_ROW, None, 0, xcols.ctypes.data, C.byref(h)))
    assert np.array_equal(xcols, np.asarray(t_dn.x_cols, dtype=np.int32))
    y_c = torch.empty((m, hidden), dtype=torch.float16, device=dev)
    ffi.check(ffi.lib().fb2_gptq_linear_forward_allreduce(h, comm._h, torch.cuda.current_stream().cuda_stream,
                                                           act_loc.data_ptr(), y_c.data_ptr(), m))
    torch.cuda.synchronize()
    assert torch.equal(y_c, y_tp), "fb2_gptq_linear_forward_allreduce != TensorParallelLinear + torch all_reduce"
    # expert-parallel FFN with the dispatch / return all-to-alls through fb2_all_to_all_f16
    ep_c = par.ExpertParallelFfn(mk(gus[rank * e_loc:(rank + 1) * e_loc], H, 2 * I), mk(dns[rank * e_loc:(rank + 1) * e_loc], I, H),
                                 K, True, B // world, E, rank, world, comm=comm)
    y_ep_c = ep_c.forward(torch.from_numpy(xs[sl].astype(np.float16)).to(dev), torch.from_numpy(lg[sl].astype(np.float16)).to(dev))
    torch.cuda.synchronize()
    assert torch.equal(y_ep_c, y_ep), "EP through fb2_all_to_all_f16 != EP through torch.distributed"
    # the same layer inside the library: fb2_moe_ep_ffn_forward (route+plan, scatter, grouped all-to-all,
    # align, two grouped GEMMs, all-to-all, combine)
    ep_n = par.ExpertParallelFfn(mk(gus[rank * e_loc:(rank + 1) * e_loc], H, 2 * I), mk(dns[rank * e_loc:(rank + 1) * e_loc], I, H),
                                 K, True, B // world, E, rank, world, comm=comm).enable_native()
    y_ep_n = ep_n.forward(torch.from_numpy(xs[sl].astype(np.float16)).to(dev), torch.from_numpy(lg[sl].astype(np.float16)).to(dev))
    torch.cuda.synchronize()
    dn_ = (y_ep_n.float() - y_ref[sl].float()).abs().max().item() / max(y_ref.float().abs().max().item(), 1e-6)
    assert dn_ < 2e-3, f"fb2_moe_ep_ffn_forward vs single-GPU rel {dn_}"
    assert torch.equal(y_ep_n, y_ep), "fb2_moe_ep_ffn_forward != the torch-level EP path"
    # ---------------- fused all-reduce: GEMM epilogue pushes its tiles to the peers (NVLink, IPC-mapped
    # buffers), one reduce kernel completes the sum.  Must match the NCCL path (same fp16 partials; the
    # sum is taken in fp32 in rank order) and be bit-identical under CUDA-graph replay.
    comm.enable_p2p(64 * hidden)
    y_p = torch.empty((m, hidden), dtype=torch.float16, device=dev)
    outs = []
    for rep in range(3):                                   # consecutive calls alternate the epoch parity
        y_p.zero_()
        ffi.check(ffi.lib().fb2_gptq_linear_forward_allreduce(h, comm._h, torch.cuda.current_stream().cuda_stream,
                                                               act_loc.data_ptr(), y_p.data_ptr(), m))
        torch.cuda.synchronize()
        outs.append(y_p.clone())
    relp = (outs[0].float() - y_full.float()).abs().max().item() / y_full.float().abs().max().item()
    assert relp < 2e-3, f"fused all-reduce vs single-GPU rel {relp}"
    assert all(torch.equal(outs[0], o) for o in outs[1:]), "fused all-reduce is not deterministic"
    # every rank must hold the same result
    gath = [torch.empty_like(y_p) for _ in range(world)]
    dist.all_gather(gath, outs[0])
    assert all(torch.equal(gath[0], g_) for g_ in gath[1:]), "ranks disagree after the fused all-reduce"
    fb2.pregrow_scratch(local, 64)
    g = torch.cuda.CUDAGraph()
    s_cap = torch.cuda.Stream()
    with torch.cuda.stream(s_cap):
        with torch.cuda.graph(g, stream=s_cap):
            ffi.check(ffi.lib().fb2_gptq_linear_forward_allreduce(h, comm._h, s_cap.cuda_stream, act_loc.data_ptr(),
                                                                   y_p.data_ptr(), m))
    for rep in range(5):
        y_p.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(y_p, outs[0]), "graph replay of the fused all-reduce differs"
    # timing: 200 calls each way (decode-sized output)
    def timed(fn):
        fn(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / 200
    st_ = torch.cuda.current_stream().cuda_stream
    t_fused = timed(lambda: ffi.check(ffi.lib().fb2_gptq_linear_forward_allreduce(h, comm._h, st_, act_loc.data_ptr(), y_p.data_ptr(), m)))
    t_nccl = timed(lambda: t_dn.forward(act_loc, y_p, m))
    if rank == 0:
        print(f"row-parallel down_proj K={inter}/{world} N={hidden} M={m}: fused push all-reduce {t_fused:.1f} us, GEMM + NCCL all-reduce {t_nccl:.1f} us")
    # prefill-sized rows through the same entry point (NCCL all-reduce; the fused push is for <= 64 rows)
    mb = 1100
    act_big = torch.from_numpy(oracle.f16_round(oracle.lcg_f32_array(31, mb * inter, -1.0, 1.0)).reshape(mb, inter).astype(np.float16)).to(dev)
    y_big_full = l_dn.forward(act_big)
    act_big_loc = act_big[:, torch.from_numpy(cols.astype(np.int64)).to(dev)].contiguous()
    y_big = torch.empty((mb, hidden), dtype=torch.float16, device=dev)
    fb2.pregrow_scratch(local, mb)
    ffi.check(ffi.lib().fb2_gptq_linear_forward_allreduce(h, comm._h, torch.cuda.current_stream().cuda_stream,
                                                           act_big_loc.data_ptr(), y_big.data_ptr(), mb))
    torch.cuda.synchronize()
    relb = (y_big.float() - y_big_full.float()).abs().max().item() / y_big_full.float().abs().max().item()
    assert relb < 2e-3, f"prefill-sized forward_allreduce vs single-GPU rel {relb}"
    ffi.lib().fb2_gptq_linear_destroy(h)
    comm.close()
    dist.barrier()
    if rank == 0:
        print(f"mgpu ok: world={world} tp_rel={rel:.2e} tp_vs_oracle={st['rel']:.2e} ep_rel={d:.2e} c_abi_comm=ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
