"""CPU tests of the multi-GPU host logic: GPTQ tensor-parallel sharding identities
(crates/ferrum-models/tests/tp_sharding_test.rs:55-170 restated on QUANTISED tensors with the
oracle) and the expert-parallel dispatch plan, run as world_size-2 gloo process groups."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, world, port):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    par = importlib.import_module("ferrum-infer-rs_b200.parallel")
    return oracle, par


def _tp_worker(rank, world, port, act):
    oracle, par = _init(rank, world, port)
    k, n, g, m = 512, 256, 128, 3
    syn = oracle.make_synthetic(k, n, g, 0x7A9, f16_scales=True)
    gi = oracle.random_balanced_g_idx(k, g, 5) if act else None
    w_full = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, g, k, n)
    x = oracle.f16_round(oracle.lcg_f32_array(11, m * k, -1.0, 1.0)).reshape(m, k)
    ref = oracle.gemm(x, w_full, m, n, k)
    # ---- row-parallel: sum of partials == full (tp_sharding_test.rs:138-170)
    qw, sc, qz, gl, kl, cols = par.row_shard(syn.qweight, syn.scales, syn.qzeros, gi, k, n, g, rank, world)
    assert kl == k // world and gl is None
    w_r = oracle.dequant_gptq(qw, sc, qz, None, 4, g, kl, n)
    part = torch.from_numpy(oracle.gemm(np.ascontiguousarray(x[:, cols]), w_r, m, n, kl))
    dist.all_reduce(part)                                          # the collective under test
    assert np.abs(part.numpy() - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())
    # every disk row is owned by exactly one rank
    owned = [torch.zeros(kl, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(owned, torch.from_numpy(np.asarray(cols, dtype=np.int64)))
    assert sorted(torch.cat(owned).tolist()) == list(range(k))
    # ---- column-parallel gate_up: gate and up halves split separately (tp_sharding_test.rs:114-134)
    qw, sc, qz, nl = par.column_shard(syn.qweight, syn.scales, syn.qzeros, k, n, g, rank, world, [n // 2, n // 2])
    w_c = oracle.dequant_gptq(qw, sc, qz, gi, 4, g, k, nl)
    s = n // 2 // world
    expect = np.concatenate([w_full[rank * s:(rank + 1) * s], w_full[n // 2 + rank * s:n // 2 + (rank + 1) * s]])
    assert np.array_equal(w_c.view(np.uint32), expect.view(np.uint32))
    # ---- qkv: q / k / v segments split separately (tp_sharding_test.rs:7-52)
    qw, sc, qz, g2, k2, n2 = par.shard_gptq("qkv_proj", syn.qweight, syn.scales, syn.qzeros, gi, k, n, g, rank, world,
                                             qkv_segments=[128, 64, 64])
    w_q = oracle.dequant_gptq(qw, sc, qz, g2, 4, g, k2, n2)
    rows = np.concatenate([np.arange(rank * 64, (rank + 1) * 64), 128 + np.arange(rank * 32, (rank + 1) * 32),
                           192 + np.arange(rank * 32, (rank + 1) * 32)])
    assert np.array_equal(w_q.view(np.uint32), w_full[rows].view(np.uint32))
    dist.destroy_process_group()


def _ep_worker(rank, world, port):
    oracle, par = _init(rank, world, port)
    hidden, inter, E, top_k, b_loc, g = 64, 32, 4, 2, 3, 32
    gu = [oracle.make_synthetic(hidden, 2 * inter, g, 0xA + e) for e in range(E)]
    dn = [oracle.make_synthetic(inter, hidden, g, 0xB + e) for e in range(E)]
    gu_w = np.stack([oracle.dequant_gptq(s.qweight, s.scales, s.qzeros, None, 4, g, hidden, 2 * inter) for s in gu])
    dn_w = np.stack([oracle.dequant_gptq(s.qweight, s.scales, s.qzeros, None, 4, g, inter, hidden) for s in dn])
    x = oracle.lcg_f32_array(100 + rank, b_loc * hidden, -1.0, 1.0).reshape(b_loc, hidden)
    lg = oracle.lcg_f32_array(200 + rank, b_loc * E, -2.0, 2.0)
    ids, wts = oracle.route(lg, b_loc, E, top_k, True)
    ref = oracle.moe_forward_cpu(x, b_loc, hidden, inter, top_k, ids, wts, gu_w, dn_w)
    cap = b_loc * top_k
    dest, slot, local, counts = par.ep_plan(ids, E, world, cap)
    send_x = np.zeros((world, cap, hidden), np.float32)
    send_meta = np.full((world, cap), -1, np.int64)
    for p in range(b_loc * top_k):
        send_x[dest[p], slot[p]] = x[p // top_k]
        send_meta[dest[p], slot[p]] = local[p]
    # all-to-all == gather everybody's send buffers and keep the slice addressed to me
    allx = [torch.zeros(world, cap, hidden) for _ in range(world)]
    allm = [torch.zeros(world, cap, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allx, torch.from_numpy(send_x))
    dist.all_gather(allm, torch.from_numpy(send_meta))
    recv_x = np.stack([allx[src][rank].numpy() for src in range(world)])
    recv_m = np.stack([allm[src][rank].numpy() for src in range(world)])
    e_loc = E // world
    out_rows = np.zeros((world, cap, hidden), np.float32)
    for src in range(world):
        for s in range(cap):
            if recv_m[src, s] < 0:
                continue
            e = rank * e_loc + int(recv_m[src, s])
            gup = oracle.gemm(recv_x[src, s], gu_w[e], 1, 2 * inter, hidden)
            act = oracle.silu_mul_split(gup, 1, inter)
            out_rows[src, s] = oracle.gemm(act, dn_w[e], 1, hidden, inter)
    allo = [torch.zeros(world, cap, hidden) for _ in range(world)]
    dist.all_gather(allo, torch.from_numpy(out_rows))
    ret = np.stack([allo[d][rank].numpy() for d in range(world)])       # return all-to-all
    out = np.zeros((b_loc, hidden), np.float32)
    for p in range(b_loc * top_k):
        out[p // top_k] += wts[p] * ret[dest[p], slot[p]]
    assert np.abs(out - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())
    dist.destroy_process_group()


@pytest.mark.parametrize("act", [False, True])
def test_tp_sharding_world2_gloo(act):
    mp.spawn(_tp_worker, args=(2, _free_port(), act), nprocs=2, join=True)


def test_ep_dispatch_world2_gloo():
    mp.spawn(_ep_worker, args=(2, _free_port()), nprocs=2, join=True)


def test_row_shard_rejects_indivisible():
    par = importlib.import_module("ferrum-infer-rs_b200.parallel")
    with pytest.raises(AssertionError):
        par.row_shard(np.zeros((48, 8), np.int32), np.zeros((3, 8), np.float32), np.zeros((3, 1), np.int32), None,
                      384, 8, 128, 0, 2)


# ---------------------------------------------------------------------------------------------
# The library's own (C++) sharding, csrc/parallel.cu, against the numpy statement above.
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("act", [False, True])
def test_cxx_row_shard_matches_numpy(world, act):
    sys.path.insert(0, ROOT)
    import oracle
    par = importlib.import_module("ferrum-infer-rs_b200.parallel")
    k, n, g = 1024, 256, 128
    syn = oracle.make_synthetic(k, n, g, 0x51AB, f16_scales=True)
    gi = oracle.random_balanced_g_idx(k, g, 9) if act else None
    bias = oracle.lcg_f32_array(3, n, -1.0, 1.0)
    for rank in range(world):
        qw, sc, qz, _, kl, cols = par.row_shard(syn.qweight, syn.scales, syn.qzeros, gi, k, n, g, rank, world)
        cqw, csc, cqz, cb, ccols, ckl, cnl = par.shard_host(syn.qweight, syn.scales, syn.qzeros, gi, bias, g, k, n,
                                                            par.SPLIT_ROW, rank, world)
        assert (ckl, cnl) == (kl, n)
        assert np.array_equal(cqw, np.asarray(qw).reshape(kl // 8, n))
        assert np.array_equal(csc, np.asarray(sc).reshape(kl // g, n))
        assert np.array_equal(cqz, np.asarray(qz).reshape(kl // g, n // 8))
        assert np.array_equal(ccols, np.asarray(cols, dtype=np.int32))
        assert np.array_equal(cb, bias if rank == 0 else np.zeros(n, np.float32))   # bias stays with rank 0


@pytest.mark.parametrize("world", [2, 4])
def test_cxx_column_shard_matches_numpy(world):
    sys.path.insert(0, ROOT)
    import oracle
    par = importlib.import_module("ferrum-infer-rs_b200.parallel")
    k, g = 256, 128
    segments = [128, 64, 64]                       # a q|k|v-like fusion
    n = sum(segments)
    syn = oracle.make_synthetic(k, n, g, 0xC01, f16_scales=True)
    bias = oracle.lcg_f32_array(4, n, -1.0, 1.0)
    for rank in range(world):
        qw, sc, qz, nl = par.column_shard(syn.qweight, syn.scales, syn.qzeros, k, n, g, rank, world, segments)
        cqw, csc, cqz, cb, ccols, ckl, cnl = par.shard_host(syn.qweight, syn.scales, syn.qzeros, None, bias, g, k, n,
                                                            par.SPLIT_COLUMN, rank, world, segments)
        assert (ckl, cnl) == (k, nl)
        assert np.array_equal(cqw, qw) and np.array_equal(csc, sc) and np.array_equal(cqz, qz)
        cols, off = [], 0
        for w in segments:
            cols.append(np.arange(off + rank * (w // world), off + (rank + 1) * (w // world)))
            off += w
        assert np.array_equal(cb, bias[np.concatenate(cols)])
        assert np.array_equal(ccols, np.arange(k, dtype=np.int32))


def test_cxx_shard_rejects_bad_splits():
    sys.path.insert(0, ROOT)
    import oracle
    par = importlib.import_module("ferrum-infer-rs_b200.parallel")
    ffi = importlib.import_module("ferrum-infer-rs_b200._ffi")
    k, n, g = 384, 128, 128                        # 3 groups: not divisible by 2
    syn = oracle.make_synthetic(k, n, g, 1, f16_scales=True)
    with pytest.raises(ffi.Fb2Error) as e:
        par.shard_host(syn.qweight, syn.scales, syn.qzeros, None, None, g, k, n, par.SPLIT_ROW, 0, 2)
    assert e.value.code == -1                      # FB2_ERR_SHAPE
    with pytest.raises(ffi.Fb2Error):              # 24-column pieces are fine, 12-column ones are not
        par.shard_host(syn.qweight, syn.scales, syn.qzeros, None, None, g, k, n, par.SPLIT_COLUMN, 0, 2, [24, 104])
    gi = np.zeros(k, dtype=np.int32)               # unbalanced act-order groups
    with pytest.raises(ffi.Fb2Error) as e:
        par.shard_host(syn.qweight, syn.scales, syn.qzeros, gi, None, g, k + 0, n, par.SPLIT_ROW, 0, 3)
    assert e.value.code == -2                      # FB2_ERR_UNSUPPORTED
