"""GPU parity tests of the stacked-expert / MoE surface (MarlinExpertStack, BackendMoeFused,
moe_forward_bucketed) against the CPU oracle, on the reference tests' fixtures."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _stack(fb2, oracle, k, n_per, g, E, seed, act=None):
    parts = [oracle.make_synthetic(k, n_per, g, seed + e, f16_scales=True) for e in range(E)]
    gi = oracle.make_desc_act_g_idx(k, g) if act else None
    st = fb2.GptqExpertStack.load_gptq_stacked([p.qweight for p in parts], [p.scales for p in parts],
                                               [p.qzeros for p in parts], gi, 4, g, k, n_per)
    w = np.stack([oracle.dequant_gptq(p.qweight, p.scales, p.qzeros, gi, 4, g, k, n_per) for p in parts])
    return st, w, parts, gi


def _rel(oracle, got, ref):
    return oracle.error_stats(got, ref)["rel"]


def test_stack_dequant_and_views(fb2, oracle, cuda):
    """cuda_stacked_offset_vs_per_expert (gptq_parity_test.rs:712-805): k=256, n_e=128, E=4."""
    k, n_per, E, m = 256, 128, 4, 3
    st, w, parts, _ = _stack(fb2, oracle, k, n_per, 128, E, 0xA00DBA11)
    assert (st.num_experts(), st.n_per_expert(), st.k()) == (E, n_per, k)
    assert np.array_equal(st.dequantize().cpu().numpy().view(np.uint32), w.view(np.uint32))
    x = oracle.f16_round(np.sin(np.arange(m * k, dtype=np.float32) * np.float32(0.0017))).reshape(m, k)
    xd = cuda.from_numpy(x.astype(np.float16)).cuda()
    for e in range(E):
        view = fb2.StackedExpertLinear.new(st, e * n_per, n_per)
        y = view.forward(xd).float().cpu().numpy()
        ref = oracle.linear_forward(x, w[e], None, m)
        assert _rel(oracle, y, ref) < 1e-2
        alone = fb2.GptqLinear.from_raw(parts[e].qweight, parts[e].scales, parts[e].qzeros, None, None, 4, 128, k, n_per)
        assert cuda.equal(alone.forward(xd), view.forward(xd))        # stacked == per-expert, bit for bit
    with pytest.raises(fb2.Fb2Error, match="OOB"):
        st.make_expert_linear(E * n_per, n_per)


@pytest.mark.parametrize("act", [None, "mod"])
def test_gemm_phase_batched_fixture(fb2, oracle, cuda, act):
    """marlin_moe_parity_test.rs:80-200: k=256, n_e=256, E=4, tokens/expert [16,8,12,4],
    input sin(0.0027 i); fused vs per-expert rel 1e-3, vs CPU 1e-2."""
    k, n_per, E = 256, 256, 4
    counts = [16, 8, 12, 4]
    st, w, _, gi = _stack(fb2, oracle, k, n_per, 128, E, 0xCAFE0000, act)
    total = sum(counts)
    x = oracle.f16_round(np.sin(np.arange(total * k, dtype=np.float32) * np.float32(0.0027))).reshape(total, k)
    xd = cuda.from_numpy(x.astype(np.float16)).cuda()
    out = cuda.full((total, n_per), 7.0, dtype=cuda.float16, device="cuda")
    disp, off = [], 0
    for e, c in enumerate(counts):
        disp.append((e, off, off, c)); off += c
    disp.sort(key=lambda d: -d[3])                      # dispatch.rs:1967-1979 sorts by m desc
    st.gemm_phase_batched(xd, disp, out, k)
    got = out.float().cpu().numpy()
    off = 0
    for e, c in enumerate(counts):
        ref = oracle.gemm(x[off:off + c], w[e], c, n_per, k)
        assert _rel(oracle, got[off:off + c], ref) < 1e-2
        view = st.make_expert_linear(e * n_per, n_per)
        per = view.forward(xd[off:off + c].contiguous()).float().cpu().numpy()
        assert _rel(oracle, got[off:off + c], per) < 1e-3
        off += c


def test_gemm_phase_batched_large_m_and_offsets(fb2, oracle, cuda):
    """m > 64 per expert (split into row blocks) and in_row != out_row."""
    k, n_per, E = 512, 128, 3
    st, w, _, _ = _stack(fb2, oracle, k, n_per, 128, E, 0x77)
    x = oracle.f16_round(oracle.lcg_f32_array(3, 150 * k, -1.0, 1.0)).reshape(150, k)
    xd = cuda.from_numpy(x.astype(np.float16)).cuda()
    out = cuda.zeros((200, n_per), dtype=cuda.float16, device="cuda")
    disp = [(2, 0, 100, 100), (0, 100, 0, 50)]
    st.gemm_phase_batched(xd, disp, out, k)
    got = out.float().cpu().numpy()
    assert _rel(oracle, got[100:200], oracle.gemm(x[0:100], w[2], 100, n_per, k, threads=8)) < 1e-2
    assert _rel(oracle, got[0:50], oracle.gemm(x[100:150], w[0], 50, n_per, k, threads=8)) < 1e-2
    assert np.all(got[50:100] == 0)
    with pytest.raises(fb2.Fb2Error):
        st.gemm_phase_batched(xd, [(5, 0, 0, 1)], out, k)
    st.zero_workspace()


# ------------------------------------------------------------------ routing kernels
@pytest.mark.parametrize("batch,E,top_k,norm", [(1, 128, 8, True), (32, 128, 8, True), (5, 4, 2, True),
                                                  (7, 64, 6, False), (3, 8, 8, True)])
def test_router_matches_host(fb2, oracle, cuda, batch, E, top_k, norm):
    """GPU router vs route_into (router.rs:113-195) on f16-rounded logits: identical ids,
    weights within 1e-6 (expf differs by <= 1-2 ulp between libm and CUDA)."""
    lg = oracle.f16_round(oracle.lcg_f32_array(0xC0FFEE + batch, batch * E, -2.0, 2.0)).reshape(batch, E)
    ids, wts = fb2.moe.route_topk_softmax(cuda.from_numpy(lg.astype(np.float16)).cuda(), top_k, norm)
    rid, rw = oracle.route(lg, batch, E, top_k, norm)
    assert np.array_equal(ids.cpu().numpy().reshape(-1), rid.astype(np.int32))
    assert np.abs(wts.cpu().numpy().reshape(-1) - rw).max() < 1e-6


def test_router_ties_and_errors(fb2, oracle, cuda):
    """moe_router_test.rs:94-119."""
    lg = cuda.ones((1, 8), dtype=cuda.float16, device="cuda")
    ids, _ = fb2.moe.route_topk_softmax(lg, 2, True)
    assert ids.cpu().numpy().tolist() == [[0, 1]]
    with pytest.raises(fb2.Fb2Error, match="top_k must be > 0"):
        fb2.moe.route_topk_softmax(lg, 0, True)
    with pytest.raises(fb2.Fb2Error, match="exceeds num_experts"):
        fb2.moe.route_topk_softmax(lg, 9, True)


@pytest.mark.parametrize("batch,E,top_k", [(1, 128, 8), (32, 128, 8), (5, 4, 2), (300, 16, 4)])
def test_build_pairs_and_align(fb2, oracle, cuda, batch, E, top_k):
    """moe_build_pairs / moe_align_block_size vs the host plan (dispatch.rs:1408-1461): this
    implementation is deterministic, so arrays are compared exactly, not as sets."""
    lg = oracle.lcg_f32_array(0xBEEF + batch, batch * E, -2.0, 2.0)
    rid, _ = oracle.route(lg, batch, E, top_k, True)
    ids = cuda.from_numpy(rid.astype(np.int32)).cuda()
    pbt, pti, offs = fb2.moe.moe_build_pairs_by_token(ids, E, top_k)
    o_offs, o_pti, o_pbt = oracle.bucket_plan(rid, batch, E, top_k)
    assert np.array_equal(offs.cpu().numpy(), o_offs)
    assert np.array_equal(pti.cpu().numpy(), o_pti.astype(np.int32))
    assert np.array_equal(pbt.cpu().numpy(), o_pbt)
    for block in (8, 16, 64):
        for pair_ids in (False, True):
            cap = batch * top_k + E * block
            s, b, t = fb2.moe.moe_align_block_size(ids, E, block, cap, pair_ids)
            os_, ob, ot = oracle.moe_align_block_size(rid.astype(np.int32), E, block, cap, pair_ids)
            assert int(t.item()) == ot
            assert np.array_equal(s.cpu().numpy(), os_)
            assert np.array_equal(b.cpu().numpy()[: ot // block], ob[: ot // block])


def test_combine_overwrites_dirty_output(fb2, oracle, cuda):
    """moe_bucketed_parity_test.rs:160-163 + capabilities.rs:684-723."""
    batch, top_k, hidden = 5, 2, 64
    pairs = batch * top_k
    pd = oracle.f16_round(oracle.lcg_f32_array(1, pairs * hidden, -1.0, 1.0)).reshape(pairs, hidden)
    w = oracle.lcg_f32_array(2, pairs, 0.0, 1.0)
    pbt = np.random.default_rng(0).permutation(pairs).astype(np.int32)
    pbt[3] = -1
    out = cuda.full((batch, hidden), 7.0, dtype=cuda.float16, device="cuda")
    fb2.moe.moe_combine(cuda.from_numpy(pd.astype(np.float16)).cuda(), cuda.from_numpy(pbt).cuda(),
                        cuda.from_numpy(w).cuda(), out, batch, hidden, top_k)
    ref = oracle.moe_combine(pd, pbt, w, batch, hidden, top_k)
    assert np.abs(out.float().cpu().numpy() - ref).max() < 2e-3
    out2 = cuda.full((batch, hidden), -11.0, dtype=cuda.float16, device="cuda")
    fb2.moe.weighted_sum_batched(cuda.from_numpy(pd.astype(np.float16)).cuda(), cuda.from_numpy(w).cuda(),
                                 out2, batch, top_k, hidden)
    ref2 = oracle.moe_combine(pd, np.arange(pairs, dtype=np.int32), w, batch, hidden, top_k)
    assert np.abs(out2.float().cpu().numpy() - ref2).max() < 2e-3


def test_silu_gelu_split(fb2, oracle, cuda):
    t, im = 9, 96
    gu = oracle.f16_round(oracle.lcg_f32_array(4, t * 2 * im, -4.0, 4.0)).reshape(t, 2 * im)
    gud = cuda.from_numpy(gu.astype(np.float16)).cuda()
    out = cuda.empty((t, im), dtype=cuda.float16, device="cuda")
    fb2.moe.fused_silu_mul_split(gud, out, t, im)
    ref = oracle.silu_mul_split(gu, t, im)
    assert np.array_equal(out.cpu().numpy().view(np.uint16), ref.astype(np.float16).view(np.uint16)) or \
        np.abs(out.float().cpu().numpy() - ref).max() < 4e-3
    fb2.moe.fused_gelu_tanh_mul_split(gud, out, t, im)
    ref = oracle.gelu_tanh_mul_split(gu, t, im)
    assert np.abs(out.float().cpu().numpy() - ref).max() < 4e-3


# ------------------------------------------------------------------ routed grouped GEMM
@pytest.mark.parametrize("block", [8, 16, 64])
@pytest.mark.parametrize("pair_ids", [False, True])
def test_gemm_phase_vllm_contract(fb2, oracle, cuda, block, pair_ids):
    """marlin_moe_vllm_parity_test.rs:75-298 / SURVEY A.6: A row = id / top_k, C row = id, both
    id conventions.  Output starts dirty: this kernel does not need a pre-zeroed buffer."""
    k, n_per, E, batch, top_k = 256, 256, 4, 10, 2
    st, w, _, _ = _stack(fb2, oracle, k, n_per, 128, E, 0xCAFE0000)
    lg = oracle.lcg_f32_array(0x51, batch * E, -2.0, 2.0)
    rid, _ = oracle.route(lg, batch, E, top_k, True)
    pairs = batch * top_k
    x = oracle.f16_round(np.sin(np.arange(batch * k, dtype=np.float32) * np.float32(0.0027))).reshape(batch, k)
    ids = cuda.from_numpy(rid.astype(np.int32)).cuda()
    s, b, t = fb2.moe.moe_align_block_size(ids, E, block, None, pair_ids)
    out = cuda.full((pairs, n_per), 7.0, dtype=cuda.float16, device="cuda")
    if pair_ids:
        st.gemm_phase_vllm(cuda.from_numpy(x.astype(np.float16)).cuda(), s, b, t, out, batch, block, top_k)
        got = out.float().cpu().numpy()
        for p in range(pairs):
            ref = oracle.gemm(x[p // top_k], w[rid[p]], 1, n_per, k)
            assert _rel(oracle, got[p], ref) < 1e-2
    else:
        offs, pti, pbt = oracle.bucket_plan(rid, batch, E, top_k)
        xp = np.ascontiguousarray(x[pti.astype(int)])
        st.gemm_phase_vllm(cuda.from_numpy(xp.astype(np.float16)).cuda(), s, b, t, out, pairs, block, 1)
        got = out.float().cpu().numpy()
        for e in range(E):
            lo, hi = offs[e], offs[e + 1]
            if hi > lo:
                assert _rel(oracle, got[lo:hi], oracle.gemm(xp[lo:hi], w[e], hi - lo, n_per, k)) < 1e-2


# -------------------------------------------------------------------------- fused FFN
def _ffn_case(fb2, oracle, cuda, hidden, inter, E, top_k, batch, g, seed, act=None):
    gu_st, gu_w, _, _ = _stack(fb2, oracle, hidden, 2 * inter, g, E, seed, act)
    dn_st, dn_w, _, _ = _stack(fb2, oracle, inter, hidden, g, E, seed + 0x100, act)
    x = oracle.f16_round(oracle.lcg_f32_array(0xC0FFEE, batch * hidden, -1.0, 1.0)).reshape(batch, hidden)
    lg = oracle.f16_round(oracle.lcg_f32_array(0xFACE, batch * E, -2.0, 2.0)).reshape(batch, E)
    ffn = fb2.moe.MoeFfn(gu_st, dn_st, top_k, True, max_batch=batch)
    out = cuda.full((batch, hidden), 7.0, dtype=cuda.float16, device="cuda")      # dirty
    ffn.forward(cuda.from_numpy(x.astype(np.float16)).cuda(), cuda.from_numpy(lg.astype(np.float16)).cuda(), out)
    cuda.cuda.synchronize()
    rid, rw = oracle.route(lg, batch, E, top_k, True)
    ref = oracle.moe_forward_cpu(x, batch, hidden, inter, top_k, rid, rw, gu_w, dn_w, threads=8)
    return oracle.error_stats(out.float().cpu().numpy(), ref)


def test_moe_ffn_bucketed_fixture(fb2, oracle, cuda):
    """moe_bucketed_parity_test.rs:82-205 shape (hidden 64, inter 32, E=4, top-2, batch 5,
    g=32): generic kernels."""
    st = _ffn_case(fb2, oracle, cuda, 64, 32, 4, 2, 5, 32, 0xA)
    assert st["rel"] < 1e-2, st


@pytest.mark.parametrize("batch", [1, 4, 32])
@pytest.mark.parametrize("act", [None, "mod"])
def test_moe_ffn_tc(fb2, oracle, cuda, batch, act):
    """Tile-format experts (hidden 512, inter 256, E=8, top-2), fused SwiGLU epilogue."""
    st = _ffn_case(fb2, oracle, cuda, 512, 256, 8, 2, batch, 128, 0x5EED, act)
    assert st["rel"] < 1e-2 and st["nmse"] < 1e-4, st


@pytest.mark.parametrize("act,batches", [(None, (1, 2, 4, 8, 16, 32)), ("mod", (1, 8, 32, 40))])
def test_moe_ffn_qwen3_layer_shape(fb2, oracle, cuda, act, batches):
    """BASELINE.json configs[2] shape at full size: E=128, top-8, hidden 2048, inter 768, every decode batch of the
    bench; with act-order stacks the chained form runs (x gathered by the routing launch, act stored in the down
    stack's column order), and batch 40 takes the unchained path (separate gather launches, 16-row blocks)."""
    for batch in batches:
        st = _ffn_case(fb2, oracle, cuda, 2048, 768, 128, 8, batch, 128, 0x9E3, act)
        assert st["rel"] < 1e-2 and st["nmse"] < 1e-4, (batch, st)


@pytest.mark.parametrize("batch,E,top_k,norm", [(1, 60, 4, True), (7, 60, 4, False), (32, 128, 8, True), (32, 8, 1, True),
                                                 (5, 40, 6, True), (33, 16, 2, True), (3, 200, 8, True), (4, 320, 8, False),
                                                 (2, 64, 40, True)])
def test_moe_ffn_odd_router_shapes(fb2, oracle, cuda, batch, E, top_k, norm):
    """fb2_moe_ffn_forward across router shapes: expert counts that are not multiples of 32, top-1,
    heavy collisions (32 tokens on 8 experts), a batch just past the single-launch route+align kernel's limit
    (33 -> the two-kernel path), more than 128 / 256 experts (the 8- and 32-registers-per-lane forms of that kernel) and
    top_k > 32 (the two-kernel path again); against moe_forward_cpu (dispatch.rs:2208-2288)."""
    torch = cuda
    hidden, inter, g = 256, 128, 128
    gus = [oracle.make_synthetic(hidden, 2 * inter, g, 0x3100 + e, f16_scales=True) for e in range(E)]
    dns = [oracle.make_synthetic(inter, hidden, g, 0x4200 + e, f16_scales=True) for e in range(E)]
    mk = lambda parts, k, n: fb2.GptqExpertStack.load_gptq_stacked([p.qweight for p in parts], [p.scales for p in parts],
                                                                   [p.qzeros for p in parts], None, 4, g, k, n)
    guw = np.stack([oracle.dequant_gptq(p.qweight, p.scales, p.qzeros, None, 4, g, hidden, 2 * inter) for p in gus])
    dnw = np.stack([oracle.dequant_gptq(p.qweight, p.scales, p.qzeros, None, 4, g, inter, hidden) for p in dns])
    x = oracle.f16_round(oracle.lcg_f32_array(21, batch * hidden, -1.0, 1.0)).reshape(batch, hidden)
    lg = oracle.f16_round(oracle.lcg_f32_array(22, batch * E, -2.0, 2.0)).reshape(batch, E)
    ffn = fb2.moe.MoeFfn(mk(gus, hidden, 2 * inter), mk(dns, inter, hidden), top_k, norm, max_batch=batch)
    out = ffn.forward(torch.from_numpy(x.astype(np.float16)).cuda(), torch.from_numpy(lg.astype(np.float16)).cuda())
    torch.cuda.synchronize()
    ids, w = oracle.route(lg, batch, E, top_k, norm)
    ref = oracle.moe_forward_cpu(x, batch, hidden, inter, top_k, ids, w, guw, dnw)
    st = oracle.error_stats(out.float().cpu().numpy(), ref)
    assert st["rel"] < 1e-2 and st["nmse"] < 1e-5, st


@pytest.mark.parametrize("batch,E,top_k", [(8, 128, 8), (3, 200, 8), (32, 16, 4)])
def test_moe_ffn_router_ties(fb2, oracle, cuda, batch, E, top_k):
    """Logits on four levels only: almost every top-k decision is a tie, which the reference breaks towards the lower
    expert index (router.rs:160-183).  The single-launch route + align kernel picks the row maximum by a warp reduction and
    the lowest index holding it; a different tie rule would route tokens to other experts and miss the oracle by far."""
    torch = cuda
    hidden, inter, g = 256, 128, 128
    gus = [oracle.make_synthetic(hidden, 2 * inter, g, 0x5100 + e, f16_scales=True) for e in range(E)]
    dns = [oracle.make_synthetic(inter, hidden, g, 0x6200 + e, f16_scales=True) for e in range(E)]
    mk = lambda parts, k, n: fb2.GptqExpertStack.load_gptq_stacked([p.qweight for p in parts], [p.scales for p in parts],
                                                                   [p.qzeros for p in parts], None, 4, g, k, n)
    guw = np.stack([oracle.dequant_gptq(p.qweight, p.scales, p.qzeros, None, 4, g, hidden, 2 * inter) for p in gus])
    dnw = np.stack([oracle.dequant_gptq(p.qweight, p.scales, p.qzeros, None, 4, g, inter, hidden) for p in dns])
    x = oracle.f16_round(oracle.lcg_f32_array(31, batch * hidden, -1.0, 1.0)).reshape(batch, hidden)
    lg = np.floor(oracle.lcg_f32_array(32, batch * E, 0.0, 3.999)).reshape(batch, E).astype(np.float32) * 0.5 - 1.0
    ffn = fb2.moe.MoeFfn(mk(gus, hidden, 2 * inter), mk(dns, inter, hidden), top_k, True, max_batch=batch)
    out = ffn.forward(torch.from_numpy(x.astype(np.float16)).cuda(), torch.from_numpy(lg.astype(np.float16)).cuda())
    torch.cuda.synchronize()
    ids, w = oracle.route(lg, batch, E, top_k, True)
    ref = oracle.moe_forward_cpu(x, batch, hidden, inter, top_k, ids, w, guw, dnw)
    st = oracle.error_stats(out.float().cpu().numpy(), ref)
    assert st["rel"] < 1e-2 and st["nmse"] < 1e-5, st
