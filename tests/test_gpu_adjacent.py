"""GPU parity tests of the adjacent fusions: residual add + RMSNorm with the act-order gather folded
into its store, the sorted-input forward that consumes it, the SwiGLU epilogue of a dense gate|up linear, row argmax.
References: crates/ferrum-kernels/kernels/fused_add_rms_norm.cu:21-100, kernels/argmax_rows.cu,
kernels/fused_silu_mul.cu:29, crates/ferrum-models/src/models/llama_family.rs:3468-3509."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rms_ref(x, r, w, eps):
    """fused_add_rms_norm_f16: sum rounded to f16 (the new residual), variance and scaling in f32, f16 out."""
    s = x.astype(np.float32) + (r.astype(np.float32) if r is not None else 0.0)
    s16 = s.astype(np.float16)
    sf = s16.astype(np.float32)
    inv = 1.0 / np.sqrt((sf * sf).mean(axis=1, keepdims=True, dtype=np.float32) + np.float32(eps))
    return s16, (sf * inv.astype(np.float32) * w.astype(np.float32)).astype(np.float16)


@pytest.mark.parametrize("tokens,hidden", [(1, 4096), (5, 2048), (33, 5376), (3, 96)])
@pytest.mark.parametrize("with_res", [False, True])
def test_fused_add_rms_norm(fb2, oracle, cuda, tokens, hidden, with_res):
    rng = np.random.default_rng(hidden + tokens)
    x = rng.uniform(-2, 2, size=(tokens, hidden)).astype(np.float16)
    r = rng.uniform(-2, 2, size=(tokens, hidden)).astype(np.float16) if with_res else None
    w = rng.uniform(0.5, 1.5, size=(hidden,)).astype(np.float16)
    xd, wd = cuda.from_numpy(x).cuda(), cuda.from_numpy(w).cuda()
    rd = cuda.from_numpy(r).cuda() if with_res else None
    out = fb2.fused_add_rms_norm(xd, wd, 1e-6, residual=rd)
    cuda.cuda.synchronize()
    res_ref, out_ref = _rms_ref(x, r, w, 1e-6)
    got = out.cpu().numpy().astype(np.float32)
    assert np.abs(got - out_ref.astype(np.float32)).max() <= 2e-3 * max(1.0, np.abs(out_ref.astype(np.float32)).max())
    if with_res:
        assert np.array_equal(rd.cpu().numpy().view(np.uint16), res_ref.view(np.uint16))     # the updated residual is exact


def test_norm_feeds_act_order_linear_without_a_gather_launch(fb2, oracle, cuda):
    """norm(out_perm = the linear's order) -> forward_sorted must equal norm -> forward bit for bit, with one launch less."""
    k, n, g, m = 1024, 512, 128, 3
    syn = oracle.make_synthetic(k, n, g, 0xAD1, f16_scales=True)
    gi = oracle.random_balanced_g_idx(k, g, 5)
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, g, k, n)
    assert lin.act_order_perm() is not None
    rng = np.random.default_rng(3)
    x = cuda.from_numpy(rng.uniform(-1, 1, size=(m, k)).astype(np.float16)).cuda()
    w = cuda.from_numpy(rng.uniform(0.5, 1.5, size=(k,)).astype(np.float16)).cuda()
    plain = fb2.fused_add_rms_norm(x, w, 1e-5)
    y_ref = lin.forward(plain)
    for mm in (m, 40):                                   # decode kernel and the 64-row-block kernel
        xx = x if mm == m else cuda.from_numpy(rng.uniform(-1, 1, size=(mm, k)).astype(np.float16)).cuda()
        a = lin.forward(fb2.fused_add_rms_norm(xx, w, 1e-5))
        fb2.reset_launch_count()
        b = lin.forward_sorted(fb2.fused_add_rms_norm(xx, w, 1e-5, out_perm=lin.act_order_perm()))
        cuda.cuda.synchronize()
        assert fb2.launch_count() == 2, fb2.launch_count()           # norm + GEMM: no gather launch
        assert cuda.equal(a, b), mm
    # a layer without act-order has no column order to offer
    lin2 = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, None, None, 4, g, k, n)
    assert lin2.act_order_perm() is None
    assert cuda.equal(lin2.forward_sorted(plain), lin2.forward(plain))
    assert y_ref.shape == (m, n)


@pytest.mark.parametrize("act", [None, "rand"])
@pytest.mark.parametrize("m", [1, 3, 8, 12, 70])
def test_dense_swiglu_epilogue(fb2, oracle, cuda, act, m):
    """gate|up linear + SiLU(gate) * up in one launch (decode kernel up to 8 rows, 64-row blocks beyond) against
    CpuGptqLinear::forward + fused_silu_mul_split of the oracle (gate | up rounded to f16 in between, as the pair does)."""
    k, inter, g = 1024, 768, 128
    syn = oracle.make_synthetic(k, 2 * inter, g, 0x5116 + m, f16_scales=True)
    gi = oracle.random_balanced_g_idx(k, g, 9) if act else None
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, g, k, 2 * inter)
    x = oracle.f16_round(oracle.lcg_f32_array(0xBEEF + m, m * k, -1.0, 1.0)).reshape(m, k)
    xd = cuda.from_numpy(x.astype(np.float16)).cuda()
    fb2.reset_launch_count()
    y = lin.forward_swiglu(xd)
    cuda.cuda.synchronize()
    w_ref = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, g, k, 2 * inter)
    ref = oracle.silu_mul_split(oracle.f16_round(oracle.gemm(x, w_ref, m, 2 * inter, k)), m, inter)
    st = oracle.error_stats(y.float().cpu().numpy(), ref)
    assert st["rel"] < 1e-2 and st["nmse"] < 1e-5, (m, act, st)
    # and against the two-launch form of this library
    two = cuda.empty((m, inter), dtype=cuda.float16, device="cuda")
    fb2.moe.fused_silu_mul_split(lin.forward(xd), two, m, inter)
    d = (y.float() - two.float()).abs().max().item() / max(two.float().abs().max().item(), 1e-6)
    assert d < 4e-3, (m, act, d)


def test_argmax_rows(fb2, oracle, cuda):
    rng = np.random.default_rng(11)
    for (m, n) in [(1, 128256), (7, 1000), (32, 151936), (3, 17)]:
        lg = rng.uniform(-8, 8, size=(m, n)).astype(np.float16)
        lg[0, n // 2] = lg[0].max()                         # a tie: the lowest index wins
        got = fb2.argmax_rows(cuda.from_numpy(lg).cuda()).cpu().numpy()
        assert np.array_equal(got, lg.astype(np.float32).argmax(axis=1).astype(np.int32)), (m, n)
