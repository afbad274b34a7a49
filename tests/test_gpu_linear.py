"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI,
against the CPU oracle on the reference tests' own fixtures, seeds and shapes.

Tolerances: dequant int4 -> f32 bit-exact; -> f16 equals f16_rn(oracle);
GEMM f16 out vs oracle (f16-rounded inputs, f64 accumulate): max-abs/max-ref < 1e-2
(reference gate 0.05, gptq_parity_test.rs:385-404) and NMSE < 1e-5."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-2
NMSE_TOL = 1e-5


def _make(oracle, k, n, g, seed, act=None, symmetric=False):
    syn = oracle.make_synthetic(k, n, g, seed, symmetric=symmetric, f16_scales=True)
    if act == "mod":
        gi = oracle.make_desc_act_g_idx(k, g)
    elif act == "rand":
        gi = oracle.random_balanced_g_idx(k, g, seed & 0xFFFF)
    elif act == "trivial":
        gi = (np.arange(k) // g).astype(np.int32)
    else:
        gi = None
    return syn, gi


def _x(oracle, m, k, seed=0xC0FFEE):
    return oracle.f16_round(oracle.lcg_f32_array(seed, m * k, -1.0, 1.0)).reshape(m, k)


def _check_forward(oracle, cuda, lin, syn, gi, m, x=None, bias=None, threads=8):
    k, n = syn.k, syn.n
    w_ref = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, syn.group_size, k, n)
    if x is None:
        x = _x(oracle, m, k)
    xd = cuda.from_numpy(x.astype(np.float16)).cuda()
    y = lin.forward(xd)
    cuda.cuda.synchronize()
    ref = oracle.linear_forward(x, w_ref, bias, m, threads=threads)
    st = oracle.error_stats(y.float().cpu().numpy(), ref)
    assert st["rel"] < REL_TOL and st["nmse"] < NMSE_TOL, (lin.kernel_name(m), m, st)
    return st


# ----------------------------------------------------------------------------- dequant
@pytest.mark.parametrize("k,n,g,act", [(256, 128, 128, None), (512, 256, 128, "mod"), (512, 256, 128, "rand"),
                                        (1024, 384, 128, "trivial"), (64, 64, 32, None), (256, 136, 64, "mod")])
def test_dequant_bit_exact(fb2, oracle, cuda, k, n, g, act):
    syn, gi = _make(oracle, k, n, g, 0xDEADBEEF, act)
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, g, k, n)
    ref = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, g, k, n)
    w32 = lin.dequantize(cuda.float32).cpu().numpy()
    assert np.array_equal(w32.view(np.uint32), ref.view(np.uint32))
    w16 = lin.dequantize(cuda.float16).cpu().numpy()
    assert np.array_equal(w16.view(np.uint16), ref.astype(np.float16).view(np.uint16))
    assert lin.is_act_order == (act in ("mod", "rand"))
    assert lin.in_features() == k and lin.out_features() == n


def test_dequant_symmetric(fb2, oracle, cuda):
    syn, _ = _make(oracle, 512, 256, 128, 0x51A7E501, symmetric=True)
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, None, None, 4, 128, 512, 256)
    ref = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, None, 4, 128, 512, 256)
    assert np.array_equal(lin.dequantize().cpu().numpy().view(np.uint32), ref.view(np.uint32))


# ------------------------------------------------------------------- reference fixtures
def test_cuda_vs_cpu(fb2, oracle, cuda):
    """gptq_parity_test.rs:329-405: K=512, N=256, seed 0xC0FFEE, cos(0.003 i), m=2."""
    syn, _ = _make(oracle, 512, 256, 128, 0xC0FFEE)
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, None, None, 4, 128, 512, 256)
    x = oracle.f16_round(np.cos(np.arange(2 * 512, dtype=np.float32) * np.float32(0.003))).reshape(2, 512)
    assert lin.kernel_name(2) == "tc-streamk"
    _check_forward(oracle, cuda, lin, syn, None, 2, x=x)


def test_cuda_desc_act_vs_cpu_reference(fb2, oracle, cuda):
    """gptq_parity_test.rs:580 (act-order through perm)."""
    syn, gi = _make(oracle, 512, 256, 128, 0x5A170A7, "mod")
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, 512, 256)
    x = oracle.f16_round(np.sin(np.arange(3 * 512, dtype=np.float32) * np.float32(0.0041))).reshape(3, 512)
    _check_forward(oracle, cuda, lin, syn, gi, 3, x=x)


@pytest.mark.parametrize("m", [1, 2, 3, 4, 8, 9, 15, 16, 17, 24, 31, 32, 33, 64, 65, 200])
@pytest.mark.parametrize("act", [None, "rand"])
def test_decode_and_small_prefill_rows(fb2, oracle, cuda, m, act):
    syn, gi = _make(oracle, 1024, 512, 128, 0xABC + m, act)
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, 1024, 512)
    _check_forward(oracle, cuda, lin, syn, gi, m)


def test_bias(fb2, oracle, cuda):
    syn, _ = _make(oracle, 512, 256, 128, 0xB1A5)
    bias = oracle.lcg_f32_array(5, 256, -1.0, 1.0)
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, None, bias, 4, 128, 512, 256)
    _check_forward(oracle, cuda, lin, syn, None, 5, bias=bias)


def test_anchor_4096(fb2, oracle, cuda):
    """BASELINE.json configs[0]: K=N=4096, g=128, M=1 (act-order on)."""
    syn, gi = _make(oracle, 4096, 4096, 128, 0xA11C, "rand")
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, 4096, 4096)
    for m in (1, 32):
        _check_forward(oracle, cuda, lin, syn, gi, m)


@pytest.mark.parametrize("k,n", [(5376, 4096), (5376, 2048), (4096, 5376)])
def test_gemma3_shapes_desc_act(fb2, oracle, cuda, k, n):
    """gptq_parity_test.rs:592-660 (Gemma3-27B shapes, desc_act).  The two 21504 / 43008-wide shapes of the
    same fixture are checked at full size by test_gemma3_wide_shapes_full_size (sampled columns)."""
    syn, gi = _make(oracle, k, n, 128, 0x6E33A + k + n, "mod")
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, k, n)
    _check_forward(oracle, cuda, lin, syn, gi, 2)


def test_generic_shapes(fb2, oracle, cuda):
    """Shapes outside the tile format (group 32/64, N not a multiple of 128) take the generic
    kernel and must match too (the CPU reference accepts them, cpu.rs:2283)."""
    for (k, n, g, act) in [(64, 64, 32, None), (256, 136, 64, "mod"), (128, 72, 128, None)]:
        syn, gi = _make(oracle, k, n, g, 0x9E4 + n, act)
        lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, g, k, n)
        assert lin.kernel_name(1) == "generic"
        for m in (1, 5, 19):
            _check_forward(oracle, cuda, lin, syn, gi, m)


def test_empty_and_host_call(fb2, oracle, cuda):
    syn, gi = _make(oracle, 512, 256, 128, 0xE0, "mod")
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, 512, 256)
    out = lin.forward(cuda.empty((0, 512), dtype=cuda.float16, device="cuda"))
    assert out.shape == (0, 256)
    x = _x(oracle, 7, 512)
    y = lin.forward_host(x.astype(np.float16)).astype(np.float32)
    ref = oracle.linear_forward(x, oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, 128, 512, 256), None, 7)
    st = oracle.error_stats(y, ref)
    assert st["rel"] < REL_TOL, st


def test_unbalanced_act_order_still_correct(fb2, oracle, cuda):
    """Unbalanced g_idx is rejected by the reference's CUDA path (native_safetensors.rs:1327);
    the CPU reference handles it, and so does the generic kernel here."""
    k, n, g = 256, 128, 128
    syn, _ = _make(oracle, k, n, g, 0x0B)
    gi = np.zeros(k, np.int32); gi[200:] = 1
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, g, k, n)
    assert lin.kernel_name(1) == "generic"
    _check_forward(oracle, cuda, lin, syn, gi, 3)


# ------------------------------------------------------ full-size, size-independent checks
def test_full_size_linearity_and_column_probe(fb2, oracle, cuda):
    """Llama-3.1-8B down_proj (K=14336, N=4096) at full size: (a) rows of a batch are
    independent (forward of a batch == forwards of its rows), (b) unit-vector inputs read back
    single weight columns == the bit-exact dequant, (c) a sampled set of output columns matches
    the oracle."""
    k, n = 14336, 4096
    syn, gi = _make(oracle, k, n, 128, 0x11A3A, "rand")
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, k, n)
    x = _x(oracle, 8, k)
    xd = cuda.from_numpy(x.astype(np.float16)).cuda()
    y8 = lin.forward(xd)
    y1 = cuda.cat([lin.forward(xd[i:i + 1].contiguous()) for i in range(8)])
    cuda.cuda.synchronize()
    assert (y8.float() - y1.float()).abs().max().item() <= 2e-2 * y8.float().abs().max().item()
    # (b) e_j probes: y = W[:, j]
    cols = [0, 1, 127, 128, 5000, k - 1]
    e = cuda.zeros((len(cols), k), dtype=cuda.float16, device="cuda")
    for i, j in enumerate(cols):
        e[i, j] = 1.0
    yw = lin.forward(e).cpu().numpy()
    w16 = lin.dequantize(cuda.float16).cpu().numpy()
    for i, j in enumerate(cols):
        assert np.array_equal(yw[i].view(np.uint16), w16[:, j].view(np.uint16))
    # (c) sampled columns vs oracle
    w_ref = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, 128, k, n)
    pick = np.arange(0, n, 37)
    ref = oracle.gemm(x, np.ascontiguousarray(w_ref[pick]), 8, len(pick), k, threads=8)
    st = oracle.error_stats(y8.float().cpu().numpy()[:, pick], ref)
    assert st["rel"] < REL_TOL and st["nmse"] < NMSE_TOL, st


def _sampled_forward_check(fb2, oracle, cuda, k, n, seed, act, ms, col_stride, row_cap=48):
    """Full-size linear against the oracle on a SAMPLE of the output: every col_stride-th block of 8 columns (the
    oracle dequantises only those columns: the packed tensors are sliced on 8-column boundaries, where the zero-point
    words split cleanly) and at most row_cap rows spread over the batch.  The GPU computes everything."""
    syn, gi = _make(oracle, k, n, 128, seed, act)
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, k, n)
    blocks = np.arange(0, n // 8, col_stride)
    cols = (blocks[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
    w_ref = oracle.dequant_gptq(np.ascontiguousarray(syn.qweight[:, cols]), np.ascontiguousarray(syn.scales[:, cols]),
                                np.ascontiguousarray(syn.qzeros[:, blocks]), gi, 4, 128, k, len(cols))
    out = {}
    for m in ms:
        x = _x(oracle, m, k, seed=0xC0FFEE + m)
        y = lin.forward(cuda.from_numpy(x.astype(np.float16)).cuda())
        cuda.cuda.synchronize()
        rows = np.unique(np.linspace(0, m - 1, min(m, row_cap)).astype(int))
        ref = oracle.gemm(np.ascontiguousarray(x[rows]), w_ref, len(rows), len(cols), k, threads=8)
        st = oracle.error_stats(y.float().cpu().numpy()[np.ix_(rows, cols)], ref)
        assert st["rel"] < REL_TOL and st["nmse"] < NMSE_TOL, (lin.kernel_name(m), k, n, m, st)
        out[m] = st
    return out


@pytest.mark.parametrize("name,k,n", [("qkv_proj", 4096, 6144), ("o_proj", 4096, 4096), ("gate_up_proj", 4096, 28672),
                                      ("down_proj", 14336, 4096)])
def test_llama_shapes_full_size(fb2, oracle, cuda, name, k, n):
    """BASELINE.json configs[1] at full size, act-order: every decode kernel class (M = 1: in-kernel gather; 8, 16:
    gather launch + decode kernel, distributed owners from 16 rows; 32: its 32-row form) and the prefill class the bench times (M = 512)."""
    _sampled_forward_check(fb2, oracle, cuda, k, n, 0x77A0 + n, "rand", (1, 2, 8, 16, 32, 512), col_stride=max(1, n // 8 // 24))


@pytest.mark.parametrize("name,k,n,ms", [("gate_up_proj", 5376, 43008, (2, 4096)), ("down_proj", 21504, 5376, (2, 4096))])
def test_gemma3_wide_shapes_full_size(fb2, oracle, cuda, name, k, n, ms):
    """gptq_parity_test.rs:592-660: the Gemma3-27B shapes with N = 43008 and K = 21504 (desc_act), at the reference
    test's M = 2 and at the prefill batch the tensor-parallel bench runs (M = 4096: 256-row blocks)."""
    _sampled_forward_check(fb2, oracle, cuda, k, n, 0x6E33A + k + n, "mod", ms, col_stride=max(1, n // 8 // 16), row_cap=24)


def test_graph_capture_and_replay(fb2, oracle, cuda):
    """capabilities.rs:205-214: forward must be capturable (no alloc, no sync, stable pointers)."""
    syn, gi = _make(oracle, 1024, 512, 128, 0x6A, "rand")
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, 1024, 512)
    x = cuda.from_numpy(_x(oracle, 4, 1024).astype(np.float16)).cuda()
    y = cuda.empty((4, 512), dtype=cuda.float16, device="cuda")
    lin.forward(x, y)
    cuda.cuda.synchronize()
    eager = y.clone()
    y.zero_()
    g = cuda.cuda.CUDAGraph()
    s = cuda.cuda.Stream()
    with cuda.cuda.stream(s):
        with cuda.cuda.graph(g, stream=s):
            lin.forward(x, y)
    for _ in range(3):
        g.replay()
    cuda.cuda.synchronize()
    assert cuda.equal(y, eager)      # deterministic split-K: bit-identical replays


@pytest.mark.parametrize("pinned", [False, True])
def test_host_batch_pageable_and_pinned(fb2, oracle, cuda, pinned):
    """fb2_gptq_linear_forward_host_batch: 40 linears (three chunks of the H2D / GEMM / D2H pipeline,
    the last one partial) from pageable host buffers (staged) and from page-locked ones (direct DMA);
    both must equal the device-buffer forward bit for bit."""
    torch = cuda
    k, n, m = 512, 256, 5
    syn, gi = _make(oracle, k, n, 128, 0x51, "rand")
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, k, n)
    lins = [lin] * 40                              # (40 entries: chunks of 8)
    mk = (lambda a: torch.from_numpy(a).pin_memory().numpy()) if pinned else (lambda a: a)
    xs = [mk(oracle.f16_round(oracle.lcg_f32_array(100 + i, m * k, -1.0, 1.0)).reshape(m, k).astype(np.float16)) for i in range(40)]
    outs = [mk(np.zeros((m, n), dtype=np.float16)) for _ in range(40)]
    fb2.forward_host_batch(lins, xs, m, outs)
    for i in (0, 15, 16, 39):
        ref = lin.forward(torch.from_numpy(np.ascontiguousarray(xs[i])).cuda()).cpu().numpy()
        assert np.array_equal(outs[i].view(np.uint16), ref.view(np.uint16)), i


def test_host_batch_cached_graphs(fb2, oracle, cuda):
    """Repeated fb2_gptq_linear_forward_host_batch calls with the same linears and m replay cached per-chunk launch graphs
    (first call eager, second captures, third replays): every call must equal the device-buffer forward bit for bit, with
    new inputs each time, and a handle destroyed and re-created in between must invalidate the cache."""
    torch = cuda
    k, n, m = 512, 256, 3
    syn, gi = _make(oracle, k, n, 128, 0x77, "rand")
    mkl = lambda: fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, k, n)
    lins = [mkl() for _ in range(20)]
    outs = [np.zeros((m, n), dtype=np.float16) for _ in range(20)]
    def check(rep):
        xs = [oracle.f16_round(oracle.lcg_f32_array(1000 * rep + i, m * k, -1.0, 1.0)).reshape(m, k).astype(np.float16) for i in range(20)]
        for o in outs: o[...] = 0
        fb2.forward_host_batch(lins, xs, m, outs)
        for i in (0, 15, 16, 19):
            ref = lins[i].forward(torch.from_numpy(np.ascontiguousarray(xs[i])).cuda()).cpu().numpy()
            assert np.array_equal(outs[i].view(np.uint16), ref.view(np.uint16)), (rep, i)
    for rep in range(4):
        check(rep)
    n0 = fb2.launch_count()
    check(4)
    assert fb2.launch_count() - n0 >= 20          # replayed graphs still count their kernels
    lins[7] = None                                 # destroy one handle ...
    import gc; gc.collect()
    lins[7] = mkl()                                # ... and create another (possibly at the same address)
    for rep in range(5, 8):
        check(rep)


def test_host_batch_rows_per_entry(fb2, oracle, cuda):
    """fb2_gptq_linear_forward_host_batch_m: one call, a different row count per entry (zero rows = skipped), three times
    (eager, capture, replay); bit-identical to the device-buffer forward."""
    torch = cuda
    k, n = 512, 256
    syn, gi = _make(oracle, k, n, 128, 0x78, "rand")
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, k, n)
    ms = [1, 2, 0, 4, 8, 16, 32, 3, 5, 7, 9, 17, 33, 1, 64, 2, 6, 31]
    lins = [lin] * len(ms)
    for rep in range(3):
        xs = [oracle.f16_round(oracle.lcg_f32_array(50 * rep + i, max(m, 1) * k, -1.0, 1.0)).reshape(max(m, 1), k).astype(np.float16) for i, m in enumerate(ms)]
        outs = [np.zeros((max(m, 1), n), dtype=np.float16) for m in ms]
        hb = fb2.HostBatch(lins, xs, outs, ms)
        hb.run()
        for i, m in enumerate(ms):
            if m == 0:
                assert not outs[i].any()
                continue
            ref = lin.forward(torch.from_numpy(np.ascontiguousarray(xs[i][:m])).cuda()).cpu().numpy()
            assert np.array_equal(outs[i][:m].view(np.uint16), ref.view(np.uint16)), (rep, i, m)


@pytest.mark.parametrize("k,n", [(512, 4096), (1792, 4096), (3584, 4096), (4096, 768), (4096, 3584), (4096, 1536),
                                 (7168, 4096), (2048, 4096)])
def test_tp_shard_shapes_decode_rows(fb2, oracle, cuda, k, n):
    """The per-rank shapes of Llama-3.1-8B under TP 2 / 4 / 8 (row-parallel o / down: K / tp; column-parallel qkv /
    gate_up: N / tp) at every decode row count the bench uses -- few units per tile and many contributors per split tile,
    the corner the stream-K exchange variants (single owner, distributed owners, counter) have to survive."""
    _sampled_forward_check(fb2, oracle, cuda, k, n, 0x7E57 + k + n, "rand", (1, 2, 4, 8, 16, 32), 4)


@pytest.mark.parametrize("k,n", [(4096, 256), (8192, 128), (2048, 384), (14336, 128)])
def test_many_contributors_per_tile(fb2, oracle, cuda, k, n):
    """Deep, narrow layers: one to three tiles cut across 16+ CTAs -- more contributors than a split tile has row
    vectors (several distributed owners idle), more than one polling round per owner -- at the row counts of every
    exchange variant, twice (the flag-in-value slots must be back to empty after a launch)."""
    syn, gi = _make(oracle, k, n, 128, 0x5EED + k + n, "rand")
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, 128, k, n)
    w_ref = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, 128, k, n)
    for m in (1, 4, 8, 16, 32):
        x = _x(oracle, m, k, seed=0xABC + m)
        xd = cuda.from_numpy(x.astype(np.float16)).cuda()
        y1 = lin.forward(xd)
        y2 = lin.forward(xd)
        cuda.cuda.synchronize()
        assert cuda.equal(y1, y2), m
        ref = oracle.gemm(x, w_ref, m, n, k, threads=8)
        st = oracle.error_stats(y1.float().cpu().numpy(), ref)
        assert st["rel"] < REL_TOL and st["nmse"] < NMSE_TOL, (lin.kernel_name(m), k, n, m, st)
