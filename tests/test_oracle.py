"""CPU tests of the oracle (oracle/): pinned against the reference's own known-answer
tests and identity tests, restated with the same fixtures, seeds and tolerances.
Reference paths are relative to the reference checkout."""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))


def silu(x):
    return x * (1.0 / (1.0 + math.exp(-x)))


# ------------------------------------------------------------------------- PRNG / fixtures
def test_lcg_c_matches_python(oracle):
    """gptq_parity_test.rs:28-38 -- the C LCG against a pure-python restatement."""
    import ctypes as C
    py = oracle.Lcg(0xDEADBEEF)
    st = C.c_uint64(0xDEADBEEF)
    for _ in range(1000):
        assert oracle.lib().oracle_rnd_u32(C.byref(st)) == py.u32()
    # known first outputs of the recurrence (computed by hand from the constants)
    s = (0 * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
    assert oracle.Lcg(0).u32() == (s >> 33)
    a = oracle.lcg_f32_array(0xC0FFEE, 64, -1.0, 1.0)
    py = oracle.Lcg(0xC0FFEE)
    assert np.array_equal(a, np.array([py.f32(-1.0, 1.0) for _ in range(64)], dtype=np.float32))
    assert (a >= -1.0).all() and (a < 1.0).all()


def test_make_synthetic_draw_order(oracle):
    """gptq_parity_test.rs:61-93: qweight words, then scales, then 8 nibble draws per zero word."""
    k, n, g = 16, 8, 8
    syn = oracle.make_synthetic(k, n, g, 0x1234)
    py = oracle.Lcg(0x1234)
    qw = np.array([py.u32() for _ in range((k // 8) * n)], dtype=np.uint32).view(np.int32)
    sc = np.array([py.f32(0.01, 0.1) for _ in range((k // g) * n)], dtype=np.float32)
    qz = []
    for _ in range((k // g) * (n // 8)):
        w = 0
        for bi in range(8):
            w |= (py.u32() & 0xF) << (bi * 4)
        qz.append(w)
    assert np.array_equal(syn.qweight.reshape(-1), qw)
    assert np.array_equal(syn.scales.reshape(-1), sc)
    assert np.array_equal(syn.qzeros.reshape(-1).view(np.uint32), np.array(qz, dtype=np.uint32))


# ------------------------------------------------------------------------------- dequant
def test_dequant_c_vs_numpy_bit_exact(oracle):
    """Two independent restatements of cpu.rs:2283-2315 agree bit for bit (sequential groups
    and g_idx groups, native_safetensors.rs:1391-1420)."""
    for (k, n, g, seed) in [(256, 128, 128, 0xDEADBEEF), (512, 256, 128, 0x51A7E5), (64, 64, 32, 7)]:
        syn = oracle.make_synthetic(k, n, g, seed)
        a = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, None, 4, g, k, n)
        b = oracle.np_dequant_gptq(syn.qweight, syn.scales, syn.qzeros, None, g, k, n)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        gi = oracle.make_desc_act_g_idx(k, g)
        a = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, g, k, n)
        b = oracle.np_dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, g, k, n)
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_dequant_rejects_bits(oracle):
    syn = oracle.make_synthetic(64, 64, 32, 1)
    with pytest.raises(ValueError, match="only bits=4"):
        oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, None, 8, 32, 64, 64)


def test_cpu_selfcheck(oracle):
    """gptq_parity_test.rs:109-145."""
    k, n, gs, m = 256, 128, 128, 2
    syn = oracle.make_synthetic(k, n, gs, 0xDEADBEEF)
    w = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, None, 4, gs, k, n)
    x = np.sin(np.arange(m * k, dtype=np.float32) * np.float32(0.001)).astype(np.float32)
    out = oracle.linear_forward(x, w, None, m)
    ref = oracle.gemm(x, oracle.np_dequant_gptq(syn.qweight, syn.scales, syn.qzeros, None, gs, k, n), m, n, k)
    assert np.abs(out - ref).max() < 1e-3
    # and against numpy float64 matmul (the f64-accumulate contract of cpu.rs:481-492)
    ref64 = (x.reshape(m, k).astype(np.float64) @ w.astype(np.float64).T).astype(np.float32)
    assert np.abs(out - ref64).max() < 1e-5


def test_gemm_thread_count_is_bit_identical(oracle):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 200)).astype(np.float32)
    b = rng.standard_normal((50, 200)).astype(np.float32)
    r1 = oracle.gemm(a, b, 3, 50, 200, threads=1)
    r4 = oracle.gemm(a, b, 3, 50, 200, threads=4)
    assert np.array_equal(r1.view(np.uint32), r4.view(np.uint32))


def test_desc_act_reference_uses_g_idx_for_scale_lookup(oracle):
    """gptq_parity_test.rs:233-254."""
    syn = oracle.make_synthetic(512, 256, 128, 0x51A7E5)
    gi = oracle.make_desc_act_g_idx(512, 128)
    assert (np.bincount(gi, minlength=4) == 128).all()
    seq = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, None, 4, 128, 512, 256)
    da = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, 128, 512, 256)
    assert np.abs(seq - da).max() > 1e-3


def test_symmetric_qzeros_encode_marlin_zero_point(oracle):
    """gptq_parity_test.rs:257-268."""
    syn = oracle.make_synthetic(512, 256, 128, 0x51A7E501, symmetric=True)
    assert (syn.qzeros.view(np.uint32) == 0x77777777).all()
    w = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, None, 4, 128, 512, 256)
    assert (w > 0).any() and (w < 0).any()


def test_desc_act_perm_gather_is_equivalent_to_g_idx_reference(oracle):
    """gptq_parity_test.rs:271-321 -- the identity the act-order fast path rests on."""
    k, n, g, m = 512, 256, 128, 3
    syn = oracle.make_synthetic(k, n, g, 0x5A170A7)
    gi = oracle.make_desc_act_g_idx(k, g)
    perm = np.argsort(gi, kind="stable")
    ref_w = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, g, k, n)
    # permute_qweight_rows_for_test (:191-230)
    qw = syn.qweight.view(np.uint32)
    unpacked = ((qw[:, None, :] >> (np.arange(8, dtype=np.uint32) * 4)[None, :, None]) & 0xF).reshape(k, n)
    srt = unpacked[perm]
    packed = np.zeros((k // 8, n), dtype=np.uint32)
    for bi in range(8):
        packed |= srt.reshape(k // 8, 8, n)[:, bi, :].astype(np.uint32) << np.uint32(bi * 4)
    sorted_w = oracle.dequant_gptq(packed.view(np.int32), syn.scales, syn.qzeros, None, 4, g, k, n)
    x = np.sin(np.arange(m * k, dtype=np.float32) * np.float32(0.0041)).astype(np.float32).reshape(m, k)
    out_ref = oracle.gemm(x, ref_w, m, n, k)
    out_perm = oracle.gemm(np.ascontiguousarray(x[:, perm]), sorted_w, m, n, k)
    assert np.abs(out_ref - out_perm).max() < 1e-3


def test_validate_desc_act_g_idx(oracle):
    """native_safetensors.rs:1327-1372 and its unit tests (:1534-1653)."""
    gi = oracle.make_desc_act_g_idx(256, 128)
    assert oracle.validate_desc_act_g_idx(gi, 256, 128) == 0
    bad = gi.copy(); bad[0] = 5
    assert oracle.validate_desc_act_g_idx(bad, 256, 128) == 3
    unbal = gi.copy(); unbal[0] = unbal[1]
    assert oracle.validate_desc_act_g_idx(unbal, 256, 128) == 4
    assert oracle.validate_desc_act_g_idx(gi[:200], 200, 128) == 2


def test_cpu_stacked_vs_per_expert_parity(oracle):
    """gptq_parity_test.rs:816-895 (tol 1e-4): a column slice of the stacked matrix equals the
    expert loaded alone."""
    k, n_per, gs, E, m = 256, 128, 128, 4, 3
    experts = [oracle.make_synthetic(k, n_per, gs, 0xA00DBA11 + e) for e in range(E)]
    x = np.sin(np.arange(m * k, dtype=np.float32) * np.float32(0.0017)).astype(np.float32)
    # stacked = row-wise interleaved concat along N (native_safetensors.rs:966-984)
    qw = np.concatenate([s.qweight for s in experts], axis=1)
    sc = np.concatenate([s.scales for s in experts], axis=1)
    qz = np.concatenate([s.qzeros for s in experts], axis=1)
    w_all = oracle.dequant_gptq(qw, sc, qz, None, 4, gs, k, E * n_per)
    for e, s in enumerate(experts):
        w = oracle.dequant_gptq(s.qweight, s.scales, s.qzeros, None, 4, gs, k, n_per)
        a = oracle.linear_forward(x, w, None, m)
        b = oracle.linear_forward(x, np.ascontiguousarray(w_all[e * n_per:(e + 1) * n_per]), None, m)
        assert np.abs(a - b).max() < 1e-4


# -------------------------------------------------------------------------------- router
@pytest.mark.parametrize("kat", KATS["router"], ids=lambda k: k["src"].split(":")[-1])
def test_router_kats(oracle, kat):
    ids, wts = oracle.route(np.array(kat["logits"], dtype=np.float32), kat["batch"], kat["experts"],
                            kat["top_k"], kat["norm"])
    if "ids" in kat:
        assert ids.tolist() == kat["ids"]
    if "weights" in kat:
        assert np.abs(wts - np.array(kat["weights"], dtype=np.float32)).max() < kat["tol"]
    if kat.get("weights_expr") == "softmax":
        lg = np.array(kat["logits"], dtype=np.float32)
        p = np.exp(lg) / np.exp(lg).sum()
        assert np.abs(wts - p[ids.astype(int)]).max() < kat["tol"]
    if "sum" in kat:
        assert abs(float(wts.sum()) - kat["sum"]) < kat["tol"]
    if "argmax_id" in kat:
        assert int(ids[int(np.argmax(wts))]) == kat["argmax_id"]
    for pos, val in kat.get("id_at", {}).items():
        assert int(ids[int(pos)]) == val


@pytest.mark.parametrize("kat", KATS["router_panics"], ids=lambda k: k["message"])
def test_router_panics(oracle, kat):
    with pytest.raises(AssertionError, match=kat["message"]):
        oracle.route(np.ones(kat["n_logits"], dtype=np.float32), kat["batch"], kat["experts"],
                     kat["top_k"], True)


def test_route_qwen3_shape_properties(oracle):
    """router.rs:197-265 (route vs route_into parity at E=128, K=8): properties of the result."""
    lg = oracle.lcg_f32_array(0xABCDEF, 4 * 128, -2.0, 2.0)
    ids, wts = oracle.route(lg, 4, 128, 8, True)
    ids, wts = ids.reshape(4, 8), wts.reshape(4, 8)
    for b in range(4):
        assert len(set(ids[b].tolist())) == 8
        assert abs(float(wts[b].sum()) - 1.0) < 1e-5
        assert (np.diff(wts[b]) <= 0).all()          # K passes of argmax -> descending
        top = np.argsort(-lg.reshape(4, 128)[b], kind="stable")[:8]
        assert ids[b].tolist() == top.tolist()


# ------------------------------------------------------------------ bucket plan / align
def test_bucket_plan_counting_sort(oracle):
    """dispatch.rs:1408-1461: expert-major runs, (b,k) order inside a run."""
    ids = np.array([2, 0, 1, 2, 0, 2], dtype=np.uint32)      # batch 3, top_k 2
    offs, pti, pbt = oracle.bucket_plan(ids, 3, 4, 2)
    assert offs.tolist() == [0, 2, 3, 6, 6]
    assert pti.tolist() == [0, 2, 1, 0, 1, 2]
    assert pbt.tolist() == [3, 0, 2, 4, 1, 5]


def test_align_block_size_oracle(oracle):
    ids = np.array([2, 0, 1, 2, 0, 2], dtype=np.int32)
    s, b, total = oracle.moe_align_block_size(ids, 4, 4, 6 + 4 * 4, pair_ids=True)
    assert total == 12
    assert s[:12].tolist() == [1, 4, 6, 6, 2, 6, 6, 6, 0, 3, 5, 6]
    assert b[:3].tolist() == [0, 1, 2]
    s2, _, _ = oracle.moe_align_block_size(ids, 4, 4, 22, pair_ids=False)
    assert s2[:12].tolist() == [0, 1, 6, 6, 2, 6, 6, 6, 3, 4, 5, 6]


# ------------------------------------------------------------------------- MoE forward
@pytest.mark.parametrize("kat", KATS["moe_forward"], ids=lambda k: k["src"].split(":")[-1])
def test_moe_forward_kats(oracle, kat):
    h, f, E = kat["hidden"], kat["ffn"], kat["experts"]
    gu = np.zeros((E, 2 * f, h), dtype=np.float32)
    dn = np.zeros((E, h, f), dtype=np.float32)
    for e, s in enumerate(kat["expert_scales"]):
        gu[e, :f] = np.eye(f, h) * s
        gu[e, f:] = np.eye(f, h) * s
        dn[e] = np.eye(h, f) * s
    out = oracle.moe_forward_cpu(np.array(kat["x"], dtype=np.float32), kat["batch"], h, f, kat["top_k"],
                                 np.array(kat["ids"], dtype=np.uint32),
                                 np.array(kat["weights"], dtype=np.float32), gu, dn)
    exp = np.array([eval(e, {"silu": silu}) for e in kat["expected_expr"]], dtype=np.float32)
    assert np.abs(out.reshape(-1) - exp).max() < kat["tol"]


def test_moe_forward_rejects_invalid_expert(oracle):
    """moe_dispatch_test.rs:208-232."""
    z = np.zeros((2, 4, 2), dtype=np.float32)
    with pytest.raises(ValueError, match="expert"):
        oracle.moe_forward_cpu(np.ones(2, np.float32), 1, 2, 2, 1, np.array([99], np.uint32),
                               np.ones(1, np.float32), z, np.zeros((2, 2, 2), np.float32))


def test_bucketed_matches_per_pair(oracle):
    """moe_bucketed_parity_test.rs:82-205 (hidden 64, inter 32, E=4, top-2, batch 5, g=32,
    seeds 0xA/0xB/0xC0FFEE, rel 1e-4): the bucketed dataflow (plan -> gather -> grouped GEMM
    -> SiLU -> grouped GEMM -> combine into a DIRTY buffer) equals per-pair moe_forward_cpu."""
    hidden, inter, E, top_k, batch, g = 64, 32, 4, 2, 5, 32
    gu = [oracle.make_synthetic(hidden, 2 * inter, g, 0xA + e) for e in range(E)]
    dn = [oracle.make_synthetic(inter, hidden, g, 0xB + e) for e in range(E)]
    gu_w = np.stack([oracle.dequant_gptq(s.qweight, s.scales, s.qzeros, None, 4, g, hidden, 2 * inter) for s in gu])
    dn_w = np.stack([oracle.dequant_gptq(s.qweight, s.scales, s.qzeros, None, 4, g, inter, hidden) for s in dn])
    import ctypes as C
    st = C.c_uint64(0xC0FFEE)
    f = oracle.lib().oracle_rnd_f32
    x = np.array([f(C.byref(st), -1.0, 1.0) for _ in range(batch * hidden)], dtype=np.float32)
    lg = np.array([f(C.byref(st), -2.0, 2.0) for _ in range(batch * E)], dtype=np.float32)
    ids, wts = oracle.route(lg, batch, E, top_k, True)
    out_a = oracle.moe_forward_cpu(x, batch, hidden, inter, top_k, ids, wts, gu_w, dn_w)
    offs, pti, pbt = oracle.bucket_plan(ids, batch, E, top_k)
    xp = x.reshape(batch, hidden)[pti.astype(int)]
    pairs = batch * top_k
    gup = np.zeros((pairs, 2 * inter), np.float32)
    for e in range(E):
        lo, hi = offs[e], offs[e + 1]
        if hi > lo:
            gup[lo:hi] = oracle.gemm(xp[lo:hi], gu_w[e], hi - lo, 2 * inter, hidden)
    act = oracle.silu_mul_split(gup, pairs, inter)
    dnp = np.zeros((pairs, hidden), np.float32)
    for e in range(E):
        lo, hi = offs[e], offs[e + 1]
        if hi > lo:
            dnp[lo:hi] = oracle.gemm(act[lo:hi], dn_w[e], hi - lo, hidden, inter)
    out_b = oracle.moe_combine(dnp, pbt, wts, batch, hidden, top_k)
    rel = np.abs(out_a - out_b).max() / max(np.abs(out_a).max(), 1e-6)
    assert rel < 1e-4


def test_silu_and_gelu(oracle):
    gu = np.array([[1.0, -2.0, 3.0, 0.5]], dtype=np.float32)
    s = oracle.silu_mul_split(gu, 1, 2)
    assert abs(s[0, 0] - silu(1.0) * 3.0) < 1e-6 and abs(s[0, 1] - silu(-2.0) * 0.5) < 1e-6
    g = oracle.gelu_tanh_mul_split(gu, 1, 2)
    ref = 0.5 * 1.0 * (1 + math.tanh(0.79788456 * (1 + 0.044715))) * 3.0
    assert abs(g[0, 0] - ref) < 1e-5


# --------------------------------------------------------------- committed golden vectors
def test_committed_golden_vectors(oracle):
    """tests/golden/gemm_small.npz was written by tests/golden/make_golden.py from this oracle;
    it pins the oracle's numerics across machines / compilers (GPU box included)."""
    z = np.load(os.path.join(HERE, "golden", "gemm_small.npz"))
    for tag in ("seq", "act"):
        k, n, g, m, seed = [int(v) for v in z[f"{tag}_meta"]]
        syn = oracle.make_synthetic(k, n, g, seed, f16_scales=True)
        gi = oracle.make_desc_act_g_idx(k, g) if tag == "act" else None
        w = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, g, k, n)
        assert np.array_equal(w.view(np.uint32), z[f"{tag}_w"].view(np.uint32))
        x = oracle.f16_round(oracle.lcg_f32_array(0xC0FFEE, m * k, -1.0, 1.0))
        y = oracle.linear_forward(x, w, None, m)
        assert np.array_equal(y.view(np.uint32), z[f"{tag}_y"].view(np.uint32))


def test_full_size_anchor_and_llama_down_against_numpy(oracle):
    """BASELINE.json configs[0] (K = N = 4096, g = 128, M = 1, act-order) and the Llama-3.1-8B down projection
    (K = 14336, N = 4096, M = 2) at FULL size: the C oracle's dequant equals the independent numpy restatement bit for
    bit, and its f64-accumulate GEMM (cpu.rs:481-492) equals numpy's float64 product rounded to f32 within one ulp."""
    for (k, n, m, seed) in [(4096, 4096, 1, 0xA11C), (14336, 4096, 2, 0xD0)]:
        syn = oracle.make_synthetic(k, n, 128, seed, f16_scales=True)
        gi = oracle.random_balanced_g_idx(k, 128, seed & 0xFFFF)
        w = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, 128, k, n)
        assert np.array_equal(w.view(np.uint32), oracle.np_dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 128, k, n).view(np.uint32))
        x = oracle.f16_round(oracle.lcg_f32_array(0xC0FFEE, m * k, -1.0, 1.0)).reshape(m, k)
        got = oracle.linear_forward(x, w, None, m, threads=8)
        ref = (x.astype(np.float64) @ w.astype(np.float64).T)
        # numpy's pairwise / blocked f64 summation differs from the sequential one by a few f64 ulps: far below f32's
        assert np.all(np.abs(got.astype(np.float64) - ref) <= np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64))
