"""Safetensors GPTQ checkpoint ingest (csrc/checkpoint.cu) against an independent
Python writer of the container format and the numpy statement of the reference's fusion rules
(native_safetensors.rs:887-1102 fused linears, :388-585 stacked experts, :1242-1246 sym
canonicalisation, :1288-1325 g_idx validation, :1475-1530 quantisation config).  The host-side
calls need no GPU; the gpu-marked tests load the same checkpoints onto the device and compare the
forward with the oracle."""
import json
import os
import struct

import numpy as np
import pytest

_DT = {np.dtype(np.float32): "F32", np.dtype(np.float16): "F16", np.dtype(np.int32): "I32",
       np.dtype(np.int64): "I64", np.dtype(np.uint8): "U8"}


def write_safetensors(path, tensors, metadata=None, bf16=()):
    """The published container: u64-LE header length, JSON header, data blob."""
    header, blob = {}, b""
    if metadata:
        header["__metadata__"] = metadata
    for name, a in tensors.items():
        a = np.ascontiguousarray(a)
        if name in bf16:        # store the upper halves of f32 words
            raw = (a.astype(np.float32).view(np.uint32) >> 16).astype(np.uint16).tobytes()
            dt = "BF16"
        else:
            raw, dt = a.tobytes(), _DT[a.dtype]
        header[name] = {"dtype": dt, "shape": list(a.shape), "data_offsets": [len(blob), len(blob) + len(raw)]}
        blob += raw
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)) + hj + blob)


def gptq_tensors(oracle, prefix, k, n, g, seed, g_idx=None, bias=None, f16_scales=True):
    syn = oracle.make_synthetic(k, n, g, seed, f16_scales=True)
    t = {f"{prefix}.qweight": syn.qweight.reshape(k // 8, n),
         f"{prefix}.scales": syn.scales.reshape(k // g, n).astype(np.float16 if f16_scales else np.float32),
         f"{prefix}.qzeros": syn.qzeros.reshape(k // g, n // 8)}
    if g_idx is not None:
        t[f"{prefix}.g_idx"] = np.asarray(g_idx, np.int32)
    if bias is not None:
        t[f"{prefix}.bias"] = np.asarray(bias, np.float16)
    return syn, t


def qcfg(dirname, **kw):
    d = {"quant_method": "gptq", "bits": 4, "group_size": 128, "desc_act": False, "sym": False}
    d.update(kw)
    with open(os.path.join(dirname, "quantize_config.json"), "w") as f:
        json.dump(d, f)


# ------------------------------------------------------------------------------ host side
def test_single_file_linear_roundtrip(fb2, oracle, tmp_path):
    k, n, g = 256, 128, 128
    bias = np.linspace(-1, 1, n).astype(np.float16)
    syn, t = gptq_tensors(oracle, "model.layers.0.self_attn.o_proj", k, n, g, 0xDEADBEEF, bias=bias)
    t["model.norm.weight"] = np.arange(8, dtype=np.float32)
    write_safetensors(tmp_path / "model.safetensors", t, metadata={"format": "pt"})
    qcfg(tmp_path)
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    q = ld.quant_config()
    assert (q.method, q.bits, q.group_size, q.desc_act, q.sym) == ("gptq", 4, 128, False, False)
    assert sorted(ld.tensor_names()) == sorted(t)
    assert ld.has_tensor("model.norm.weight") and not ld.has_tensor("model.norm.bias")
    assert ld.tensor_info("model.layers.0.self_attn.o_proj.scales") == ("F16", (2, 128), 2 * 128 * 2)
    np.testing.assert_array_equal(ld.read_f32("model.norm.weight"), t["model.norm.weight"])
    info = ld.gptq_info("model.layers.0.self_attn.o_proj")
    assert info == {"k": k, "n": n, "has_g_idx": False, "has_bias": True, "desc_act": False, "part_n": [n]}
    qw, sc, qz, gi, bs, kk, nn = ld.read_gptq_raw("model.layers.0.self_attn.o_proj")
    assert (kk, nn) == (k, n) and gi is None
    np.testing.assert_array_equal(qw, syn.qweight)
    np.testing.assert_array_equal(sc, syn.scales)                    # f16 on disk, widened exactly
    np.testing.assert_array_equal(qz, syn.qzeros)
    np.testing.assert_array_equal(bs, bias.astype(np.float32))
    with pytest.raises(fb2.Fb2Error) as e:
        ld.read_i32("model.norm.weight")
    assert e.value.code == -6 and "expected I32" in str(e.value)
    with pytest.raises(fb2.Fb2Error) as e:
        ld.read_f32("nope")
    assert "not in index" in str(e.value)


def test_bf16_and_f32_scales_widen_exactly(fb2, oracle, tmp_path):
    k, n, g = 128, 64, 128
    syn, t = gptq_tensors(oracle, "lin", k, n, g, 7, f16_scales=False)
    vals = np.array([1.0, -2.5, 3.140625, 1e-3, 65280.0, -0.0], np.float32)
    vals = (vals.view(np.uint32) & 0xFFFF0000).view(np.float32)
    t["b"] = vals
    write_safetensors(tmp_path / "model.safetensors", t, bf16=("b",))
    qcfg(tmp_path)
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    assert ld.tensor_info("b")[0] == "BF16"
    np.testing.assert_array_equal(ld.read_f32("b").view(np.uint32), vals.view(np.uint32))
    np.testing.assert_array_equal(ld.read_gptq_raw("lin")[1], syn.scales)


def test_sharded_index_and_fusion_shims(fb2, oracle, tmp_path):
    """qkv_proj / gate_up_proj are synthesised from their parts (row-major concat along N), parts may
    live in different shard files, shards are the sorted distinct values of weight_map."""
    k, g = 256, 128
    gi = oracle.make_desc_act_g_idx(k, g)
    pre = "model.layers.3.self_attn."
    widths = {"q_proj": 256, "k_proj": 64, "v_proj": 64}
    parts, shard_a, shard_b = {}, {}, {}
    for i, (p, w) in enumerate(widths.items()):
        syn, t = gptq_tensors(oracle, pre + p, k, w, g, 100 + i, g_idx=gi, bias=np.full(w, i + 1, np.float16))
        parts[p] = syn
        (shard_a if p != "k_proj" else shard_b).update(t)
    mg, tg = gptq_tensors(oracle, "model.layers.3.mlp.gate_proj", k, 128, g, 200)
    mu, tu = gptq_tensors(oracle, "model.layers.3.mlp.up_proj", k, 128, g, 201)
    shard_b.update(tg); shard_b.update(tu)
    write_safetensors(tmp_path / "model-00002-of-00002.safetensors", shard_b)
    write_safetensors(tmp_path / "model-00001-of-00002.safetensors", shard_a)
    wm = {name: "model-00001-of-00002.safetensors" for name in shard_a}
    wm.update({name: "model-00002-of-00002.safetensors" for name in shard_b})
    with open(tmp_path / "model.safetensors.index.json", "w") as f:
        json.dump({"metadata": {"total_size": 1}, "weight_map": wm}, f)
    # transformers-style config instead of quantize_config.json (:1487-1527)
    with open(tmp_path / "config.json", "w") as f:
        json.dump({"hidden_size": k, "quantization_config": {"quant_method": "GPTQ", "bits": 4, "group_size": 128,
                                                             "desc_act": True, "sym": False, "extra": [1, {"a": None}]}}, f)
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    q = ld.quant_config()
    assert (q.method, q.bits, q.group_size, q.desc_act) == ("gptq", 4, 128, True)
    info = ld.gptq_info(pre + "qkv_proj")
    assert info == {"k": k, "n": 384, "has_g_idx": True, "has_bias": True, "desc_act": True, "part_n": [256, 64, 64]}
    qw, sc, qz, g_idx, bs, kk, nn = ld.read_gptq_raw(pre + "qkv_proj")
    order = ["q_proj", "k_proj", "v_proj"]
    np.testing.assert_array_equal(qw, np.concatenate([parts[p].qweight.reshape(k // 8, -1) for p in order], axis=1))
    np.testing.assert_array_equal(sc, np.concatenate([parts[p].scales.reshape(k // g, -1) for p in order], axis=1))
    np.testing.assert_array_equal(qz, np.concatenate([parts[p].qzeros.reshape(k // g, -1) for p in order], axis=1))
    np.testing.assert_array_equal(g_idx, gi)
    np.testing.assert_array_equal(bs, np.concatenate([np.full(widths[p], i + 1, np.float32) for i, p in enumerate(order)]))
    # the fused weight equals the concatenation of the parts' reference dequants
    w = oracle.dequant_gptq(qw.ravel(), sc.ravel(), qz.ravel(), g_idx, 4, g, k, 384)
    ref = np.concatenate([oracle.dequant_gptq(parts[p].qweight, parts[p].scales, parts[p].qzeros, gi, 4, g, k, widths[p])
                          for p in order], axis=0)
    np.testing.assert_array_equal(w, ref)
    # gate_up without g_idx although desc_act=true is configured: validate_gptq_g_idx rejects it
    with pytest.raises(fb2.Fb2Error) as e:
        ld.read_gptq_raw("model.layers.3.mlp.gate_up_proj")
    assert e.value.code == -6 and "desc_act=true but no g_idx" in str(e.value)
    with pytest.raises(fb2.Fb2Error) as e:
        ld.gptq_info("model.layers.3.mlp.down_proj")
    assert "no fusable GPTQ parts" in str(e.value)


def test_sym_canonicalisation_and_g_idx_validation(fb2, oracle, tmp_path):
    k, n, g = 256, 64, 128
    syn, t = gptq_tensors(oracle, "a", k, n, g, 5)
    _, t2 = gptq_tensors(oracle, "b", k, n, g, 6, g_idx=np.arange(k) // g)           # trivial order
    bad = (np.arange(k) // g).astype(np.int32); bad[17] = 2
    _, t3 = gptq_tensors(oracle, "c", k, n, g, 7, g_idx=bad)                          # group out of range
    _, t4 = gptq_tensors(oracle, "d", k, n, g, 8, g_idx=np.zeros(k - 8, np.int32))    # wrong length
    for d in (t2, t3, t4):
        t.update(d)
    write_safetensors(tmp_path / "model.safetensors", t)
    qcfg(tmp_path, sym=True)
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    assert ld.quant_config().sym
    qz = ld.read_gptq_raw("a")[2]
    assert (qz == 0x77777777).all() and not (syn.qzeros == 0x77777777).all()
    assert ld.gptq_info("b")["desc_act"] is False and ld.gptq_info("b")["has_g_idx"]
    with pytest.raises(fb2.Fb2Error) as e:
        ld.gptq_info("c")
    assert "g_idx[17]=2 outside expected group range 0..1" in str(e.value)
    with pytest.raises(fb2.Fb2Error) as e:
        ld.gptq_info("d")
    assert "g_idx length 248 must match K=256" in str(e.value)


def test_fusion_error_cases(fb2, oracle, tmp_path):
    k, g = 128, 128
    gi = oracle.make_desc_act_g_idx(256, g)
    t = {}
    t.update(gptq_tensors(oracle, "x.gate_proj", k, 64, g, 1, bias=np.zeros(64))[1])     # bias on one part only
    t.update(gptq_tensors(oracle, "x.up_proj", k, 64, g, 2)[1])
    t.update(gptq_tensors(oracle, "y.gate_proj", k, 64, g, 3)[1])                        # K mismatch
    t.update(gptq_tensors(oracle, "y.up_proj", 256, 64, g, 4)[1])
    t.update(gptq_tensors(oracle, "z.gate_proj", 256, 64, g, 5, g_idx=gi)[1])            # g_idx on one part only
    t.update(gptq_tensors(oracle, "z.up_proj", 256, 64, g, 6)[1])
    t.update(gptq_tensors(oracle, "w.gate_proj", 256, 64, g, 7, g_idx=gi)[1])            # differing g_idx
    t.update(gptq_tensors(oracle, "w.up_proj", 256, 64, g, 8, g_idx=gi[::-1].copy())[1])
    write_safetensors(tmp_path / "model.safetensors", t)
    qcfg(tmp_path)
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    for name, msg in (("x.gate_up_proj", "inconsistent bias presence"), ("y.gate_up_proj", "row mismatch"),
                      ("z.gate_up_proj", "missing g_idx for 'z.up_proj'"), ("w.gate_up_proj", "g_idx mismatch")):
        with pytest.raises(fb2.Fb2Error) as e:
            ld.read_gptq_raw(name)
        assert e.value.code == -6 and msg in str(e.value), (name, str(e.value))


def test_stacked_experts_host(fb2, oracle, tmp_path):
    """load_stacked_gptq_experts: per expert, gate || up concatenated along N; one shared g_idx."""
    k, inter, g, E = 256, 128, 128, 4
    gi = oracle.make_desc_act_g_idx(k, g)
    t, syn = {}, {}
    for e in range(E):
        for j, p in enumerate(("gate_proj", "up_proj")):
            s, d = gptq_tensors(oracle, f"model.layers.0.mlp.experts.{e}.{p}", k, inter, g, 0xA00DBA11 + 2 * e + j, g_idx=gi)
            syn[(e, p)] = s
            t.update(d)
        s, d = gptq_tensors(oracle, f"model.layers.0.mlp.experts.{e}.down_proj", inter, k, g, 0xCAFE0000 + e)
        syn[(e, "down_proj")] = s
        t.update(d)
    write_safetensors(tmp_path / "model.safetensors", t)
    qcfg(tmp_path)
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    fmt = "model.layers.0.mlp.experts.{e}."
    assert ld.stacked_gptq_info(fmt, ("gate_proj", "up_proj")) == (k, 2 * inter, True)
    assert ld.stacked_gptq_info(fmt, ("down_proj",), 3) == (inter, k, False)
    for e in range(E):
        qw, sc, qz, g_idx = ld.read_stacked_gptq_expert(fmt, ("gate_proj", "up_proj"), e)
        a, b = syn[(e, "gate_proj")], syn[(e, "up_proj")]
        np.testing.assert_array_equal(qw, np.concatenate([a.qweight.reshape(k // 8, -1), b.qweight.reshape(k // 8, -1)], 1))
        np.testing.assert_array_equal(sc, np.concatenate([a.scales.reshape(k // g, -1), b.scales.reshape(k // g, -1)], 1))
        np.testing.assert_array_equal(qz, np.concatenate([a.qzeros.reshape(k // g, -1), b.qzeros.reshape(k // g, -1)], 1))
        np.testing.assert_array_equal(g_idx, gi)
    with pytest.raises(fb2.Fb2Error) as e:
        ld.read_stacked_gptq_expert(fmt, ("gate_proj", "up_proj"), E)
    assert "not in index" in str(e.value)


def test_malformed_containers(fb2, tmp_path):
    def expect(msg, code=-6):
        with pytest.raises(fb2.Fb2Error) as e:
            fb2.NativeSafetensorsLoader.open(tmp_path)
        assert e.value.code == code and msg in str(e.value), str(e.value)

    expect("no safetensors files")
    p = tmp_path / "model.safetensors"
    p.write_bytes(b"abc")
    expect("smaller than the 8-byte header length")
    p.write_bytes(struct.pack("<Q", 1 << 40) + b"{}")
    expect("exceeds the file")
    p.write_bytes(struct.pack("<Q", 2) + b"[]")
    expect("not a JSON object")
    h = json.dumps({"t": {"dtype": "F32", "shape": [4], "data_offsets": [0, 12]}}).encode()
    p.write_bytes(struct.pack("<Q", len(h)) + h + b"\0" * 16)
    expect("do not match")
    h = json.dumps({"t": {"dtype": "F32", "shape": [4], "data_offsets": [0, 16]}}).encode()
    p.write_bytes(struct.pack("<Q", len(h)) + h + b"\0" * 8)
    expect("do not match")
    h = json.dumps({"t": {"dtype": "Q4", "shape": [4], "data_offsets": [0, 16]}}).encode()
    p.write_bytes(struct.pack("<Q", len(h)) + h + b"\0" * 16)
    expect("unknown dtype")
    # a good container but a bad quantize_config.json: serde rejects unknown variants / negative usize
    h = json.dumps({"t": {"dtype": "F32", "shape": [4], "data_offsets": [0, 16]}}).encode()
    p.write_bytes(struct.pack("<Q", len(h)) + h + b"\0" * 16)
    (tmp_path / "quantize_config.json").write_text('{"quant_method": "marlin", "bits": 4}')
    expect("unknown quant_method")
    (tmp_path / "quantize_config.json").write_text('{"quant_method": "gptq", "bits": 4, "group_size": -1}')
    expect("non-negative")
    (tmp_path / "quantize_config.json").write_text('{"method": "gptq", "bits": 4, "group_size": 128, "desc_act": true}')
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    assert ld.quant_config().method == "gptq" and ld.quant_config().desc_act          # `method` alias
    os.remove(tmp_path / "quantize_config.json")
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    assert ld.quant_config() is None
    with pytest.raises(fb2.Fb2Error) as e:
        ld.gptq_info("t")
    assert "requires quantize_config.json" in str(e.value)


def test_load_fails_loudly_without_gpu(fb2, oracle, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _, t = gptq_tensors(oracle, "lin", 128, 128, 128, 3)
    write_safetensors(tmp_path / "model.safetensors", t)
    qcfg(tmp_path)
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    with pytest.raises(fb2.Fb2Error) as e:
        ld.load_linear("lin")
    assert e.value.code == -3


# ------------------------------------------------------------------------------- on the GPU
@pytest.mark.gpu
def test_gpu_load_linear_fused_act_order(fb2, oracle, cuda, tmp_path):
    torch = cuda
    k, g, m = 512, 128, 3
    gi = oracle.random_balanced_g_idx(k, g, 11)
    widths = {"q_proj": 256, "k_proj": 128, "v_proj": 128}
    t = {}
    for i, (p, w) in enumerate(widths.items()):
        t.update(gptq_tensors(oracle, "m.attn." + p, k, w, g, 300 + i, g_idx=gi, bias=np.full(w, 0.25 * i, np.float16))[1])
    write_safetensors(tmp_path / "model.safetensors", t)
    qcfg(tmp_path, desc_act=True)
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    qw, sc, qz, g_idx, bs, kk, n = ld.read_gptq_raw("m.attn.qkv_proj")
    lin = ld.load_linear("m.attn.qkv_proj")
    assert (lin.in_features(), lin.out_features(), lin.is_act_order) == (k, 512, True)
    w_ref = oracle.dequant_gptq(qw.ravel(), sc.ravel(), qz.ravel(), g_idx, 4, g, k, n)
    w = lin.dequantize().cpu().numpy()
    assert np.array_equal(w.view(np.uint32), w_ref.view(np.uint32))
    x = oracle.f16_round(np.cos(np.arange(m * k, dtype=np.float32) * np.float32(0.003))).reshape(m, k)
    y = lin.forward(torch.from_numpy(x.astype(np.float16)).cuda())
    torch.cuda.synchronize()
    st = oracle.error_stats(y.float().cpu().numpy(), oracle.linear_forward(x, w_ref, bs, m))
    assert st["rel"] < 1e-2 and st["nmse"] < 1e-5, st
    # column-parallel shard of the fused layer: q | k | v are split separately
    outs = []
    for r in range(2):
        sh, _ = ld.load_linear_sharded("m.attn.qkv_proj", r, 2, 0)
        assert sh.out_features() == 256
        outs.append(sh.forward(torch.from_numpy(x.astype(np.float16)).cuda()).float().cpu().numpy())
    full = y.float().cpu().numpy()
    for r in range(2):
        ref = np.concatenate([full[:, 128 * r:128 * r + 128], full[:, 256 + 64 * r:256 + 64 * r + 64],
                              full[:, 384 + 64 * r:384 + 64 * r + 64]], axis=1)
        np.testing.assert_array_equal(outs[r], ref)


@pytest.mark.gpu
def test_gpu_load_stacked_experts(fb2, oracle, cuda, tmp_path):
    torch = cuda
    k, inter, g, E = 256, 128, 128, 4
    t = {}
    for e in range(E):
        for j, p in enumerate(("gate_proj", "up_proj")):
            t.update(gptq_tensors(oracle, f"L.experts.{e}.{p}", k, inter, g, 0xA00DBA11 + 2 * e + j)[1])
    write_safetensors(tmp_path / "model.safetensors", t)
    qcfg(tmp_path)
    ld = fb2.NativeSafetensorsLoader.open(tmp_path)
    stack, n_e, kk = ld.load_stacked_gptq_experts("L.experts.{e}.", E, ("gate_proj", "up_proj"))
    assert (n_e, kk, stack.num_experts()) == (2 * inter, k, E)
    w = stack.dequantize().cpu().numpy()
    for e in range(E):
        qw, sc, qz, _ = ld.read_stacked_gptq_expert("L.experts.{e}.", ("gate_proj", "up_proj"), e)
        ref = oracle.dequant_gptq(qw.ravel(), sc.ravel(), qz.ravel(), None, 4, g, k, n_e)
        assert np.array_equal(w[e].view(np.uint32), ref.view(np.uint32))
    # an expert-parallel rank loads its own half
    half, _, _ = ld.load_stacked_gptq_experts("L.experts.{e}.", E // 2, ("gate_proj", "up_proj"), expert_begin=2)
    np.testing.assert_array_equal(half.dequantize().cpu().numpy(), w[2:])


def test_header_json_corner_cases(fb2, tmp_path):
    """The container header is arbitrary JSON: escapes, unicode names, nested metadata, exponents, padding."""
    blob = np.arange(6, dtype=np.float32).tobytes()
    name = "ténsor \"q\"\\\n\U0001F600"
    header = {"__metadata__": {"note": "a\tb", "nested": "{\"x\": [1, 2.5e3, -1E-2, true, false, null]}"},
              name: {"dtype": "F32", "shape": [2, 3], "data_offsets": [0, 24]},
              "empty": {"dtype": "F16", "shape": [0, 4], "data_offsets": [24, 24]},
              "scalar": {"dtype": "I32", "shape": [], "data_offsets": [24, 28]}}
    for ascii_only in (True, False):
        hj = json.dumps(header, ensure_ascii=ascii_only).encode("utf-8")
        hj = b"  " + hj + b" \n\t  "
        (tmp_path / "model.safetensors").write_bytes(struct.pack("<Q", len(hj)) + hj + blob + struct.pack("<i", -7))
        ld = fb2.NativeSafetensorsLoader.open(tmp_path)
        assert sorted(ld.tensor_names()) == sorted([name, "empty", "scalar"])
        np.testing.assert_array_equal(ld.read_f32(name), np.arange(6, dtype=np.float32).reshape(2, 3))
        assert ld.tensor_info("empty") == ("F16", (0, 4), 0) and ld.read_f32("empty").shape == (0, 4)
        assert ld.tensor_info("scalar") == ("I32", (), 4) and int(ld.read_i32("scalar")) == -7
    for bad in (b'{"t": {"dtype": "F32", "shape": [1], "data_offsets": [0, 4]},}', b'{"t": {"dtype": "F32" "shape": [1]}}',
                b'{"t": {"dtype": "F32", "shape": [1.5], "data_offsets": [0, 4]}}', b'{"t": "x"}', b'{"a": [' * 100):
        (tmp_path / "model.safetensors").write_bytes(struct.pack("<Q", len(bad)) + bad + b"\0" * 8)
        with pytest.raises(fb2.Fb2Error) as e:
            fb2.NativeSafetensorsLoader.open(tmp_path)
        assert e.value.code == -6


def test_header_strings_are_strict_json(fb2, tmp_path):
    """The reference parses headers with serde_json, which rejects invalid UTF-8, raw control characters and lone
    surrogates; so does this parser (a tensor name is handed back to the caller as a C string)."""
    entry = b'{"dtype": "F32", "shape": [1], "data_offsets": [0, 4]}'
    bad_names = [b'"t\xed\xa0\x80"',          # UTF-8-encoded surrogate
                 b'"t\xff"', b'"t\xc3"',      # invalid lead byte, truncated sequence
                 b'"t\xc0\xaf"',              # overlong '/'
                 b'"t\xf4\x90\x80\x80"',      # above U+10FFFF
                 b'"t\x01"', b'"t\nx"',       # raw control characters
                 b'"\\ud800"', b'"\\udc00"', b'"\\ud800\\u0041"']   # lone / mismatched surrogates
    for nm in bad_names:
        hj = b"{" + nm + b": " + entry + b"}"
        (tmp_path / "model.safetensors").write_bytes(struct.pack("<Q", len(hj)) + hj + b"\0" * 8)
        with pytest.raises(fb2.Fb2Error) as e:
            fb2.NativeSafetensorsLoader.open(tmp_path)
        assert e.value.code == -6, nm
    good = '{"\\ud83d\\ude00 é\\u00e9": '.encode("utf-8") + entry + b"}"
    (tmp_path / "model.safetensors").write_bytes(struct.pack("<Q", len(good)) + good + b"\0" * 8)
    assert fb2.NativeSafetensorsLoader.open(tmp_path).tensor_names() == ["\U0001F600 éé"]


def test_fuzzed_containers_never_crash_the_parser():
    """scripts/fuzz_checkpoint.py: 300 mutated containers / config files / shard indices are each either loaded or
    rejected with a status; a crash of the C++ ingest kills the child process and fails here."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "fuzz_checkpoint.py"), "300", "11"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "no crash" in out.stdout, (out.returncode, out.stdout[-500:], out.stderr[-2000:])
