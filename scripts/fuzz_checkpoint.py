"""Mutation fuzzer for the safetensors / quantize_config ingest (csrc/checkpoint.cu).  Every case runs in this process:
a crash (SIGSEGV / abort) of the C++ parser kills the script and names the seed; a clean run means every mutated
container was either loaded or rejected with a status code.  usage: python scripts/fuzz_checkpoint.py [cases] [seed]"""
import importlib
import json
import os
import shutil
import struct
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
fb2 = importlib.import_module("ferrum-infer-rs_b200")
from test_checkpoint import qcfg, write_safetensors  # noqa: E402  (the independent container writer)

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
k, n, g = 256, 128, 128
rng0 = np.random.default_rng(7)
base = {
    "model.layers.0.self_attn.q_proj.qweight": rng0.integers(-2**31, 2**31, (k // 8, n), dtype=np.int64).astype(np.int32),
    "model.layers.0.self_attn.q_proj.scales": rng0.uniform(0.01, 0.1, (k // g, n)).astype(np.float16),
    "model.layers.0.self_attn.q_proj.qzeros": rng0.integers(-2**31, 2**31, (k // g, n // 8), dtype=np.int64).astype(np.int32),
    "model.layers.0.self_attn.q_proj.g_idx": (np.arange(k) // g).astype(np.int32),
    "model.norm.weight": rng0.uniform(-1, 1, (n,)).astype(np.float32),
}
work = tempfile.mkdtemp(prefix="fb2fuzz_")
good = os.path.join(work, "good.safetensors")
write_safetensors(good, base)
blob = open(good, "rb").read()
hlen = struct.unpack("<Q", blob[:8])[0]
ok = rejected = 0
for case in range(cases):
    rng = np.random.default_rng(seed0 * 1000003 + case)
    d = os.path.join(work, f"c{case % 8}")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    b = bytearray(blob)
    kind = rng.integers(0, 7)
    if kind == 0:                                   # flip bytes inside the JSON header
        for _ in range(rng.integers(1, 6)):
            b[8 + rng.integers(0, hlen)] = rng.integers(0, 256)
    elif kind == 1:                                 # lie about the header length
        lens = [0, 1, 7, hlen - 1, hlen + 1, len(b), len(b) + 8, 2**31, 2**63, 2**64 - 1]
        b[:8] = struct.pack("<Q", lens[rng.integers(0, len(lens))])
    elif kind == 2:                                 # truncate the file
        b = b[:rng.integers(0, len(b))]
    elif kind == 3:                                 # rewrite data_offsets / shape with hostile numbers
        h = json.loads(blob[8:8 + hlen])
        name = list(h)[rng.integers(0, len(h))]
        big = [0, 1, 2**31 - 1, 2**31, 2**32, 2**53, 2**63 - 1, -1, -2**31]
        pick = lambda: big[rng.integers(0, len(big))]  # noqa: E731
        if rng.integers(0, 2):
            h[name]["data_offsets"] = [pick(), pick()]
        else:
            h[name]["shape"] = [pick() for _ in range([0, 1, 2, 3, 4, 8, 9, 40][rng.integers(0, 8)])]
        hj = json.dumps(h).encode()
        b = bytearray(struct.pack("<Q", len(hj)) + hj + blob[8 + hlen:])
    elif kind == 4:                                 # structurally odd JSON
        alts = [b"[]", b"{}", b"null", b'{"a":', b'{"a":{"dtype":"F32","shape":[1],"data_offsets":[0,4]},"a":{}}',
                b'{"x":{"dtype":"F32","shape":"no","data_offsets":[0,4]}}', b'{"x":{"dtype":7,"shape":[1],"data_offsets":[0,4]}}',
                b'{"x":{"dtype":"F32","shape":[1],"data_offsets":[4,0]}}', b'{' + b'"k":{' * 2000, b'"' + b"\\" * 999,
                b'{"\\ud800":{"dtype":"F32","shape":[1],"data_offsets":[0,4]}}', b'{"x":' + b"[" * 5000 + b"]" * 5000 + b"}"]
        hj = alts[rng.integers(0, len(alts))]
        b = bytearray(struct.pack("<Q", len(hj)) + hj + b"\0" * 64)
    elif kind == 5:                                 # hostile quantize_config.json / config.json next to a good file
        txt = [b"", b"{", b"[]", b'{"bits":"four"}', b'{"bits":4,"group_size":-1}', b'{"bits":1e400,"group_size":1e400}',
               b'{"quant_method":null,"bits":4,"group_size":128,"desc_act":"yes","sym":[]}',
               b'{"bits":4,"group_size":18446744073709551616}', b"\xff\xfe\x00", b'{"bits":4,"group_size":128' + b" " * 70000][rng.integers(0, 10)]
        open(os.path.join(d, ["quantize_config.json", "config.json"][rng.integers(0, 2)]), "wb").write(
            txt if rng.integers(0, 2) else b'{"quantization_config":' + txt + b"}")
    else:                                           # hostile shard index
        idx = [b"{}", b'{"weight_map":[]}', b'{"weight_map":{"a":5}}', b'{"weight_map":{"a":"../../etc/passwd"}}',
               b'{"weight_map":{"a":"missing.safetensors"}}', b'{"weight_map":{"a":"model.safetensors","b":"model.safetensors"}}'][rng.integers(0, 6)]
        open(os.path.join(d, "model.safetensors.index.json"), "wb").write(idx)
    open(os.path.join(d, "model.safetensors"), "wb").write(bytes(b))
    with open(os.path.join(work, "current_case.txt"), "w") as f:      # a crash leaves the culprit's id behind
        f.write(f"seed {seed0} case {case} kind {kind}\n")
    try:
        ld = fb2.NativeSafetensorsLoader.open(d)
    except fb2.Fb2Error:
        rejected += 1
        continue
    soft = (fb2.Fb2Error, ValueError, MemoryError, ZeroDivisionError, OverflowError, UnicodeDecodeError, AttributeError, TypeError)
    try:
        try:
            ld.quant_config()
        except soft:
            pass
        for name in ld.tensor_names():
            try:
                _, shape, nbytes = ld.tensor_info(name)
                if nbytes <= (1 << 26) and int(np.prod(shape, dtype=object)) <= (1 << 24):
                    ld.read_f32(name)
            except soft:
                pass
        for lin in ("model.layers.0.self_attn.q_proj", "model.layers.0.self_attn.qkv_proj", "model.norm"):
            try:
                info = ld.gptq_info(lin)
                if info["k"] * info["n"] <= (1 << 24):
                    ld.read_gptq_raw(lin)
            except soft:
                pass
        ok += 1
    finally:
        ld.close()
print(f"{cases} cases: {ok} opened, {rejected} rejected at open, no crash")
shutil.rmtree(work, ignore_errors=True)
