"""Chunked dense launches (launch_tc_gemm): kept in its own, last-sorting file because it runs the library in a
subprocess with a debug environment variable."""
import pytest

pytestmark = pytest.mark.gpu

_CHUNK_SCRIPT = r"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import oracle
fb2 = importlib.import_module("ferrum-infer-rs_b200")
k, n, g, m = 512, 256, 128, 700
for act in (False, True):
    syn = oracle.make_synthetic(k, n, g, 0xBADC0DE, f16_scales=True)
    gi = oracle.random_balanced_g_idx(k, g, 77) if act else None
    lin = fb2.GptqLinear.from_raw(syn.qweight, syn.scales, syn.qzeros, gi, None, 4, g, k, n)
    fb2.pregrow_scratch(0, m)
    x = oracle.f16_round(oracle.lcg_f32_array(0xC0FFEE, m * k, -1.0, 1.0)).reshape(m, k)
    fb2.reset_launch_count()
    y = lin.forward(torch.from_numpy(x.astype(np.float16)).cuda())
    torch.cuda.synchronize()
    launches = fb2.launch_count() - (1 if act else 0)          # minus the act-order gather
    assert launches >= 2, f"expected the batch to be cut into several GEMM launches, saw {launches}"
    w = oracle.dequant_gptq(syn.qweight, syn.scales, syn.qzeros, gi, 4, g, k, n)
    st = oracle.error_stats(y.float().cpu().numpy(), oracle.linear_forward(x, w, None, m, threads=8))
    assert st["rel"] < 1e-2 and st["nmse"] < 1e-5, st
print("chunked ok")
"""


def test_chunked_dense_launch(fb2, oracle, cuda):
    """Dense batches whose units x CTAs would overflow the kernels' 32-bit bookkeeping are cut into several
    launches of whole row blocks (launch_tc_gemm).  FB2_DEBUG_MAX_UNIT_CTAS (read once per process, hence the
    subprocess) lowers the limit so that 700 rows reach that path; every row must still match the oracle."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FB2_DEBUG_MAX_UNIT_CTAS="100")
    out = subprocess.run([sys.executable, "-c", _CHUNK_SCRIPT], cwd=root, env=env, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and "chunked ok" in out.stdout, out.stdout + out.stderr
