"""ferrum-infer-rs_b200: host-side mirror of ferrum-infer-rs's GPTQ-Int4 W4A16 operator
interface over the B200-native C ABI (libfb2_w4a16.so).

The names follow the reference crates this path lives in:
  ferrum-quantization  -> QuantConfig, GptqLinear, StackedExpertLinear, load_gptq_stacked
  ferrum-kernels       -> MarlinExpertStack surface (GptqExpertStack), BackendMoeFused ops
  ferrum-models::moe   -> moe_forward_bucketed (moe_ffn_forward), route
torch is used only to own device memory and streams.  Import as
`importlib.import_module("ferrum-infer-rs_b200")` (the directory name is the reference
repo's name, which is not a Python identifier).
"""
from ._ffi import Fb2Error, lib, launch_count, reset_launch_count, LIB_PATH  # noqa: F401
from .gptq import (QuantConfig, GptqLinear, StackedExpertLinear, GptqExpertStack,  # noqa: F401
                   gemm_algorithmic_bytes, gemm_flops, pregrow_scratch, forward_host_batch, HostBatch,
                   fused_add_rms_norm, argmax_rows)
from . import moe  # noqa: F401

__version__ = "0.1.0"
from .checkpoint import NativeSafetensorsLoader  # noqa: F401,E402
