"""ctypes binding of libfb2_w4a16.so (include/fb2_w4a16.h).

This is the *reference-side binding stub* exercised from Python: the same symbols a Rust
`extern "C"` block binds.  There is no fallback of any kind: if the
shared library is missing the import fails, and if no sm_100 device is usable every
computing entry point returns FB2_ERR_CUDA, which is raised as `Fb2Error`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FB2_LIB selects another build of the same library (A/B measurements, scripts/build_variant.sh)
LIB_PATH = os.environ.get("FB2_LIB") or os.path.join(_HERE, "libfb2_w4a16.so")

FB2_OK = 0
ERR_NAMES = {-1: "FB2_ERR_SHAPE", -2: "FB2_ERR_UNSUPPORTED", -3: "FB2_ERR_CUDA",
             -4: "FB2_ERR_NCCL", -5: "FB2_ERR_ARG", -6: "FB2_ERR_MODEL", -7: "FB2_ERR_IO"}


class Fb2Error(RuntimeError):
    """A non-zero status from the C ABI.  `.code` is the FB2_ERR_* value; the reference maps
    UNSUPPORTED -> FerrumError::unsupported and MODEL/SHAPE -> FerrumError::model."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


class QuantConfigC(C.Structure):
    _fields_ = [("present", C.c_int), ("method", C.c_int), ("bits", C.c_uint32), ("group_size", C.c_size_t),
                ("desc_act", C.c_int), ("sym", C.c_int)]


class GptqInfoC(C.Structure):
    _fields_ = [("k", C.c_size_t), ("n", C.c_size_t), ("has_g_idx", C.c_int), ("has_bias", C.c_int),
                ("desc_act", C.c_int), ("n_parts", C.c_int), ("part_n", C.c_size_t * 4)]


class Dispatch(C.Structure):
    _fields_ = [("expert", C.c_int64), ("in_row", C.c_int64), ("out_row", C.c_int64),
                ("m", C.c_int64)]


_vp, _sz, _i32, _u32, _u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64

# name -> (restype, argtypes); kept in one table so tests can check that every symbol the
# header declares is exported.
SIGNATURES = {
    "fb2_version": (C.c_char_p, []),
    "fb2_last_error": (C.c_char_p, []),
    "fb2_launch_count": (_u64, []),
    "fb2_reset_launch_count": (None, []),
    "fb2_gptq_linear_create": (_i32, [_vp, _vp, _vp, _vp, _vp, _u32, _sz, _sz, _sz, _i32, C.POINTER(_vp)]),
    "fb2_gptq_linear_destroy": (None, [_vp]),
    "fb2_gptq_linear_in_features": (_sz, [_vp]),
    "fb2_gptq_linear_out_features": (_sz, [_vp]),
    "fb2_gptq_linear_is_act_order": (_i32, [_vp]),
    "fb2_gptq_linear_kernel_name": (C.c_char_p, [_vp, _sz]),
    "fb2_gptq_linear_forward": (_i32, [_vp, _vp, _vp, _vp, _sz]),
    "fb2_gptq_linear_forward_host": (_i32, [_vp, _vp, _vp, _sz]),
    "fb2_gptq_linear_forward_host_batch": (_i32, [_vp, _sz, _vp, _vp, _sz]),
    "fb2_gptq_linear_forward_host_batch_m": (_i32, [_vp, _sz, _vp, _vp, _vp]),
    "fb2_gptq_linear_dequant_f32": (_i32, [_vp, _vp, _vp]),
    "fb2_gptq_linear_dequant_f16": (_i32, [_vp, _vp, _vp]),
    "fb2_gemm_algorithmic_bytes": (_u64, [_sz, _sz, _sz, _sz, _i32]),
    "fb2_gemm_flops": (_u64, [_sz, _sz, _sz]),
    "fb2_pregrow_scratch": (_i32, [_i32, _sz]),
    "fb2_gptq_stack_create": (_i32, [_vp, _vp, _vp, _vp, _u32, _sz, _sz, _sz, _sz, _i32, C.POINTER(_vp)]),
    "fb2_gptq_stack_destroy": (None, [_vp]),
    "fb2_gptq_stack_n_per_expert": (_sz, [_vp]),
    "fb2_gptq_stack_k": (_sz, [_vp]),
    "fb2_gptq_stack_num_experts": (_sz, [_vp]),
    "fb2_gptq_stack_zero_workspace": (_i32, [_vp, _vp]),
    "fb2_gptq_stack_make_expert_linear": (_i32, [_vp, _sz, _sz, _vp, C.POINTER(_vp)]),
    "fb2_gptq_stack_gemm_batched": (_i32, [_vp, _vp, _vp, C.POINTER(Dispatch), _sz, _vp, _sz]),
    "fb2_gptq_stack_gemm_routed": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _sz, _sz, _sz]),
    "fb2_gptq_stack_gemm_routed_silu": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _sz, _sz, _sz]),
    "fb2_gptq_stack_dequant_f32": (_i32, [_vp, _vp, _vp]),
    "fb2_moe_route_topk_softmax": (_i32, [_vp, _vp, _vp, _vp, _sz, _sz, _sz, _i32]),
    "fb2_moe_build_pairs_by_token": (_i32, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _sz]),
    "fb2_moe_align_block_size": (_i32, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _sz, _sz, _i32]),
    "fb2_moe_combine": (_i32, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _sz]),
    "fb2_moe_weighted_sum_batched": (_i32, [_vp, _vp, _vp, _vp, _sz, _sz, _sz]),
    "fb2_fused_silu_mul_split": (_i32, [_vp, _vp, _vp, _sz, _sz]),
    "fb2_fused_gelu_tanh_mul_split": (_i32, [_vp, _vp, _vp, _sz, _sz]),
    "fb2_fused_add_rms_norm": (_i32, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, C.c_float, _vp]),
    "fb2_gptq_linear_act_order_perm": (_i32, [_vp, C.POINTER(_vp)]),
    "fb2_gptq_linear_forward_sorted": (_i32, [_vp, _vp, _vp, _vp, _sz]),
    "fb2_gptq_linear_forward_swiglu": (_i32, [_vp, _vp, _vp, _vp, _sz, _i32]),
    "fb2_argmax_rows_f16": (_i32, [_vp, _vp, _sz, _sz, _vp]),
    "fb2_moe_ffn_workspace_bytes": (_sz, [_vp, _vp, _sz, _sz]),
    "fb2_moe_ffn_forward": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _sz, _i32, _vp, _sz]),
    "fb2_gptq_shard_dims": (_i32, [_sz, _sz, _sz, _i32, _i32, _i32, _vp, _i32, C.POINTER(_sz), C.POINTER(_sz)]),
    "fb2_gptq_shard_host": (_i32, [_vp, _vp, _vp, _vp, _vp, _sz, _sz, _sz, _i32, _i32, _i32, _vp, _i32,
                                   _vp, _vp, _vp, _vp, _vp]),
    "fb2_gptq_linear_create_sharded": (_i32, [_vp, _vp, _vp, _vp, _vp, _u32, _sz, _sz, _sz, _i32, _i32, _i32, _i32,
                                              _vp, _i32, _vp, C.POINTER(_vp)]),
    "fb2_comm_get_unique_id": (_i32, [_vp]),
    "fb2_comm_init_rank": (_i32, [_vp, _i32, _i32, _i32, C.POINTER(_vp)]),
    "fb2_comm_destroy": (None, [_vp]),
    "fb2_comm_world_size": (_i32, [_vp]),
    "fb2_comm_rank": (_i32, [_vp]),
    "fb2_all_reduce_f16": (_i32, [_vp, _vp, _vp, _sz]),
    "fb2_all_gather_f16": (_i32, [_vp, _vp, _vp, _vp, _sz]),
    "fb2_broadcast_f16": (_i32, [_vp, _vp, _vp, _sz, _i32]),
    "fb2_all_to_all_f16": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fb2_gptq_linear_forward_allreduce": (_i32, [_vp, _vp, _vp, _vp, _vp, _sz]),
    "fb2_comm_enable_p2p": (_i32, [_vp, _vp, _sz]),
    "fb2_moe_ep_workspace_bytes": (_sz, [_vp, _vp, _sz, _sz, _i32]),
    "fb2_moe_ep_ffn_forward": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _sz, _i32, _vp, _sz]),
    "fb2_ckpt_open": (_i32, [C.c_char_p, C.POINTER(_vp)]),
    "fb2_ckpt_close": (None, [_vp]),
    "fb2_ckpt_quant_config": (_i32, [_vp, C.POINTER(QuantConfigC)]),
    "fb2_ckpt_num_tensors": (_sz, [_vp]),
    "fb2_ckpt_tensor_name": (C.c_char_p, [_vp, _sz]),
    "fb2_ckpt_has_tensor": (_i32, [_vp, C.c_char_p]),
    "fb2_ckpt_tensor_info": (_i32, [_vp, C.c_char_p, C.POINTER(_i32), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "fb2_ckpt_read_f32": (_i32, [_vp, C.c_char_p, _vp, _sz]),
    "fb2_ckpt_read_i32": (_i32, [_vp, C.c_char_p, _vp, _sz]),
    "fb2_ckpt_gptq_info": (_i32, [_vp, C.c_char_p, C.POINTER(GptqInfoC)]),
    "fb2_ckpt_read_gptq": (_i32, [_vp, C.c_char_p, _vp, _vp, _vp, _vp, _vp]),
    "fb2_ckpt_load_gptq_linear": (_i32, [_vp, C.c_char_p, _i32, C.POINTER(_vp)]),
    "fb2_ckpt_load_gptq_linear_sharded": (_i32, [_vp, C.c_char_p, _i32, _i32, _i32, _i32, _vp, C.POINTER(_vp)]),
    "fb2_ckpt_stacked_gptq_info": (_i32, [_vp, C.c_char_p, _vp, _sz, _sz, C.POINTER(_sz), C.POINTER(_sz),
                                          C.POINTER(_i32)]),
    "fb2_ckpt_read_stacked_gptq_expert": (_i32, [_vp, C.c_char_p, _vp, _sz, _sz, _vp, _vp, _vp, _vp]),
    "fb2_ckpt_load_stacked_gptq_experts": (_i32, [_vp, C.c_char_p, _sz, _sz, _vp, _sz, _i32, C.POINTER(_vp),
                                                  C.POINTER(_sz), C.POINTER(_sz)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the product library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (make -C ferrum-infer-rs_b200/csrc).  There is no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != FB2_OK:
        raise Fb2Error(rc, lib().fb2_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(lib().fb2_launch_count())


def reset_launch_count() -> None:
    lib().fb2_reset_launch_count()
