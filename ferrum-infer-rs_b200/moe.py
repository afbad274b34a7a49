"""BackendMoeFused surface (crates/ferrum-kernels/src/backend/capabilities.rs:305-723) and the
fused MoE FFN (crates/ferrum-models/src/moe/dispatch.rs:1558-2192) over the C ABI.
All tensors are CUDA tensors; every call appends to the current stream."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._ffi import check, lib
from .gptq import GptqExpertStack, _stream


def _i32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not (t.is_cuda and t.dtype == torch.int32 and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous CUDA int32 tensor")
    return t


def route_topk_softmax(logits: torch.Tensor, top_k: int, norm_topk_prob: bool,
                       out_ids: Optional[torch.Tensor] = None, out_weights: Optional[torch.Tensor] = None
                       ) -> Tuple[torch.Tensor, torch.Tensor]:
    """capabilities.rs:337; logits [batch, E] f16 -> ids [batch, K] i32, weights [batch, K] f32."""
    batch, e = logits.shape
    if not (logits.is_cuda and logits.dtype == torch.float16 and logits.is_contiguous()):
        raise ValueError("logits must be a contiguous CUDA float16 tensor")
    if out_ids is None:
        out_ids = torch.empty((batch, top_k), dtype=torch.int32, device=logits.device)
    if out_weights is None:
        out_weights = torch.empty((batch, top_k), dtype=torch.float32, device=logits.device)
    check(lib().fb2_moe_route_topk_softmax(_stream(), logits.data_ptr(), out_ids.data_ptr(),
                                           out_weights.data_ptr(), batch, e, top_k, 1 if norm_topk_prob else 0))
    return out_ids, out_weights


def moe_build_pairs_by_token(expert_ids: torch.Tensor, num_experts: int, top_k: int):
    """capabilities.rs:407 -> (pairs_by_token [pairs], packed_token_idx [pairs], expert_offsets [E+1])."""
    ids = _i32(expert_ids, "expert_ids").reshape(-1)
    pairs = ids.numel()
    dev = ids.device
    pbt = torch.empty(pairs, dtype=torch.int32, device=dev)
    pti = torch.empty(pairs, dtype=torch.int32, device=dev)
    offs = torch.empty(num_experts + 1, dtype=torch.int32, device=dev)
    check(lib().fb2_moe_build_pairs_by_token(_stream(), ids.data_ptr(), pbt.data_ptr(), pti.data_ptr(),
                                             offs.data_ptr(), pairs, num_experts, top_k))
    return pbt, pti, offs


def moe_align_block_size(expert_ids: torch.Tensor, num_experts: int, block_size: int,
                         sorted_max_size: Optional[int] = None, pair_ids: bool = False):
    """capabilities.rs:427,447 -> (sorted_token_ids [cap], block_ids [cap/block + 1], total [1])."""
    ids = _i32(expert_ids, "expert_ids").reshape(-1)
    pairs = ids.numel()
    if sorted_max_size is None:
        sorted_max_size = pairs + num_experts * block_size          # dispatch.rs:1865
    dev = ids.device
    sorted_ids = torch.empty(sorted_max_size, dtype=torch.int32, device=dev)
    block_ids = torch.full((sorted_max_size // block_size + 1,), -1, dtype=torch.int32, device=dev)
    total = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib().fb2_moe_align_block_size(_stream(), ids.data_ptr(), sorted_ids.data_ptr(),
                                         block_ids.data_ptr(), total.data_ptr(), pairs, num_experts,
                                         block_size, sorted_max_size, 1 if pair_ids else 0))
    return sorted_ids, block_ids, total


def moe_combine(packed_down: torch.Tensor, pairs_by_token: torch.Tensor, pair_weights: torch.Tensor,
                out: torch.Tensor, batch: int, hidden: int, top_k: int) -> torch.Tensor:
    """capabilities.rs:684: out is OVERWRITTEN."""
    check(lib().fb2_moe_combine(_stream(), packed_down.data_ptr(), _i32(pairs_by_token, "pairs_by_token").data_ptr(),
                                pair_weights.data_ptr(), out.data_ptr(), batch, hidden, top_k))
    return out


def weighted_sum_batched(slots: torch.Tensor, weights: torch.Tensor, out: torch.Tensor, batch: int,
                         top_k: int, hidden: int) -> torch.Tensor:
    """capabilities.rs:565: slots [batch, top_k, hidden]."""
    check(lib().fb2_moe_weighted_sum_batched(_stream(), slots.data_ptr(), weights.data_ptr(),
                                             out.data_ptr(), batch, top_k, hidden))
    return out


def fused_silu_mul_split(gate_up: torch.Tensor, out: torch.Tensor, tokens: int, im: int) -> torch.Tensor:
    check(lib().fb2_fused_silu_mul_split(_stream(), gate_up.data_ptr(), out.data_ptr(), tokens, im))
    return out


def fused_gelu_tanh_mul_split(gate_up: torch.Tensor, out: torch.Tensor, tokens: int, im: int) -> torch.Tensor:
    check(lib().fb2_fused_gelu_tanh_mul_split(_stream(), gate_up.data_ptr(), out.data_ptr(), tokens, im))
    return out


class MoeFfn:
    """One layer's MoE FFN: moe_forward_bucketed (dispatch.rs:1558) with gather, SwiGLU and
    combine fused around the two grouped GEMMs.  Owns its workspace (graph-capturable)."""

    def __init__(self, gate_up: GptqExpertStack, down: GptqExpertStack, top_k: int,
                 norm_topk_prob: bool, max_batch: int):
        self.gate_up, self.down = gate_up, down
        self.top_k, self.norm, self.max_batch = top_k, norm_topk_prob, max_batch
        self.hidden = gate_up.k()
        nbytes = int(lib().fb2_moe_ffn_workspace_bytes(gate_up._h, down._h, max_batch, top_k))
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{gate_up.device}")

    def forward(self, x: torch.Tensor, router_logits: torch.Tensor, out: Optional[torch.Tensor] = None
                ) -> torch.Tensor:
        batch = x.shape[0]
        if batch > self.max_batch:
            raise ValueError(f"batch {batch} > max_batch {self.max_batch}")
        if out is None:
            out = torch.empty((batch, self.hidden), dtype=torch.float16, device=x.device)
        check(lib().fb2_moe_ffn_forward(self.gate_up._h, self.down._h, _stream(), x.data_ptr(),
                                        router_logits.data_ptr(), out.data_ptr(), batch, self.top_k,
                                        1 if self.norm else 0, self.workspace.data_ptr(),
                                        self.workspace.numel()))
        return out

    __call__ = forward
