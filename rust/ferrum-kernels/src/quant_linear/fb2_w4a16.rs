//! `Linear<CudaBackend>` / `MarlinExpertStack<CudaBackend>` / `BackendQuantMarlin` over `libfb2_w4a16.so`.
//!
//! Drop this file into the reference at `crates/ferrum-kernels/src/quant_linear/fb2_w4a16.rs`, add
//! `fb2-w4a16-sys` (this repository, `rust/fb2-w4a16-sys`) as an optional dependency behind a cargo
//! feature `fb2-w4a16`, and `#[cfg(feature = "fb2-w4a16")] pub mod fb2_w4a16;` in `quant_linear/mod.rs`.
//! With the feature on, `backend/cuda/quant.rs` forwards its three `BackendQuantMarlin` methods to the
//! free functions at the bottom of this file instead of the Marlin / vLLM-Marlin / Triton-PTX stores
//! (`quant.rs:598-760`); nothing above the trait boundary (ferrum-quantization's `GptqLinear`,
//! ferrum-models' `ExpertStack`, ferrum-engine) changes.  One visibility edit is needed:
//! `current_device_ordinal` (backend/cuda/mod.rs:210) from `pub(super)` to `pub(crate)`, because the
//! two load functions have no `Context` argument to read the ordinal from.
//!
//! Interfaces replaced, by reference file:line —
//!   `Linear::forward`                       linear.rs:123        (CUDA impl quant_linear/cuda_marlin.rs:83)
//!   `MarlinExpertStack`                     marlin_expert_stack.rs:31-109 (CUDA impl cuda_marlin_stack.rs:47-220)
//!   `BackendQuantMarlin::load_gptq`         backend/capabilities.rs:136   (CUDA impl backend/cuda/quant.rs:612)
//!   `BackendQuantMarlin::load_gptq_stacked` backend/capabilities.rs:180
//!   `pregrow_marlin_gather_scratch`         backend/capabilities.rs:212
//!
//! NOT COMPILED in the repository this file ships in (no Rust toolchain in that image).  It is written
//! against the reference's trait signatures as read at the lines above and against `fb2-w4a16-sys`,
//! whose `extern "C"` block is generated from the C header and checked against it by the test-suite.
#![cfg(all(feature = "cuda", feature = "fb2-w4a16"))]

use crate::backend::cuda::{CudaBackend, CudaState};
use crate::backend::CudaBuf;
use crate::marlin_expert_stack::MarlinExpertStack;
use crate::Linear;
use cudarc::driver::{DevicePtr, DevicePtrMut};
use fb2_w4a16_sys as fb2;
use ferrum_types::{FerrumError, Result};
use std::os::raw::c_void;
use std::sync::atomic::{AtomicUsize, Ordering};
use std::sync::Arc;

/// Widest `in_features` loaded so far; turns the reference's element-count scratch request into rows.
static WIDEST_K: AtomicUsize = AtomicUsize::new(0);

/// `Fb2Error` -> the `FerrumError` kind the reference raises for the same condition: unsupported
/// quantisation (`quant.rs:623`), malformed tensors (`native_safetensors.rs:1308-1322`), I/O.
fn ferr(what: &str, e: fb2::Fb2Error) -> FerrumError {
    match e {
        fb2::Fb2Error::Unsupported(m) => FerrumError::unsupported(format!("{what}: {m}")),
        fb2::Fb2Error::Io(m) => FerrumError::io(format!("{what}: {m}")),
        other => FerrumError::model(format!("{what}: {other}")),
    }
}

#[inline]
fn raw_stream(ctx: &CudaState) -> fb2::Stream {
    ctx.stream.cu_stream() as *mut c_void
}

// ---------------------------------------------------------------------------------------- Linear
/// One GPTQ-Int4 projection (dense, or a column-slice view of an expert stack).
pub struct Fb2GptqLinear {
    inner: fb2::GptqLinear,
}

impl Fb2GptqLinear {
    pub fn new(inner: fb2::GptqLinear) -> Self {
        Self { inner }
    }
    pub fn handle(&self) -> &fb2::GptqLinear {
        &self.inner
    }
}

impl Linear<CudaBackend> for Fb2GptqLinear {
    fn in_features(&self) -> usize {
        self.inner.in_features()
    }
    fn out_features(&self) -> usize {
        self.inner.out_features()
    }

    /// `forward` has no `Result` in the trait (linear.rs:123); the reference's CUDA impl panics on a
    /// kernel error (cuda_marlin.rs:170,258) and so does this one.  The act-order row gather, the bias
    /// add and the kernel choice (split-K decode vs tcgen05 prefill) all happen inside the library.
    fn forward(&self, ctx: &mut CudaState, input: &CudaBuf, out: &mut CudaBuf, m: usize) {
        let stream = ctx.stream.clone();
        let (x, _xg) = input.as_f16().device_ptr(&stream);
        let (y, _yg) = out.as_f16_mut().device_ptr_mut(&stream);
        let rc = unsafe { self.inner.forward(raw_stream(ctx), x as fb2::DevF16, y as fb2::DevF16Mut, m) };
        if let Err(e) = rc {
            panic!("fb2 GPTQ forward (m={m}, K={}, N={}): {e}", self.in_features(), self.out_features());
        }
    }
}

// ----------------------------------------------------------------------------------------- stack
/// `MarlinExpertStack<CudaBackend>` over one `fb2_stack_t` (all experts of one role, tile-stripe layout).
pub struct Fb2ExpertStack {
    inner: Arc<fb2::ExpertStack>,
}

impl Fb2ExpertStack {
    pub fn handle(&self) -> &Arc<fb2::ExpertStack> {
        &self.inner
    }
}

impl MarlinExpertStack<CudaBackend> for Fb2ExpertStack {
    fn n_per_expert(&self) -> usize {
        self.inner.n_per_expert()
    }
    fn k(&self) -> usize {
        self.inner.k()
    }
    fn num_experts(&self) -> usize {
        self.inner.num_experts()
    }

    /// The routed single-launch path serves every zero-point and both id conventions, and the bucketed
    /// path is served as well, so the caller is free to choose (`dispatch.rs:1558` picks by batch).
    fn requires_vllm_moe(&self) -> bool {
        false
    }

    fn as_any(&self) -> &dyn std::any::Any {
        self
    }

    /// The stream-K fix-up slots clean themselves; kept as a cheap call so `moe_forward_bucketed`
    /// needs no feature test.
    fn zero_workspace(&self, ctx: &mut CudaState) -> Result<()> {
        self.inner.zero_workspace(raw_stream(ctx)).map_err(|e| ferr("fb2 zero_workspace", e))
    }

    fn gemm_phase_batched(&self, ctx: &mut CudaState, input: &CudaBuf, dispatches: &[(usize, usize, usize, usize)],
                          output: &mut CudaBuf, k: usize) -> Result<()> {
        let stream = ctx.stream.clone();
        let (x, _xg) = input.as_f16().device_ptr(&stream);
        let (y, _yg) = output.as_f16_mut().device_ptr_mut(&stream);
        unsafe { self.inner.gemm_phase_batched(raw_stream(ctx), x as fb2::DevF16, dispatches, y as fb2::DevF16Mut, k) }
            .map_err(|e| ferr("fb2 gemm_phase_batched", e))
    }

    #[allow(clippy::too_many_arguments)]
    fn gemm_phase_vllm(&self, ctx: &mut CudaState, input: &CudaBuf, sorted_token_ids: &CudaBuf, expert_ids: &CudaBuf,
                       num_tokens_past_padded: &CudaBuf, output: &mut CudaBuf, prob_m: usize, moe_block_size: usize,
                       top_k: usize) -> Result<()> {
        let stream = ctx.stream.clone();
        let sorted = sorted_token_ids.as_i32();
        let (x, _g0) = input.as_f16().device_ptr(&stream);
        let (s, _g1) = sorted.device_ptr(&stream);
        let (e, _g2) = expert_ids.as_i32().device_ptr(&stream);
        let (t, _g3) = num_tokens_past_padded.as_i32().device_ptr(&stream);
        let (y, _g4) = output.as_f16_mut().device_ptr_mut(&stream);
        unsafe {
            self.inner.gemm_phase_vllm(raw_stream(ctx), x as fb2::DevF16, s as *const i32, e as *const i32, t as *const i32,
                                       y as fb2::DevF16Mut, prob_m, moe_block_size, top_k, sorted.len())
        }
        .map_err(|e| ferr("fb2 gemm_phase_vllm", e))
    }

    fn make_expert_linear(self: Arc<Self>, expert_offset: usize, expert_n: usize, bias_host: Option<&[f32]>)
                          -> Result<Box<dyn Linear<CudaBackend> + Send + Sync>> {
        let view = self.inner.make_expert_linear(expert_offset, expert_n, bias_host)
            .map_err(|e| ferr("fb2 make_expert_linear", e))?;
        Ok(Box::new(Fb2GptqLinear::new(view)))
    }
}

// ------------------------------------------------------------- BackendQuantMarlin entry points
/// Body for `<CudaBackend as BackendQuantMarlin>::load_gptq` (capabilities.rs:136).
#[allow(clippy::too_many_arguments)]
pub fn load_gptq(qweight: &[i32], scales: &[f32], qzeros: &[i32], g_idx: Option<&[i32]>, bias_host: Option<&[f32]>,
                 bits: u32, group_size: usize, k: usize, n: usize) -> Result<Box<dyn Linear<CudaBackend> + Send + Sync>> {
    let device = crate::backend::cuda::current_device_ordinal() as i32;
    let lin = fb2::GptqLinear::load_gptq(qweight, scales, qzeros, g_idx, bias_host, bits, group_size, k, n, device)
        .map_err(|e| ferr("CUDA GPTQ", e))?;
    WIDEST_K.fetch_max(k, Ordering::Relaxed);
    Ok(Box::new(Fb2GptqLinear::new(lin)))
}

/// Body for `<CudaBackend as BackendQuantMarlin>::load_gptq_stacked` (capabilities.rs:180).
#[allow(clippy::too_many_arguments)]
pub fn load_gptq_stacked(qweights: &[&[i32]], scales: &[&[f32]], qzeros: &[&[i32]], g_idx: Option<&[i32]>, bits: u32,
                         group_size: usize, k: usize, n_per_expert: usize)
                         -> Result<Arc<dyn MarlinExpertStack<CudaBackend>>> {
    let device = crate::backend::cuda::current_device_ordinal() as i32;
    let st = fb2::ExpertStack::load_gptq_stacked(qweights, scales, qzeros, g_idx, bits, group_size, k, n_per_expert, device)
        .map_err(|e| ferr("CUDA GPTQ stacked", e))?;
    WIDEST_K.fetch_max(k, Ordering::Relaxed);
    Ok(Arc::new(Fb2ExpertStack { inner: Arc::new(st) }))
}

/// Body for `pregrow_marlin_gather_scratch` (capabilities.rs:212): the act-order gather scratch must
/// exist before a CUDA-graph capture begins.  The reference passes `required` in ELEMENTS,
/// `m_total * intermediate_size` (llama_family_forward_batched.rs:1860-1908); the library sizes its
/// scratch (permuted activations and split-K partials) by ROWS for every weight loaded on the device,
/// so the request is divided by the widest K loaded so far (= `intermediate_size` once `down_proj` exists).
pub fn pregrow_marlin_gather_scratch(ctx: &mut CudaState, required: usize) {
    let widest = WIDEST_K.load(Ordering::Relaxed);
    if widest == 0 {
        return; // no GPTQ weight loaded: nothing to size
    }
    let max_m = (required + widest - 1) / widest;
    if let Err(e) = fb2::pregrow_scratch(ctx.ordinal as i32, max_m) {
        panic!("fb2 pregrow scratch ({required} elements = {max_m} rows): {e}");
    }
}

/// Whole-FFN call for `moe_forward_bucketed` (ferrum-models/src/moe/dispatch.rs:1558): route, align,
/// gate_up + SiLU, down, weighted combine in one library call on `ctx.stream`.  `workspace` is an i8
/// `CudaBuf` of at least `fb2::moe_ffn_workspace_bytes(..)` bytes, allocated once at model build.
#[allow(clippy::too_many_arguments)]
pub fn moe_ffn_forward(ctx: &mut CudaState, gate_up: &dyn MarlinExpertStack<CudaBackend>,
                       down: &dyn MarlinExpertStack<CudaBackend>, x: &CudaBuf, router_logits: &CudaBuf, out: &mut CudaBuf,
                       workspace: &mut CudaBuf, batch: usize, top_k: usize, norm_topk_prob: bool) -> Result<()> {
    let (gu, dn) = match (gate_up.as_any().downcast_ref::<Fb2ExpertStack>(), down.as_any().downcast_ref::<Fb2ExpertStack>()) {
        (Some(g), Some(d)) => (g, d),
        _ => return Err(FerrumError::unsupported("moe_ffn_forward: expert stacks are not fb2 stacks")),
    };
    let stream = ctx.stream.clone();
    let ws_bytes = workspace.len();
    let (xp, _g0) = x.as_f16().device_ptr(&stream);
    let (lp, _g1) = router_logits.as_f16().device_ptr(&stream);
    let (op, _g2) = out.as_f16_mut().device_ptr_mut(&stream);
    let (wp, _g3) = match workspace {
        CudaBuf::I8(s) => s.device_ptr_mut(&stream),
        _ => return Err(FerrumError::model("moe_ffn_forward: workspace must be an I8 buffer")),
    };
    unsafe {
        fb2::moe_ffn_forward(&gu.inner, &dn.inner, raw_stream(ctx), xp as fb2::DevF16, lp as fb2::DevF16, op as fb2::DevF16Mut,
                             batch, top_k, norm_topk_prob, wp as *mut c_void, ws_bytes)
    }
    .map_err(|e| ferr("fb2 moe_ffn_forward", e))
}

// ---------------------------------------------------------------- BackendMoeFused entry points
// Bodies for `impl BackendMoeFused for CudaBackend` (backend/cuda/moe.rs:145-560).  Buffer dtypes follow
// that impl: logits / activations F16, ids and index arrays I32, routing weights F32.

/// `route_topk_softmax` (capabilities.rs:321, CUDA impl moe.rs:146).
#[allow(clippy::too_many_arguments)]
pub fn route_topk_softmax(ctx: &mut CudaState, logits: &CudaBuf, out_ids: &mut CudaBuf, out_weights: &mut CudaBuf,
                          batch: usize, num_experts: usize, top_k: usize, norm_topk_prob: bool) -> Result<()> {
    let stream = ctx.stream.clone();
    let (l, _g0) = logits.as_f16().device_ptr(&stream);
    let (i, _g1) = out_ids.as_i32_mut().device_ptr_mut(&stream);
    let (w, _g2) = out_weights.as_f32_mut().device_ptr_mut(&stream);
    unsafe {
        fb2::moe::route_topk_softmax(raw_stream(ctx), l as fb2::DevF16, i as *mut i32, w as *mut f32, batch, num_experts,
                                     top_k, norm_topk_prob)
    }
    .map_err(|e| ferr("fb2 route_topk_softmax", e))
}

/// `moe_build_pairs_by_token` (capabilities.rs:372, CUDA impl moe.rs:314).
#[allow(clippy::too_many_arguments)]
pub fn moe_build_pairs_by_token(ctx: &mut CudaState, expert_ids: &CudaBuf, pairs_by_token: &mut CudaBuf,
                                packed_token_idx: &mut CudaBuf, expert_offsets: &mut CudaBuf, batch_x_topk: usize,
                                num_experts: usize, top_k: usize) -> Result<()> {
    let stream = ctx.stream.clone();
    let (e, _g0) = expert_ids.as_i32().device_ptr(&stream);
    let (p, _g1) = pairs_by_token.as_i32_mut().device_ptr_mut(&stream);
    let (t, _g2) = packed_token_idx.as_i32_mut().device_ptr_mut(&stream);
    let (o, _g3) = expert_offsets.as_i32_mut().device_ptr_mut(&stream);
    unsafe {
        fb2::moe::build_pairs_by_token(raw_stream(ctx), e as *const i32, p as *mut i32, t as *mut i32, o as *mut i32,
                                       batch_x_topk, num_experts, top_k)
    }
    .map_err(|e| ferr("fb2 moe_build_pairs_by_token", e))
}

#[allow(clippy::too_many_arguments)]
fn align_impl(ctx: &mut CudaState, expert_ids_per_pair: &CudaBuf, sorted_token_ids: &mut CudaBuf, block_ids: &mut CudaBuf,
              total_tokens_post_pad: &mut CudaBuf, batch_x_topk: usize, num_experts: usize, block_size: usize,
              sorted_max_size: usize, pair_ids: bool) -> Result<()> {
    let stream = ctx.stream.clone();
    let (e, _g0) = expert_ids_per_pair.as_i32().device_ptr(&stream);
    let (s, _g1) = sorted_token_ids.as_i32_mut().device_ptr_mut(&stream);
    let (b, _g2) = block_ids.as_i32_mut().device_ptr_mut(&stream);
    let (t, _g3) = total_tokens_post_pad.as_i32_mut().device_ptr_mut(&stream);
    unsafe {
        fb2::moe::align_block_size(raw_stream(ctx), e as *const i32, s as *mut i32, b as *mut i32, t as *mut i32,
                                   batch_x_topk, num_experts, block_size, sorted_max_size, pair_ids)
    }
    .map_err(|e| ferr("fb2 moe_align_block_size", e))
}

/// `moe_align_block_size` (capabilities.rs:389, CUDA impl moe.rs:358): sorted ids are token indices.
#[allow(clippy::too_many_arguments)]
pub fn moe_align_block_size(ctx: &mut CudaState, expert_ids_per_pair: &CudaBuf, sorted_token_ids: &mut CudaBuf,
                            block_ids: &mut CudaBuf, total_tokens_post_pad: &mut CudaBuf, batch_x_topk: usize,
                            num_experts: usize, block_size: usize, sorted_max_size: usize) -> Result<()> {
    align_impl(ctx, expert_ids_per_pair, sorted_token_ids, block_ids, total_tokens_post_pad, batch_x_topk, num_experts,
               block_size, sorted_max_size, false)
}

/// `moe_align_block_size_pair_ids` (capabilities.rs:406, CUDA impl moe.rs:420): sorted ids are pair indices.
#[allow(clippy::too_many_arguments)]
pub fn moe_align_block_size_pair_ids(ctx: &mut CudaState, expert_ids_per_pair: &CudaBuf, sorted_token_ids: &mut CudaBuf,
                                     block_ids: &mut CudaBuf, total_tokens_post_pad: &mut CudaBuf, batch_x_topk: usize,
                                     num_experts: usize, block_size: usize, sorted_max_size: usize) -> Result<()> {
    align_impl(ctx, expert_ids_per_pair, sorted_token_ids, block_ids, total_tokens_post_pad, batch_x_topk, num_experts,
               block_size, sorted_max_size, true)
}

/// `moe_combine` (capabilities.rs:683, CUDA impl moe.rs:476): no `Result` in the trait, a failed launch panics there too.
#[allow(clippy::too_many_arguments)]
pub fn moe_combine(ctx: &mut CudaState, packed_down: &CudaBuf, pairs_by_token: &CudaBuf, pair_weights: &CudaBuf,
                   out: &mut CudaBuf, batch: usize, hidden: usize, top_k: usize, _total_pairs: usize) {
    let stream = ctx.stream.clone();
    let (d, _g0) = packed_down.as_f16().device_ptr(&stream);
    let (p, _g1) = pairs_by_token.as_i32().device_ptr(&stream);
    let (w, _g2) = pair_weights.as_f32().device_ptr(&stream);
    let (o, _g3) = out.as_f16_mut().device_ptr_mut(&stream);
    let rc = unsafe {
        fb2::moe::combine(raw_stream(ctx), d as fb2::DevF16, p as *const i32, w as *const f32, o as fb2::DevF16Mut, batch,
                          hidden, top_k)
    };
    if let Err(e) = rc {
        panic!("fb2 moe_combine: {e}");
    }
}

/// `weighted_sum_batched` (capabilities.rs:571, CUDA impl moe.rs:519).
pub fn weighted_sum_batched(ctx: &mut CudaState, slots: &CudaBuf, weights: &CudaBuf, out: &mut CudaBuf, batch: usize,
                            top_k: usize, hidden: usize) -> Result<()> {
    let stream = ctx.stream.clone();
    let (s, _g0) = slots.as_f16().device_ptr(&stream);
    let (w, _g1) = weights.as_f32().device_ptr(&stream);
    let (o, _g2) = out.as_f16_mut().device_ptr_mut(&stream);
    unsafe {
        fb2::moe::weighted_sum_batched(raw_stream(ctx), s as fb2::DevF16, w as *const f32, o as fb2::DevF16Mut, batch, top_k,
                                       hidden)
    }
    .map_err(|e| ferr("fb2 weighted_sum_batched", e))
}

/// `fused_silu_mul_split` (Backend trait; CPU statement backend/cpu.rs:1666).
pub fn fused_silu_mul_split(ctx: &mut CudaState, gate_up: &CudaBuf, out: &mut CudaBuf, tokens: usize, im: usize) {
    let stream = ctx.stream.clone();
    let (g, _g0) = gate_up.as_f16().device_ptr(&stream);
    let (o, _g1) = out.as_f16_mut().device_ptr_mut(&stream);
    let rc = unsafe { fb2::moe::fused_silu_mul_split(raw_stream(ctx), g as fb2::DevF16, o as fb2::DevF16Mut, tokens, im) };
    if let Err(e) = rc {
        panic!("fb2 fused_silu_mul_split: {e}");
    }
}

// ------------------------------------------------------------------------------- collectives
/// The shape of `NcclRank` (backend/cuda/nccl_comm.rs:8-56) over the library's communicator, one per
/// rank process.  `tp_decode.rs:351-375` calls `all_reduce_f16_inplace` after each row-parallel GEMM;
/// `forward_allreduce` below replaces that pair with the fused peer-memory kernel.
pub struct Fb2NcclRank {
    comm: fb2::Comm,
    stream: Arc<cudarc::driver::CudaStream>,
}
unsafe impl Send for Fb2NcclRank {}
unsafe impl Sync for Fb2NcclRank {}

impl Fb2NcclRank {
    /// Rank 0 creates the id (`fb2::Comm::unique_id`) and hands it to the other rank processes.
    pub fn init_rank(id: &[u8; fb2::ffi::FB2_UNIQUE_ID_BYTES], rank: usize, world_size: usize, ordinal: usize,
                     stream: Arc<cudarc::driver::CudaStream>) -> Result<Self> {
        let comm = fb2::Comm::init_rank(id, rank as i32, world_size as i32, ordinal as i32).map_err(|e| ferr("fb2 comm init", e))?;
        Ok(Self { comm, stream })
    }
    pub fn rank(&self) -> usize {
        self.comm.rank()
    }
    pub fn world_size(&self) -> usize {
        self.comm.world_size()
    }
    /// `NcclRank::all_reduce_f16_inplace` (nccl_comm.rs:40): sum over ranks, in place, on this rank's stream.
    pub fn all_reduce_f16_inplace(&self, buf: &mut cudarc::driver::CudaSlice<half::f16>) -> Result<()> {
        let len = buf.len();
        let (p, _g) = buf.device_ptr_mut(&self.stream);
        unsafe { self.comm.all_reduce_f16(self.stream.cu_stream() as *mut c_void, p as fb2::DevF16Mut, len) }
            .map_err(|e| ferr("fb2 all_reduce", e))
    }
    /// Peer-memory receive slots for `forward_allreduce`; collective, outside graph capture.
    pub fn enable_p2p(&mut self, max_elems: usize) -> Result<()> {
        self.comm.enable_p2p(self.stream.cu_stream() as *mut c_void, max_elems).map_err(|e| ferr("fb2 enable_p2p", e))
    }
    /// Row-parallel GEMM + all-reduce in one call (the epilogue stores tiles into every peer's slot over NVLink).
    pub fn forward_allreduce(&self, lin: &Fb2GptqLinear, input: &CudaBuf, out: &mut CudaBuf, m: usize) -> Result<()> {
        let (x, _g0) = input.as_f16().device_ptr(&self.stream);
        let (y, _g1) = out.as_f16_mut().device_ptr_mut(&self.stream);
        unsafe {
            lin.handle().forward_allreduce(&self.comm, self.stream.cu_stream() as *mut c_void, x as fb2::DevF16,
                                           y as fb2::DevF16Mut, m)
        }
        .map_err(|e| ferr("fb2 forward_allreduce", e))
    }
}
