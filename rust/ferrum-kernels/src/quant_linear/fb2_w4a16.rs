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
