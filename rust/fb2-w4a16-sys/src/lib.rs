//! Safe, owning handles over the C ABI of `libfb2_w4a16.so`.
//!
//! `ffi` is generated from `include/fb2_w4a16.h`.  The types here add ownership (`Drop`), the
//! `Send + Sync` promise the header makes ("handles are immutable after create"), and the status ->
//! error mapping the reference uses (`FerrumError::{unsupported, model, io}`).  Nothing in this crate
//! touches CUDA directly: device pointers and streams are passed through as the caller owns them
//! (`CudaSlice` device pointers / `CUstream` in the reference, `backend/cuda/mod.rs:234-237,612`).
//!
//! This crate has not been compiled in the repository it ships in (no Rust toolchain in that image);
//! the same ABI is exercised from Python by the test-suite.
//!
//! The impls of ferrum's own traits (`Linear<CudaBackend>`, `MarlinExpertStack<CudaBackend>`,
//! `BackendQuantMarlin`) over these handles live in `rust/ferrum-kernels/src/quant_linear/fb2_w4a16.rs`:
//! they name `crate::backend::...` types and so belong inside the reference's `ferrum-kernels` crate.
pub mod ffi;

use std::ffi::{CStr, CString};
use std::os::raw::{c_int, c_void};
use std::ptr;

/// A device pointer to IEEE fp16 values (the ABI types them as `u16` bit patterns).
pub type DevF16 = *const u16;
pub type DevF16Mut = *mut u16;
/// `CUstream` / `cudaStream_t`; null = the default stream.
pub type Stream = *mut c_void;

#[derive(Debug, Clone, PartialEq, Eq)]
pub enum Fb2Error {
    Shape(String),
    Unsupported(String),
    Cuda(String),
    Nccl(String),
    Arg(String),
    Model(String),
    Io(String),
}

impl std::fmt::Display for Fb2Error {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        let (kind, msg) = match self {
            Fb2Error::Shape(m) => ("shape", m),
            Fb2Error::Unsupported(m) => ("unsupported", m),
            Fb2Error::Cuda(m) => ("cuda", m),
            Fb2Error::Nccl(m) => ("nccl", m),
            Fb2Error::Arg(m) => ("argument", m),
            Fb2Error::Model(m) => ("model", m),
            Fb2Error::Io(m) => ("io", m),
        };
        write!(f, "fb2 {kind} error: {msg}")
    }
}
impl std::error::Error for Fb2Error {}
pub type Result<T> = std::result::Result<T, Fb2Error>;

/// Status code -> error, with the thread-local message of the failing call.
pub fn check(rc: c_int) -> Result<()> {
    if rc == ffi::FB2_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::fb2_last_error()) }.to_string_lossy().into_owned();
    Err(match rc {
        ffi::FB2_ERR_SHAPE => Fb2Error::Shape(msg),
        ffi::FB2_ERR_UNSUPPORTED => Fb2Error::Unsupported(msg),
        ffi::FB2_ERR_CUDA => Fb2Error::Cuda(msg),
        ffi::FB2_ERR_NCCL => Fb2Error::Nccl(msg),
        ffi::FB2_ERR_ARG => Fb2Error::Arg(msg),
        ffi::FB2_ERR_IO => Fb2Error::Io(msg),
        _ => Fb2Error::Model(msg),
    })
}

fn opt_ptr<T>(s: Option<&[T]>) -> *const T {
    s.map_or(ptr::null(), |v| v.as_ptr())
}

// ------------------------------------------------------------------------------------ linear
/// One GPTQ linear on one device (`Box<dyn Linear<B>>` in the reference).
pub struct GptqLinear {
    h: *mut ffi::Fb2Linear,
    /// keeps a stack alive while this linear is a view of it (`StackedExpertLinear`, gptq.rs:154-198)
    _stack: Option<std::sync::Arc<ExpertStack>>,
}
unsafe impl Send for GptqLinear {}
unsafe impl Sync for GptqLinear {}

impl GptqLinear {
    /// `BackendQuantMarlin::load_gptq` (capabilities.rs:136), argument for argument.
    #[allow(clippy::too_many_arguments)]
    pub fn load_gptq(qweight: &[i32], scales: &[f32], qzeros: &[i32], g_idx: Option<&[i32]>, bias: Option<&[f32]>,
                     bits: u32, group_size: usize, k: usize, n: usize, device: i32) -> Result<Self> {
        if group_size == 0 || k % 8 != 0 || n % 8 != 0 || k % group_size != 0 {
            return Err(Fb2Error::Shape(format!("GPTQ linear K={k} N={n} group_size={group_size}")));
        }
        let g = k / group_size;
        if qweight.len() != k / 8 * n || scales.len() != g * n || qzeros.len() != g * (n / 8)
            || g_idx.map_or(false, |v| v.len() != k) || bias.map_or(false, |v| v.len() != n) {
            return Err(Fb2Error::Model(format!("GPTQ tensor lengths do not match K={k} N={n} group_size={group_size}")));
        }
        let mut h = ptr::null_mut();
        check(unsafe {
            ffi::fb2_gptq_linear_create(qweight.as_ptr(), scales.as_ptr(), qzeros.as_ptr(), opt_ptr(g_idx), opt_ptr(bias),
                                        bits, group_size, k, n, device, &mut h)
        })?;
        Ok(Self { h, _stack: None })
    }

    /// This rank's tensor-parallel shard (`fb2_gptq_linear_create_sharded`); `x_cols` receives the
    /// activation columns a row-parallel shard consumes.
    #[allow(clippy::too_many_arguments)]
    pub fn load_gptq_sharded(qweight: &[i32], scales: &[f32], qzeros: &[i32], g_idx: Option<&[i32]>, bias: Option<&[f32]>,
                             bits: u32, group_size: usize, k: usize, n: usize, device: i32, tp_rank: i32, tp_size: i32,
                             split: Split, segments: &[usize], x_cols: Option<&mut [i32]>) -> Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe {
            ffi::fb2_gptq_linear_create_sharded(qweight.as_ptr(), scales.as_ptr(), qzeros.as_ptr(), opt_ptr(g_idx),
                                                opt_ptr(bias), bits, group_size, k, n, device, tp_rank, tp_size, split as c_int,
                                                if segments.is_empty() { ptr::null() } else { segments.as_ptr() },
                                                segments.len() as c_int, x_cols.map_or(ptr::null_mut(), |v| v.as_mut_ptr()),
                                                &mut h)
        })?;
        Ok(Self { h, _stack: None })
    }

    pub fn in_features(&self) -> usize { unsafe { ffi::fb2_gptq_linear_in_features(self.h) } }
    pub fn out_features(&self) -> usize { unsafe { ffi::fb2_gptq_linear_out_features(self.h) } }
    pub fn is_act_order(&self) -> bool { unsafe { ffi::fb2_gptq_linear_is_act_order(self.h) != 0 } }
    pub fn as_ptr(&self) -> *const ffi::Fb2Linear { self.h }

    /// `Linear::forward` (linear.rs:123): appends the GEMM to `stream`; no sync, no allocation.
    ///
    /// # Safety
    /// `x` / `y` must be device pointers to `m * K` / `m * N` halves on this linear's device.
    pub unsafe fn forward(&self, stream: Stream, x: DevF16, y: DevF16Mut, m: usize) -> Result<()> {
        check(ffi::fb2_gptq_linear_forward(self.h, stream, x, y, m))
    }

    /// Row-parallel forward + all-reduce (`tp_decode.rs:351-375` does GEMM then `all_reduce_f16_inplace`).
    ///
    /// # Safety
    /// as `forward`; every rank of `comm` must make the same call.
    pub unsafe fn forward_allreduce(&self, comm: &Comm, stream: Stream, x: DevF16, y: DevF16Mut, m: usize) -> Result<()> {
        check(ffi::fb2_gptq_linear_forward_allreduce(self.h, comm.h, stream, x, y, m))
    }

    /// Host-buffer form of the call (synchronous).
    pub fn forward_host(&mut self, x: &[u16], y: &mut [u16], m: usize) -> Result<()> {
        if x.len() < m * self.in_features() || y.len() < m * self.out_features() {
            return Err(Fb2Error::Shape("forward_host: buffer too small".into()));
        }
        check(unsafe { ffi::fb2_gptq_linear_forward_host(self.h, x.as_ptr(), y.as_mut_ptr(), m) })
    }
}
impl Drop for GptqLinear {
    fn drop(&mut self) { unsafe { ffi::fb2_gptq_linear_destroy(self.h) } }
}

#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum Split { Column = 0, Row = 1 }

/// `BackendQuantMarlin::pregrow_marlin_gather_scratch` (capabilities.rs:212).
pub fn pregrow_scratch(device: i32, max_m: usize) -> Result<()> {
    check(unsafe { ffi::fb2_pregrow_scratch(device, max_m) })
}

// ------------------------------------------------------------------------------------- stack
/// `Arc<dyn MarlinExpertStack<B>>` (marlin_expert_stack.rs:31-109).
pub struct ExpertStack { h: *mut ffi::Fb2Stack }
unsafe impl Send for ExpertStack {}
unsafe impl Sync for ExpertStack {}

impl ExpertStack {
    /// `BackendQuantMarlin::load_gptq_stacked` (capabilities.rs:180).
    #[allow(clippy::too_many_arguments)]
    pub fn load_gptq_stacked(qweights: &[&[i32]], scales: &[&[f32]], qzeros: &[&[i32]], g_idx: Option<&[i32]>, bits: u32,
                             group_size: usize, k: usize, n_per_expert: usize, device: i32) -> Result<Self> {
        if qweights.is_empty() || scales.len() != qweights.len() || qzeros.len() != qweights.len() {
            return Err(Fb2Error::Model("load_gptq_stacked: empty or ragged expert lists".into()));
        }
        let qw: Vec<*const i32> = qweights.iter().map(|s| s.as_ptr()).collect();
        let sc: Vec<*const f32> = scales.iter().map(|s| s.as_ptr()).collect();
        let qz: Vec<*const i32> = qzeros.iter().map(|s| s.as_ptr()).collect();
        let mut h = ptr::null_mut();
        check(unsafe {
            ffi::fb2_gptq_stack_create(qw.as_ptr(), sc.as_ptr(), qz.as_ptr(), opt_ptr(g_idx), bits, group_size, k,
                                       n_per_expert, qweights.len(), device, &mut h)
        })?;
        Ok(Self { h })
    }

    pub fn n_per_expert(&self) -> usize { unsafe { ffi::fb2_gptq_stack_n_per_expert(self.h) } }
    pub fn k(&self) -> usize { unsafe { ffi::fb2_gptq_stack_k(self.h) } }
    pub fn num_experts(&self) -> usize { unsafe { ffi::fb2_gptq_stack_num_experts(self.h) } }
    pub fn as_ptr(&self) -> *const ffi::Fb2Stack { self.h }

    /// `zero_workspace` (marlin_expert_stack.rs:57): nothing to clear in this layout, kept for the trait.
    pub fn zero_workspace(&self, stream: Stream) -> Result<()> {
        check(unsafe { ffi::fb2_gptq_stack_zero_workspace(self.h, stream) })
    }

    /// `make_expert_linear` (marlin_expert_stack.rs:103): the view keeps the stack alive.
    pub fn make_expert_linear(self: &std::sync::Arc<Self>, expert_offset: usize, expert_n: usize, bias: Option<&[f32]>)
                              -> Result<GptqLinear> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::fb2_gptq_stack_make_expert_linear(self.h, expert_offset, expert_n, opt_ptr(bias), &mut h) })?;
        Ok(GptqLinear { h, _stack: Some(self.clone()) })
    }

    /// `gemm_phase_batched` (marlin_expert_stack.rs:64-70).
    ///
    /// # Safety
    /// device pointers sized for the dispatch list.
    pub unsafe fn gemm_phase_batched(&self, stream: Stream, x: DevF16, dispatches: &[(usize, usize, usize, usize)],
                                     y: DevF16Mut, k: usize) -> Result<()> {
        let d: Vec<ffi::Fb2Dispatch> = dispatches.iter()
            .map(|&(e, i, o, m)| ffi::Fb2Dispatch { expert: e as i64, in_row: i as i64, out_row: o as i64, m: m as i64 })
            .collect();
        check(ffi::fb2_gptq_stack_gemm_batched(self.h, stream, x, d.as_ptr(), d.len(), y, k))
    }

    /// `gemm_phase_vllm` (marlin_expert_stack.rs:82-96); the output need not be pre-zeroed.
    ///
    /// # Safety
    /// device pointers; `sorted_capacity` = length of `sorted_token_ids`.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn gemm_phase_vllm(&self, stream: Stream, x: DevF16, sorted_token_ids: *const i32, expert_ids: *const i32,
                                  num_tokens_past_padded: *const i32, y: DevF16Mut, prob_m: usize, moe_block_size: usize,
                                  top_k: usize, sorted_capacity: usize) -> Result<()> {
        check(ffi::fb2_gptq_stack_gemm_routed(self.h, stream, x, sorted_token_ids, expert_ids, num_tokens_past_padded, y,
                                              prob_m, moe_block_size, top_k, sorted_capacity))
    }

    /// gate_up GEMM with the SiLU-and-multiply epilogue (`fused_silu_mul_split`, backend/cpu.rs:1666) fused:
    /// `act` receives `[pairs, n_per_expert / 2]`.
    ///
    /// # Safety
    /// as `gemm_phase_vllm`.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn gemm_phase_vllm_silu(&self, stream: Stream, x: DevF16, sorted_token_ids: *const i32, expert_ids: *const i32,
                                       num_tokens_past_padded: *const i32, act: DevF16Mut, prob_m: usize,
                                       moe_block_size: usize, top_k: usize, sorted_capacity: usize) -> Result<()> {
        check(ffi::fb2_gptq_stack_gemm_routed_silu(self.h, stream, x, sorted_token_ids, expert_ids, num_tokens_past_padded,
                                                   act, prob_m, moe_block_size, top_k, sorted_capacity))
    }
}
impl Drop for ExpertStack {
    fn drop(&mut self) { unsafe { ffi::fb2_gptq_stack_destroy(self.h) } }
}

/// The whole MoE FFN of one layer (`moe_forward_bucketed`, dispatch.rs:1558-2192).
///
/// # Safety
/// device pointers; `workspace` holds `moe_ffn_workspace_bytes` bytes.
#[allow(clippy::too_many_arguments)]
pub unsafe fn moe_ffn_forward(gate_up: &ExpertStack, down: &ExpertStack, stream: Stream, x: DevF16, router_logits: DevF16,
                              out: DevF16Mut, batch: usize, top_k: usize, norm_topk_prob: bool, workspace: *mut c_void,
                              workspace_bytes: usize) -> Result<()> {
    check(ffi::fb2_moe_ffn_forward(gate_up.h, down.h, stream, x, router_logits, out, batch, top_k, norm_topk_prob as c_int,
                                   workspace, workspace_bytes))
}
pub fn moe_ffn_workspace_bytes(gate_up: &ExpertStack, down: &ExpertStack, max_batch: usize, top_k: usize) -> usize {
    unsafe { ffi::fb2_moe_ffn_workspace_bytes(gate_up.h, down.h, max_batch, top_k) }
}

/// Expert-parallel form: this rank holds `num_experts / world_size` experts in `gate_up` / `down`; tokens travel
/// by one grouped all-to-all each way (`fb2_moe_ep_ffn_forward`).
///
/// # Safety
/// device pointers; every rank of `comm` makes the same call; `workspace` holds `moe_ep_workspace_bytes` bytes.
#[allow(clippy::too_many_arguments)]
pub unsafe fn moe_ep_ffn_forward(gate_up: &ExpertStack, down: &ExpertStack, comm: &Comm, stream: Stream, x: DevF16,
                                 router_logits: DevF16, out: DevF16Mut, batch_local: usize, top_k: usize,
                                 norm_topk_prob: bool, workspace: *mut c_void, workspace_bytes: usize) -> Result<()> {
    check(ffi::fb2_moe_ep_ffn_forward(gate_up.h, down.h, comm.h, stream, x, router_logits, out, batch_local, top_k,
                                      norm_topk_prob as c_int, workspace, workspace_bytes))
}
pub fn moe_ep_workspace_bytes(gate_up: &ExpertStack, down: &ExpertStack, max_batch_local: usize, top_k: usize,
                              world_size: usize) -> usize {
    unsafe { ffi::fb2_moe_ep_workspace_bytes(gate_up.h, down.h, max_batch_local, top_k, world_size as c_int) }
}

// --------------------------------------------------------------------------------------- moe
/// The routing pre/post-ops of `BackendMoeFused` (capabilities.rs:305-723).  Every function appends
/// work to `stream` and returns; all pointers are device pointers owned by the caller.
pub mod moe {
    use super::{check, ffi, DevF16, DevF16Mut, Result, Stream};
    use std::os::raw::c_int;

    /// `route_topk_softmax` (capabilities.rs:321; CPU `moe/router.rs:183-257`): ids `[batch, top_k]` i32,
    /// weights `[batch, top_k]` f32; ties go to the lower expert index.
    ///
    /// # Safety
    /// device pointers of the stated shapes.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn route_topk_softmax(stream: Stream, logits: DevF16, out_ids: *mut i32, out_weights: *mut f32, batch: usize,
                                     num_experts: usize, top_k: usize, norm_topk_prob: bool) -> Result<()> {
        check(ffi::fb2_moe_route_topk_softmax(stream, logits, out_ids, out_weights, batch, num_experts, top_k,
                                              norm_topk_prob as c_int))
    }

    /// `moe_build_pairs_by_token` (capabilities.rs:372): deterministic (token-ascending inside an expert).
    ///
    /// # Safety
    /// device pointers: `expert_ids`, `pairs_by_token`, `packed_token_idx` hold `batch_x_topk` i32;
    /// `expert_offsets` holds `num_experts + 1`.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn build_pairs_by_token(stream: Stream, expert_ids: *const i32, pairs_by_token: *mut i32,
                                       packed_token_idx: *mut i32, expert_offsets: *mut i32, batch_x_topk: usize,
                                       num_experts: usize, top_k: usize) -> Result<()> {
        check(ffi::fb2_moe_build_pairs_by_token(stream, expert_ids, pairs_by_token, packed_token_idx, expert_offsets,
                                                batch_x_topk, num_experts, top_k))
    }

    /// `moe_align_block_size` / `moe_align_block_size_pair_ids` (capabilities.rs:389,406): `pair_ids`
    /// selects the second convention (sorted ids are pair indices, not token indices).
    ///
    /// # Safety
    /// device pointers; `sorted_token_ids` holds `sorted_max_size` i32, `block_ids` holds
    /// `sorted_max_size / block_size`, `total_tokens_post_pad` one i32.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn align_block_size(stream: Stream, expert_ids_per_pair: *const i32, sorted_token_ids: *mut i32,
                                   block_ids: *mut i32, total_tokens_post_pad: *mut i32, batch_x_topk: usize,
                                   num_experts: usize, block_size: usize, sorted_max_size: usize, pair_ids: bool)
                                   -> Result<()> {
        check(ffi::fb2_moe_align_block_size(stream, expert_ids_per_pair, sorted_token_ids, block_ids, total_tokens_post_pad,
                                            batch_x_topk, num_experts, block_size, sorted_max_size, pair_ids as c_int))
    }

    /// `moe_combine` (capabilities.rs:683): `out[b] = sum_k w[b,k] * packed_down[pairs_by_token[b,k]]`, fp32 sum.
    ///
    /// # Safety
    /// device pointers of the stated shapes.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn combine(stream: Stream, packed_down: DevF16, pairs_by_token: *const i32, pair_weights: *const f32,
                          out: DevF16Mut, batch: usize, hidden: usize, top_k: usize) -> Result<()> {
        check(ffi::fb2_moe_combine(stream, packed_down, pairs_by_token, pair_weights, out, batch, hidden, top_k))
    }

    /// `weighted_sum_batched` (capabilities.rs:571): slots `[batch, top_k, hidden]`.
    ///
    /// # Safety
    /// device pointers of the stated shapes.
    pub unsafe fn weighted_sum_batched(stream: Stream, slots: DevF16, weights: *const f32, out: DevF16Mut, batch: usize,
                                       top_k: usize, hidden: usize) -> Result<()> {
        check(ffi::fb2_moe_weighted_sum_batched(stream, slots, weights, out, batch, top_k, hidden))
    }

    /// `fused_silu_mul_split` (backend/cpu.rs:1666): `[tokens, 2*im]` -> `[tokens, im]`.
    ///
    /// # Safety
    /// device pointers of the stated shapes.
    pub unsafe fn fused_silu_mul_split(stream: Stream, gate_up: DevF16, out: DevF16Mut, tokens: usize, im: usize) -> Result<()> {
        check(ffi::fb2_fused_silu_mul_split(stream, gate_up, out, tokens, im))
    }

    /// The Gemma form of the same op (`fused_gelu_tanh_mul_split`).
    ///
    /// # Safety
    /// device pointers of the stated shapes.
    pub unsafe fn fused_gelu_tanh_mul_split(stream: Stream, gate_up: DevF16, out: DevF16Mut, tokens: usize, im: usize)
                                            -> Result<()> {
        check(ffi::fb2_fused_gelu_tanh_mul_split(stream, gate_up, out, tokens, im))
    }
}

// -------------------------------------------------------------------------------------- comm
/// One NCCL communicator per process (`NcclRank`, nccl_comm.rs:12-105; `BackendCollective`, capabilities.rs:85-107).
pub struct Comm { h: *mut ffi::Fb2Comm }
unsafe impl Send for Comm {}

impl Comm {
    pub fn unique_id() -> Result<[u8; ffi::FB2_UNIQUE_ID_BYTES]> {
        let mut id = [0u8; ffi::FB2_UNIQUE_ID_BYTES];
        check(unsafe { ffi::fb2_comm_get_unique_id(id.as_mut_ptr() as *mut c_void) })?;
        Ok(id)
    }
    pub fn init_rank(id: &[u8; ffi::FB2_UNIQUE_ID_BYTES], rank: i32, world_size: i32, device: i32) -> Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { ffi::fb2_comm_init_rank(id.as_ptr() as *const c_void, rank, world_size, device, &mut h) })?;
        Ok(Self { h })
    }
    pub fn world_size(&self) -> usize { unsafe { ffi::fb2_comm_world_size(self.h) as usize } }
    pub fn rank(&self) -> usize { unsafe { ffi::fb2_comm_rank(self.h) as usize } }
    /// Peer-memory receive buffers for the fused GEMM + all-reduce (collective over all ranks).
    pub fn enable_p2p(&mut self, stream: Stream, max_elems: usize) -> Result<()> {
        check(unsafe { ffi::fb2_comm_enable_p2p(self.h, stream, max_elems) })
    }
    /// `all_reduce_f16_inplace` (nccl_comm.rs:40).
    ///
    /// # Safety
    /// `buf` is a device pointer to `len` halves; every rank makes the same call.
    pub unsafe fn all_reduce_f16(&self, stream: Stream, buf: DevF16Mut, len: usize) -> Result<()> {
        check(ffi::fb2_all_reduce_f16(self.h, stream, buf, len))
    }
    /// `BackendCollective::all_gather` (capabilities.rs:96): `recv` holds `world_size * len_per_rank` halves.
    ///
    /// # Safety
    /// device pointers; every rank makes the same call.
    pub unsafe fn all_gather_f16(&self, stream: Stream, send: DevF16, recv: DevF16Mut, len_per_rank: usize) -> Result<()> {
        check(ffi::fb2_all_gather_f16(self.h, stream, send, recv, len_per_rank))
    }
    /// `BackendCollective::broadcast` (capabilities.rs:105).
    ///
    /// # Safety
    /// device pointer; every rank makes the same call.
    pub unsafe fn broadcast_f16(&self, stream: Stream, buf: DevF16Mut, len: usize, root: usize) -> Result<()> {
        check(ffi::fb2_broadcast_f16(self.h, stream, buf, len, root as c_int))
    }
    /// Variable-count all-to-all (expert-parallel dispatch / combine): element offsets and counts per peer.
    ///
    /// # Safety
    /// device pointers; the four slices have `world_size` entries; every rank makes the matching call.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn all_to_all_f16(&self, stream: Stream, send: DevF16, send_offsets: &[usize], send_counts: &[usize],
                                 recv: DevF16Mut, recv_offsets: &[usize], recv_counts: &[usize]) -> Result<()> {
        let w = self.world_size();
        if send_offsets.len() != w || send_counts.len() != w || recv_offsets.len() != w || recv_counts.len() != w {
            return Err(Fb2Error::Arg(format!("all_to_all_f16: offset / count slices must have {w} entries")));
        }
        check(ffi::fb2_all_to_all_f16(self.h, stream, send, send_offsets.as_ptr(), send_counts.as_ptr(), recv,
                                      recv_offsets.as_ptr(), recv_counts.as_ptr()))
    }
}
impl Drop for Comm {
    fn drop(&mut self) { unsafe { ffi::fb2_comm_destroy(self.h) } }
}

// -------------------------------------------------------------------------------- checkpoint
/// `NativeSafetensorsLoader` (native_safetensors.rs:130-365): the GPTQ side.
pub struct Checkpoint { h: *mut ffi::Fb2Ckpt }
unsafe impl Send for Checkpoint {}
unsafe impl Sync for Checkpoint {}

impl Checkpoint {
    pub fn open(model_dir: &std::path::Path) -> Result<Self> {
        let dir = CString::new(model_dir.to_string_lossy().as_bytes()).map_err(|e| Fb2Error::Arg(e.to_string()))?;
        let mut h = ptr::null_mut();
        check(unsafe { ffi::fb2_ckpt_open(dir.as_ptr(), &mut h) })?;
        Ok(Self { h })
    }
    pub fn quant_config(&self) -> Result<Option<ffi::Fb2QuantConfig>> {
        let mut q = ffi::Fb2QuantConfig::default();
        check(unsafe { ffi::fb2_ckpt_quant_config(self.h, &mut q) })?;
        Ok(if q.present != 0 { Some(q) } else { None })
    }
    pub fn has_tensor(&self, name: &str) -> bool {
        CString::new(name).map_or(false, |n| unsafe { ffi::fb2_ckpt_has_tensor(self.h, n.as_ptr()) != 0 })
    }
    /// `load_linear` for a GPTQ layer, stored or fused from its parts (native_safetensors.rs:597-634).
    pub fn load_linear(&self, name: &str, device: i32) -> Result<GptqLinear> {
        let n = CString::new(name).map_err(|e| Fb2Error::Arg(e.to_string()))?;
        let mut h = ptr::null_mut();
        check(unsafe { ffi::fb2_ckpt_load_gptq_linear(self.h, n.as_ptr(), device, &mut h) })?;
        Ok(GptqLinear { h, _stack: None })
    }
    /// `load_stacked_gptq_experts` (native_safetensors.rs:388): returns `(stack, n_per_expert, k)`.
    pub fn load_stacked_gptq_experts(&self, expert_prefix_fmt: &str, expert_begin: usize, num_experts: usize,
                                     proj_names: &[&str], device: i32) -> Result<(ExpertStack, usize, usize)> {
        let fmt = CString::new(expert_prefix_fmt).map_err(|e| Fb2Error::Arg(e.to_string()))?;
        let names: Vec<CString> = proj_names.iter().map(|p| CString::new(*p).unwrap()).collect();
        let ptrs: Vec<*const std::os::raw::c_char> = names.iter().map(|c| c.as_ptr()).collect();
        let (mut h, mut n, mut k) = (ptr::null_mut(), 0usize, 0usize);
        check(unsafe {
            ffi::fb2_ckpt_load_stacked_gptq_experts(self.h, fmt.as_ptr(), expert_begin, num_experts, ptrs.as_ptr(), ptrs.len(),
                                                    device, &mut h, &mut n, &mut k)
        })?;
        Ok((ExpertStack { h }, n, k))
    }
}
impl Drop for Checkpoint {
    fn drop(&mut self) { unsafe { ffi::fb2_ckpt_close(self.h) } }
}
