//! FFI bindings for libvqb200.so, the B200-native engine of the vq hot path (include/vqb200.h).
//!
//! Twin of `src/core/hsdlib_ffi.rs` (reference lines 38-83): an `extern "C"` block, a status enum with the same codes as
//! hsdlib's (hsdlib.h:32-38) plus the two VqError kinds the engine distinguishes, and thin safe helpers.
//! Compiled only with the `b200` feature.  Every function of the header is declared here; `tests/test_shim_signatures.py`
//! in the engine repository keeps this block and the header in step.

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

use crate::core::error::{VqError, VqResult};

#[repr(C)]
pub struct VqbCtx { _private: [u8; 0] }
#[repr(C)]
pub struct VqbPq { _private: [u8; 0] }
#[repr(C)]
pub struct VqbTsvq { _private: [u8; 0] }

/// Stands in for `data.choose(&mut rng)` (src/core/vector.rs:450): global row that re-seeds the next empty cluster.
pub type ReseedFn = Option<unsafe extern "C" fn(user: *mut c_void, subspace: u32) -> u64>;
/// In-place float sum over all ranks of a device buffer, ordered after prior work on `cuda_stream`.
pub type AllreduceFn =
    Option<unsafe extern "C" fn(user: *mut c_void, buf: *mut f32, count: usize, cuda_stream: *mut c_void) -> c_int>;

pub const VQB_UPDATE_ORDERED: u32 = 0;
pub const VQB_UPDATE_FAST: u32 = 1;
pub const VQB_ASSIGN_AUTO: u32 = 0;
pub const VQB_ASSIGN_EXACT: u32 = 1;
pub const VQB_ASSIGN_TENSOR: u32 = 2;
pub const VQB_TRAIN_USE_COMM: u32 = 1;
pub const VQB_COMM_ID_BYTES: usize = 128;
/// EXTENSION metric id (max_i |a_i - b_i|): not a variant of the crate's `Distance` (src/core/distance.rs:8-17).
pub const VQB_CHEBYSHEV: c_int = 5;

#[repr(C)]
pub struct VqbTrainOpts {
    pub struct_size: u32,
    pub update_mode: u32,
    pub assign_mode: u32,
    pub flags: u32,
    pub reseed: ReseedFn,
    pub reseed_user: *mut c_void,
    pub allreduce: AllreduceFn,
    pub allreduce_user: *mut c_void,
    pub row_offset: u64,
    pub n_global: u64,
    pub iter_ms: *mut f32,
}

/// Status codes of the engine: hsdlib's (hsdlib_ffi.rs:10-17) plus EmptyInput / DimensionMismatch.
#[repr(i32)]
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum VqbStatus {
    Success = 0,
    ErrNullPtr = -1,
    ErrEmptyInput = -2,
    ErrInvalidInput = -3,
    ErrUnsupportedDevice = -4,
    ErrDimMismatch = -5,
    Failure = -99,
}

impl From<c_int> for VqbStatus {
    fn from(value: c_int) -> Self {
        match value {
            0 => VqbStatus::Success,
            -1 => VqbStatus::ErrNullPtr,
            -2 => VqbStatus::ErrEmptyInput,
            -3 => VqbStatus::ErrInvalidInput,
            -4 => VqbStatus::ErrUnsupportedDevice,
            -5 => VqbStatus::ErrDimMismatch,
            _ => VqbStatus::Failure,
        }
    }
}

unsafe extern "C" {
    pub fn vqb_ctx_create(device: c_int, out: *mut *mut VqbCtx) -> c_int;
    pub fn vqb_ctx_destroy(ctx: *mut VqbCtx) -> c_int;
    pub fn vqb_ctx_synchronize(ctx: *mut VqbCtx) -> c_int;
    pub fn vqb_ctx_stream(ctx: *mut VqbCtx) -> *mut c_void;
    pub fn vqb_ctx_set_stream(ctx: *mut VqbCtx, cuda_stream: *mut c_void) -> c_int;
    pub fn vqb_last_error(ctx: *mut VqbCtx) -> *const c_char;
    pub fn vqb_backend_name() -> *const c_char;
    pub fn vqb_ctx_launch_count(ctx: *mut VqbCtx) -> u64;
    pub fn vqb_malloc(ctx: *mut VqbCtx, bytes: usize, dptr: *mut *mut c_void) -> c_int;
    pub fn vqb_free(ctx: *mut VqbCtx, dptr: *mut c_void) -> c_int;
    pub fn vqb_host_alloc(ctx: *mut VqbCtx, bytes: usize, hptr: *mut *mut c_void) -> c_int;
    pub fn vqb_host_free(ctx: *mut VqbCtx, hptr: *mut c_void) -> c_int;
    pub fn vqb_memcpy(ctx: *mut VqbCtx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn vqb_comm_unique_id(id_out: *mut c_void) -> c_int;
    pub fn vqb_comm_init_rank(ctx: *mut VqbCtx, id: *const c_void, rank: c_int, world: c_int) -> c_int;
    pub fn vqb_comm_destroy(ctx: *mut VqbCtx) -> c_int;
    pub fn vqb_comm_info(ctx: *mut VqbCtx, rank: *mut c_int, world: *mut c_int) -> c_int;
    pub fn vqb_comm_allreduce(ctx: *mut VqbCtx, buf: *mut f32, count: usize) -> c_int;
    pub fn vqb_distance_batch(ctx: *mut VqbCtx, metric: c_int, a: *const f32, b: *const f32, rows: usize, n: usize, out: *mut f32) -> c_int;
    pub fn vqb_bq_quantize(ctx: *mut VqbCtx, x: *const f32, n: usize, threshold: f32, low: u8, high: u8, out: *mut u8) -> c_int;
    pub fn vqb_bq_dequantize(ctx: *mut VqbCtx, codes: *const u8, n: usize, low: u8, high: u8, out: *mut f32) -> c_int;
    pub fn vqb_sq_quantize(ctx: *mut VqbCtx, x: *const f32, n: usize, min: f32, max: f32, step: f32, levels: u32, out: *mut u8) -> c_int;
    pub fn vqb_sq_dequantize(ctx: *mut VqbCtx, codes: *const u8, n: usize, min: f32, step: f32, out: *mut f32) -> c_int;
    pub fn vqb_f16_dequantize(ctx: *mut VqbCtx, q: *const u16, n: usize, out: *mut f32) -> c_int;
    pub fn vqb_pq_train(ctx: *mut VqbCtx, x: *const f32, n: usize, dim: usize, m: usize, k: usize, max_iters: usize, init_idx: *const u64, opts: *const VqbTrainOpts, codebooks: *mut f32, iters_run: *mut u32) -> c_int;
    pub fn vqb_pq_assign_train(ctx: *mut VqbCtx, x: *const f32, n: usize, dim: usize, m: usize, k: usize, codebooks: *const f32, assign_mode: u32, codes_out: *mut u32) -> c_int;
    pub fn vqb_pq_train_step(ctx: *mut VqbCtx, x: *const f32, n: usize, dim: usize, m: usize, k: usize, codebooks_inout: *mut f32, opts: *const VqbTrainOpts, changed_out: *mut u32, counts_out: *mut u32) -> c_int;
    pub fn vqb_pq_create(ctx: *mut VqbCtx, codebooks: *const f32, m: usize, k: usize, sub_dim: usize, metric: c_int, out: *mut *mut VqbPq) -> c_int;
    pub fn vqb_pq_destroy(pq: *mut VqbPq) -> c_int;
    pub fn vqb_pq_codebooks(pq: *mut VqbPq, out: *mut f32) -> c_int;
    pub fn vqb_pq_encode(pq: *mut VqbPq, x: *const f32, n: usize, assign_mode: u32, codes_out: *mut c_void, code_bytes: u32, recon_out: *mut u16) -> c_int;
    pub fn vqb_pq_decode(pq: *mut VqbPq, codes: *const c_void, code_bytes: u32, n: usize, out: *mut f32) -> c_int;
    pub fn vqb_debug_tc_scores(ctx: *mut VqbCtx, cosine: c_int, x: *const f32, n: usize, dim: usize, m: usize, k: usize, codebooks: *const f32, sub: c_int, scores_out: *mut f32, rescans_out: *mut u64, codes_out: *mut u32) -> c_int;
    pub fn vqb_debug_tc_timeline(ctx: *mut VqbCtx, x: *const f32, n: usize, dim: usize, m: usize, k: usize, codebooks: *const f32, ts_out: *mut u64, units: c_int) -> c_int;
    pub fn vqb_debug_tc_variant(variant: c_int) -> c_int;
    pub fn vqb_tsvq_train(ctx: *mut VqbCtx, x: *const f32, n: usize, dim: usize, max_depth: usize, metric: c_int, out: *mut *mut VqbTsvq) -> c_int;
    pub fn vqb_tsvq_create(ctx: *mut VqbCtx, centroids: *const f32, left: *const i32, right: *const i32, n_nodes: usize, dim: usize, metric: c_int, out: *mut *mut VqbTsvq) -> c_int;
    pub fn vqb_tsvq_destroy(t: *mut VqbTsvq) -> c_int;
    pub fn vqb_tsvq_num_nodes(t: *mut VqbTsvq, n_nodes: *mut usize, dim: *mut usize) -> c_int;
    pub fn vqb_tsvq_export(t: *mut VqbTsvq, centroids: *mut f32, left: *mut i32, right: *mut i32, split_dim: *mut i32, median: *mut f32, count: *mut u64) -> c_int;
    pub fn vqb_tsvq_encode(t: *mut VqbTsvq, x: *const f32, n: usize, leaf_out: *mut u32, recon_out: *mut u16) -> c_int;
}

/// Message of the last failure on `ctx` (owned by the context).
pub fn last_error(ctx: *mut VqbCtx) -> String {
    // SAFETY: vqb_last_error returns a NUL-terminated string owned by the context (or a static one for a null context).
    unsafe { CStr::from_ptr(vqb_last_error(ctx)) }.to_string_lossy().into_owned()
}

/// Status -> VqError (src/core/error.rs:5-28).  The shim validates in Rust first, so the parameter errors below only
/// appear if the two sides disagree.
pub fn check(ctx: *mut VqbCtx, rc: c_int) -> VqResult<()> {
    match VqbStatus::from(rc) {
        VqbStatus::Success => Ok(()),
        VqbStatus::ErrEmptyInput => Err(VqError::EmptyInput),
        VqbStatus::ErrInvalidInput | VqbStatus::ErrDimMismatch => {
            Err(VqError::InvalidParameter { parameter: "engine", reason: last_error(ctx) })
        }
        VqbStatus::ErrUnsupportedDevice => {
            Err(VqError::FfiError("no sm_100 (B200) GPU: the b200 feature has no CPU fallback".to_string()))
        }
        _ => Err(VqError::FfiError(last_error(ctx))),
    }
}

/// Replaces `get_simd_backend()` (src/core/hsdlib_ffi.rs:144-155).
pub fn get_backend() -> String {
    // SAFETY: returns a static NUL-terminated string.
    unsafe { CStr::from_ptr(vqb_backend_name()) }.to_string_lossy().into_owned()
}

/// One engine context per process and device, created on first use (quantizers are `Send + Sync`: every entry point
/// serialises on the context's own mutex).
pub struct Engine(pub *mut VqbCtx);
unsafe impl Send for Engine {}
unsafe impl Sync for Engine {}

pub fn engine() -> VqResult<&'static Engine> {
    use std::sync::OnceLock;
    static ENGINE: OnceLock<Result<Engine, String>> = OnceLock::new();
    let e = ENGINE.get_or_init(|| {
        let device = std::env::var("VQ_B200_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut ctx: *mut VqbCtx = std::ptr::null_mut();
        // SAFETY: `ctx` is a valid out-pointer.
        let rc = unsafe { vqb_ctx_create(device, &mut ctx) };
        if rc == 0 { Ok(Engine(ctx)) } else { Err(format!("vqb_ctx_create({device}) failed with status {rc}")) }
    });
    e.as_ref().map_err(|m| VqError::FfiError(m.clone()))
}
