//! Raw declarations of include/msm_b200.h.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct msm_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct msm_bases {
    _private: [u8; 0],
}

pub const MSM_OK: c_int = 0;
pub const MSM_ERR_INVALID: c_int = 1;
pub const MSM_ERR_CUDA: c_int = 2;
pub const MSM_ERR_BUSY: c_int = 3;
pub const MSM_ERR_ABORTED: c_int = 4;
pub const MSM_ERR_NO_DEVICE: c_int = 5;
pub const MSM_ERR_TOO_LARGE: c_int = 6;

pub const MSM_CURVE_BN254_G1: c_int = 0;
pub const MSM_CURVE_BLS12_381_G1: c_int = 1;

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct msm_timings {
    pub h2d_ms: f32,
    pub sort_ms: f32,
    pub accumulate_ms: f32,
    pub reduce_ms: f32,
    pub total_ms: f32,
    pub window_bits: u32,
    pub num_windows: u32,
    pub num_entries: u64,
    pub kernel_launches: u64,
    pub scatter_passes: u32,
    pub sub_batches: u32,
}

extern "C" {
    pub fn msm_device_count() -> c_int;
    pub fn msm_ctx_create(curve: c_int, device_ids: *const c_int, n_devices: c_int, out: *mut *mut msm_ctx) -> c_int;
    pub fn msm_ctx_destroy(ctx: *mut msm_ctx) -> c_int;
    pub fn msm_ctx_num_devices(ctx: *const msm_ctx) -> c_int;
    pub fn msm_set_abort_flag(ctx: *mut msm_ctx, flag: *const c_int) -> c_int;
    pub fn msm_last_error(ctx: *const msm_ctx) -> *const c_char;
    pub fn msm_last_timings(ctx: *const msm_ctx, out: *mut msm_timings) -> c_int;
    pub fn msm_bases_upload(ctx: *mut msm_ctx, xy_mont: *const c_void, n_points: usize, out: *mut *mut msm_bases) -> c_int;
    pub fn msm_bases_upload_sharded(ctx: *mut msm_ctx, xy_mont: *const c_void, n_points: usize, out: *mut *mut msm_bases) -> c_int;
    pub fn msm_bases_precompute(ctx: *mut msm_ctx, b: *mut msm_bases, window_bits: u32) -> c_int;
    pub fn msm_bases_precompute_chunked(ctx: *mut msm_ctx, b: *mut msm_bases, chunk_len: usize) -> c_int;
    /// 0 = off, 1 = lazy (default: the 2nd call of a shape builds the window table), 2 = eager.
    pub fn msm_bases_set_table_policy(ctx: *mut msm_ctx, b: *mut msm_bases, policy: c_int) -> c_int;
    /// Upload pipelining the engine would use for a call shape (host arithmetic, no device work): sub-batches of one
    /// MSM or task groups of a many-task row, and the factor their sizes grow by.
    pub fn msm_pipeline_shape(
        l: usize, n_lines: u32, num_chunks: u32, h2d_gbs: f32, device_ms: f32, sub_batches: *mut u32, growth: *mut f64,
    ) -> c_int;
    pub fn msm_bases_size_bytes(b: *const msm_bases) -> usize;
    pub fn msm_bases_free(b: *mut msm_bases) -> c_int;
    pub fn msm_multiple_multiexp(
        ctx: *mut msm_ctx, bases: *const msm_bases, scalars: *const c_void, l: usize, num_chunks: u32,
        window_hint: u32, neg_is_cheap: c_int, out_jacobian: *mut c_void,
    ) -> c_int;
    pub fn msm_multiexp(ctx: *mut msm_ctx, bases_xy_mont: *const c_void, scalars: *const c_void, n: usize, out_jacobian: *mut c_void) -> c_int;
    pub fn msm_multiexp_resident(ctx: *mut msm_ctx, bases: *const msm_bases, skip: usize, scalars: *const c_void, n: usize, out_jacobian: *mut c_void) -> c_int;
    pub fn msm_multiple_multiexp_montgomery(
        ctx: *mut msm_ctx, bases: *const msm_bases, scalars_mont: *const c_void, l: usize, num_chunks: u32,
        out_jacobian: *mut c_void,
    ) -> c_int;
    pub fn msm_ec_fft(ctx: *mut msm_ctx, jacobian_inout: *mut c_void, log_n: u32, omegas_mont: *const c_void, n_omegas: u32) -> c_int;
    pub fn msm_scalar_fft(ctx: *mut msm_ctx, fr_inout: *mut c_void, log_n: u32, omega_mont: *const c_void) -> c_int;
}
