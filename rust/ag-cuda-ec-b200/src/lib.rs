//! Same public surface as the reference's `ag_cuda_ec` (ag-cuda-ec/src/{lib,multiexp,pairing_suite}.rs)
//! and `ec_gpu_proxy::multiexp::MultiexpKernel` (ec-gpu-proxy/src/multiexp.rs:256-403), implemented
//! over libmsm_b200.so.  NOT COMPILED in this repository (no Rust toolchain in the image).
//!
//! Layout contract (unchanged from the reference, ag-types/src/impls.rs:7-58):
//!   `<Affine as GpuRepr>::Repr = [Fq; 2]`  -> 2 x N little-endian u32 limbs, Montgomery form
//!   `<Scalar as PrimeFieldRepr>::Repr = BigInt<4>` -> 32 bytes canonical little-endian
//!   `Curve = Projective { x, y, z }` (Jacobian, Montgomery)  <- the engine's output bytes
pub mod pairing_suite {
    #[cfg(feature = "bn254")]
    pub use ark_bn254::{Fr as Scalar, G1Affine as Affine, G1Projective as Curve};
    #[cfg(feature = "bls12-381")]
    pub use ark_bls12_381::{Fr as Scalar, G1Affine as Affine, G1Projective as Curve};
    #[cfg(feature = "bn254")]
    pub const CURVE_ID: i32 = msm_b200_sys::MSM_CURVE_BN254_G1;
    #[cfg(feature = "bls12-381")]
    pub const CURVE_ID: i32 = msm_b200_sys::MSM_CURVE_BLS12_381_G1;
}

use ag_types::{GpuRepr, PrimeFieldRepr};
use ark_std::Zero;
use msm_b200_sys as sys;
use once_cell::sync::Lazy;
use pairing_suite::{Affine, Curve, Scalar, CURVE_ID};
use std::{cell::RefCell, ffi::CStr, os::raw::c_void, ptr};

/// rustacuda::error::CudaError stand-in: the variants this path can produce.
#[derive(Debug, Clone, PartialEq, Eq)]
pub enum CudaError {
    ContextAlreadyInUse, // MSM_ERR_BUSY  (ag-cuda-proxy/src/context.rs:20-27)
    InvalidValue,        // MSM_ERR_INVALID / MSM_ERR_TOO_LARGE
    NoDevice,            // MSM_ERR_NO_DEVICE
    UnknownError(String),
}
pub type CudaResult<T> = Result<T, CudaError>;

fn check(ctx: *const sys::msm_ctx, rc: i32) -> CudaResult<()> {
    match rc {
        sys::MSM_OK => Ok(()),
        sys::MSM_ERR_BUSY => Err(CudaError::ContextAlreadyInUse),
        sys::MSM_ERR_INVALID | sys::MSM_ERR_TOO_LARGE => Err(CudaError::InvalidValue),
        sys::MSM_ERR_NO_DEVICE => Err(CudaError::NoDevice),
        _ => Err(CudaError::UnknownError(unsafe { CStr::from_ptr(sys::msm_last_error(ctx)) }.to_string_lossy().into_owned())),
    }
}

/// CudaWorkspace (ag-cuda-proxy/src/module.rs:13-62): one engine context on device 0.
pub struct CudaWorkspace(*mut sys::msm_ctx);
unsafe impl Send for CudaWorkspace {}
unsafe impl Sync for CudaWorkspace {} // the context itself rejects concurrent use with MSM_ERR_BUSY
impl CudaWorkspace {
    fn new() -> Self {
        let mut ctx = ptr::null_mut();
        let rc = unsafe { sys::msm_ctx_create(CURVE_ID, ptr::null(), 1, &mut ctx) };
        check(ptr::null(), rc).unwrap(); // the reference unwraps too (ag-cuda-ec/src/lib.rs:13)
        CudaWorkspace(ctx)
    }
}
impl Drop for CudaWorkspace {
    fn drop(&mut self) {
        unsafe { sys::msm_ctx_destroy(self.0) };
    }
}

// construct_workspace! (ag-cuda-workspace-macro/src/lib.rs:58-78)
pub static GLOBAL: Lazy<CudaWorkspace> = Lazy::new(CudaWorkspace::new);
thread_local! { pub static LOCAL: RefCell<Option<CudaWorkspace>> = RefCell::new(None); }
pub fn init_global_workspace() {
    Lazy::force(&GLOBAL);
}
pub fn init_local_workspace() {
    LOCAL.with(|l| l.borrow_mut().get_or_insert_with(CudaWorkspace::new));
}

/// DeviceData (ag-cuda-proxy/src/params.rs:173-218): resident bases, freed on drop.  The handle keeps its
/// context alive (msm_b200.h, "Lifetime"): dropping it after its workspace -- a thread-local workspace that died
/// before a `DeviceData` sent to another thread -- is safe.
pub struct DeviceData(*mut sys::msm_bases);
unsafe impl Send for DeviceData {}
impl DeviceData {
    pub fn size(&self) -> usize {
        unsafe { sys::msm_bases_size_bytes(self.0) }
    }
}
impl Drop for DeviceData {
    fn drop(&mut self) {
        unsafe { sys::msm_bases_free(self.0) };
    }
}

pub mod multiexp {
    use super::*;

    // Window tables need no call here: the second `multiple_multiexp_*` call with the same
    // `(exponents.len(), num_chunks)` on one `DeviceData` builds the table that shape wants (msm_b200.h, "Window
    // tables by policy").  The two functions below only override that policy.

    /// Engine extension: 0 = never build a table behind the caller's back, 1 = lazily (default), 2 = now.
    pub fn set_table_policy_st(bases_gpu: &DeviceData, policy: i32) -> CudaResult<()> {
        check(GLOBAL.0, unsafe { sys::msm_bases_set_table_policy(GLOBAL.0, bases_gpu.0, policy) })
    }

    /// Engine extension: window table for calls whose tasks have `chunk_len` points each (the
    /// per-segment commitment and AMT shapes, ag-cuda-ec/benches/{multiexp,amt}.rs).  Results are unchanged.
    pub fn precompute_chunked_st(bases_gpu: &DeviceData, chunk_len: usize) -> CudaResult<()> {
        check(GLOBAL.0, unsafe { sys::msm_bases_precompute_chunked(GLOBAL.0, bases_gpu.0, chunk_len) })
    }

    fn upload(ws: &CudaWorkspace, bases: &[Affine]) -> CudaResult<DeviceData> {
        // ag-cuda-ec/src/multiexp.rs:15-16: strip the `infinity` flag, identity -> (0,0)
        let repr: Vec<<Affine as GpuRepr>::Repr> = bases.iter().map(GpuRepr::to_gpu_repr).collect();
        let mut h = ptr::null_mut();
        check(ws.0, unsafe { sys::msm_bases_upload(ws.0, repr.as_ptr() as *const c_void, repr.len(), &mut h) })?;
        Ok(DeviceData(h))
    }

    fn run(
        ws: &CudaWorkspace, bases_gpu: &DeviceData, exponents: &[<Scalar as PrimeFieldRepr>::Repr],
        num_chunks: usize, window_size: usize, neg_is_cheap: bool,
    ) -> CudaResult<Vec<Curve>> {
        let num_bases = bases_gpu.size() / std::mem::size_of::<<Affine as GpuRepr>::Repr>();
        let num_lines = num_bases / exponents.len(); // ag-cuda-ec/src/multiexp.rs:28-30
        let mut output = vec![Curve::zero(); num_chunks * num_lines];
        check(ws.0, unsafe {
            sys::msm_multiple_multiexp(
                ws.0, bases_gpu.0, exponents.as_ptr() as *const c_void, exponents.len(), num_chunks as u32,
                window_size as u32, neg_is_cheap as i32, output.as_mut_ptr() as *mut c_void,
            )
        })?;
        Ok(output)
    }

    // #[auto_workspace] expansion (ag-cuda-workspace-macro/src/lib.rs:8-55)
    pub fn upload_multiexp_bases_st(bases: &[Affine]) -> CudaResult<DeviceData> {
        upload(&GLOBAL, bases)
    }
    pub fn upload_multiexp_bases_mt(bases: &[Affine]) -> CudaResult<DeviceData> {
        init_local_workspace();
        LOCAL.with(|l| upload(l.borrow().as_ref().unwrap(), bases))
    }
    pub fn multiple_multiexp_st(
        bases_gpu: &DeviceData, exponents: &[<Scalar as PrimeFieldRepr>::Repr], num_chunks: usize,
        window_size: usize, neg_is_cheap: bool,
    ) -> CudaResult<Vec<Curve>> {
        run(&GLOBAL, bases_gpu, exponents, num_chunks, window_size, neg_is_cheap)
    }
    pub fn multiple_multiexp_mt(
        bases_gpu: &DeviceData, exponents: &[<Scalar as PrimeFieldRepr>::Repr], num_chunks: usize,
        window_size: usize, neg_is_cheap: bool,
    ) -> CudaResult<Vec<Curve>> {
        init_local_workspace();
        LOCAL.with(|l| run(l.borrow().as_ref().unwrap(), bases_gpu, exponents, num_chunks, window_size, neg_is_cheap))
    }
}

/// ag_cuda_ec::ec_fft (ag-cuda-ec/src/ec_fft.rs:13-99): same signature, the transform runs in msm_ec_fft.
pub mod ec_fft {
    use super::*;

    fn run(ws: &CudaWorkspace, input: &mut Vec<Curve>, omegas: &[Scalar]) -> CudaResult<()> {
        let n = input.len();
        let log_n = n.ilog2();
        assert_eq!(n, 1 << log_n);
        // Vec<Curve> is the {x, y, z} Montgomery record the engine reads and writes in place;
        // Scalar is arkworks' Fp<MontBackend, 4>: 4 x u64 little-endian, Montgomery form
        check(ws.0, unsafe {
            sys::msm_ec_fft(ws.0, input.as_mut_ptr() as *mut c_void, log_n, omegas.as_ptr() as *const c_void, omegas.len() as u32)
        })
    }
    pub fn radix_ec_fft_st(input: &mut Vec<Curve>, omegas: &[Scalar]) -> CudaResult<()> {
        run(&GLOBAL, input, omegas)
    }
    pub fn radix_ec_fft_mt(input: &mut Vec<Curve>, omegas: &[Scalar]) -> CudaResult<()> {
        init_local_workspace();
        LOCAL.with(|l| run(l.borrow().as_ref().unwrap(), input, omegas))
    }
}

/// ec_gpu_proxy::multiexp (legacy multi-GPU entry), ec-gpu-proxy/src/multiexp.rs:256-403.
pub mod legacy {
    use super::*;
    use std::sync::{atomic::{AtomicI32, Ordering}, Arc};

    #[derive(Debug)]
    pub enum EcError {
        Simple(&'static str),
        Aborted,
        GpuTools(String),
    }
    pub type EcResult<T> = Result<T, EcError>;

    pub struct MultiexpKernel<'a> {
        ctx: *mut sys::msm_ctx,
        maybe_abort: Option<&'a (dyn Fn() -> bool + Send + Sync)>,
        abort_flag: Box<AtomicI32>,
    }

    impl<'a> MultiexpKernel<'a> {
        /// `devices`: CUDA ordinals (the reference passes rust_gpu_tools Programs/Devices).
        pub fn create(devices: &[i32]) -> EcResult<Self> {
            Self::create_optional_abort(devices, None)
        }
        pub fn create_with_abort(devices: &[i32], maybe_abort: &'a (dyn Fn() -> bool + Send + Sync)) -> EcResult<Self> {
            Self::create_optional_abort(devices, Some(maybe_abort))
        }
        fn create_optional_abort(devices: &[i32], maybe_abort: Option<&'a (dyn Fn() -> bool + Send + Sync)>) -> EcResult<Self> {
            let mut ctx = ptr::null_mut();
            let rc = unsafe { sys::msm_ctx_create(CURVE_ID, devices.as_ptr(), devices.len() as i32, &mut ctx) };
            if rc != sys::MSM_OK {
                return Err(EcError::Simple("No working GPUs found!")); // multiexp.rs:305-307
            }
            let abort_flag = Box::new(AtomicI32::new(0));
            unsafe { sys::msm_set_abort_flag(ctx, abort_flag.as_ptr() as *const i32) };
            Ok(MultiexpKernel { ctx, maybe_abort, abort_flag })
        }
        pub fn num_kernels(&self) -> usize {
            unsafe { sys::msm_ctx_num_devices(self.ctx) as usize }
        }
        /// multiexp.rs:372-400: uses `bases[skip..skip + exps.len()]`; the pool argument is kept for
        /// signature compatibility (the engine runs one host thread per GPU internally).
        pub fn multiexp<W>(
            &mut self, _pool: &W, bases_arc: Arc<Vec<Affine>>, exps: Arc<Vec<<Scalar as PrimeFieldRepr>::Repr>>,
            skip: usize,
        ) -> EcResult<Curve> {
            if let Some(f) = self.maybe_abort {
                if f() {
                    self.abort_flag.store(1, Ordering::SeqCst);
                    return Err(EcError::Aborted);
                }
            }
            let bases = &bases_arc[skip..(skip + exps.len())];
            let repr: Vec<<Affine as GpuRepr>::Repr> = bases.iter().map(GpuRepr::to_gpu_repr).collect();
            let mut out = Curve::zero();
            let rc = unsafe {
                sys::msm_multiexp(self.ctx, repr.as_ptr() as *const c_void, exps.as_ptr() as *const c_void, exps.len(),
                                  &mut out as *mut Curve as *mut c_void)
            };
            match rc {
                sys::MSM_OK => Ok(out),
                sys::MSM_ERR_ABORTED => Err(EcError::Aborted),
                _ => Err(EcError::GpuTools(unsafe { CStr::from_ptr(sys::msm_last_error(self.ctx)) }.to_string_lossy().into_owned())),
            }
        }
    }
    impl<'a> Drop for MultiexpKernel<'a> {
        fn drop(&mut self) {
            unsafe { sys::msm_ctx_destroy(self.ctx) };
        }
    }
}
