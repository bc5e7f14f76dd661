/* msm_b200.h -- C ABI of the B200-native MSM engine (libmsm_b200.so).
 *
 * The reference has no FFI on this path: its Rust host code reaches the device through
 * rustacuda's cuLaunchKernel by mangled kernel name (ag-cuda-ec/src/multiexp.rs:57-72).  This
 * header is therefore the boundary a thin Rust shim (rust/ in this repo) binds
 * one level up, under the two host APIs that stay frozen:
 *
 *   ag_cuda_ec::{init_global_workspace, init_local_workspace}   ag-cuda-workspace-macro/src/lib.rs:58-78
 *   ag_cuda_ec::multiexp::upload_multiexp_bases_{st,mt}          ag-cuda-ec/src/multiexp.rs:12-19
 *   ag_cuda_ec::multiexp::multiple_multiexp_{st,mt}              ag-cuda-ec/src/multiexp.rs:22-81
 *   ec_gpu_proxy::multiexp::MultiexpKernel::{create,create_with_abort,multiexp,num_kernels}
 *                                                                ec-gpu-proxy/src/multiexp.rs:266-403
 *
 * Memory layouts (identical to the reference):
 *   base point   {x, y}: 2 x N little-endian u32 limbs, Montgomery form (R = 2^(32N));
 *                N = 8 (BN254, 64 B/point), N = 12 (BLS12-381, 96 B/point); identity = all zero
 *                (GpuRepr for Affine, ag-types/src/impls.rs:48-58)
 *   scalar       32 bytes little-endian, canonical (non-Montgomery) integer < r
 *                (PrimeFieldRepr::to_bigint -> BigInt<4>, ag-types/src/impls.rs:7-18)
 *   result point {x, y, z}: Jacobian, Montgomery form, infinity <=> z == 0
 *                (POINT_jacobian, ag-build/cl/ec.cl:10-14).  Any representative of the group
 *                element may be returned; callers compare as group elements / after into_affine()
 *                (ag-cuda-ec/src/multiexp.rs:125, ec-gpu-proxy/tests/multiexp.rs:99).
 *
 * Conventions: every function returns 0 (MSM_OK) or an msm_status; no exceptions cross the
 * boundary; every handle is owned by the caller and freed by the matching *_destroy / *_free.
 * A context is not re-entrant: a second concurrent call on the same context returns
 * MSM_ERR_BUSY, mirroring CudaError::ContextAlreadyInUse (ag-cuda-proxy/src/context.rs:20-27).
 * Distinct contexts are independent (the reference's per-thread "_mt" workspaces).
 * There is no CPU fallback anywhere behind this interface.
 */
#ifndef MSM_B200_H
#define MSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  MSM_OK = 0,
  MSM_ERR_INVALID = 1,     /* bad argument (null pointer, size mismatch, unknown curve) */
  MSM_ERR_CUDA = 2,        /* CUDA runtime failure; msm_last_error() has the text (EcError::GpuTools) */
  MSM_ERR_BUSY = 3,        /* context already in use (CudaError::ContextAlreadyInUse) */
  MSM_ERR_ABORTED = 4,     /* abort flag observed (EcError::Aborted, ec-gpu-proxy/src/multiexp.rs:140-144) */
  MSM_ERR_NO_DEVICE = 5,   /* "No working GPUs found!" (ec-gpu-proxy/src/multiexp.rs:305-307) */
  MSM_ERR_TOO_LARGE = 6    /* sizes beyond the u32 index space the reference also assumes */
} msm_status;

/* G2: the same engine over Fq2 = Fq[u]/(u^2+1); a coordinate is {c0, c1}, each in
 * the base field's layout (GpuRepr for the quadratic extension, ag-types/src/impls.rs:36-46): points are
 * 128 B (BN254) / 192 B (BLS12-381) affine and 192 / 288 B Jacobian.  Every entry point below works for all
 * four ids; "N" in the layout notes is then 16 / 24. */
typedef enum { MSM_CURVE_BN254_G1 = 0, MSM_CURVE_BLS12_381_G1 = 1, MSM_CURVE_BN254_G2 = 2, MSM_CURVE_BLS12_381_G2 = 3 } msm_curve;

typedef struct msm_ctx msm_ctx;     /* = CudaWorkspace / MultiexpKernel (one or more devices) */
typedef struct msm_bases msm_bases; /* = DeviceData holding resident bases */

/* Per-phase device times of the last call on device 0 of the context, milliseconds (CUDA events). */
typedef struct {
  float h2d_ms;        /* scalar upload */
  float sort_ms;       /* digit decomposition + bucket sort */
  float accumulate_ms; /* bucket accumulation (the IMAD-bound phase) */
  float reduce_ms;     /* bucket reduction + window combine (+ cross-GPU gather) */
  float total_ms;      /* first kernel to result ready */
  uint32_t window_bits;
  uint32_t num_windows;
  uint64_t num_entries;   /* non-zero digits sorted */
  uint64_t kernel_launches;
  uint32_t scatter_passes; /* bucket-range passes of the scatter kernel */
  uint32_t sub_batches;    /* > 1: host scalars were uploaded and processed in pipelined sub-batches */
} msm_timings;

/* ---- contexts -------------------------------------------------------------------------- */
/* Number of visible CUDA devices (<= 0: none). */
int msm_device_count(void);
/* Create a context over `n_devices` devices (device_ids == NULL: devices 0..n_devices-1;
 * n_devices == 0: all visible devices).  Replaces CudaWorkspace::from_bytes
 * (ag-cuda-proxy/src/module.rs:24-42) and MultiexpKernel::create
 * (ec-gpu-proxy/src/multiexp.rs:266-322). */
int msm_ctx_create(int curve, const int* device_ids, int n_devices, msm_ctx** out);
int msm_ctx_destroy(msm_ctx* ctx);
/* MultiexpKernel::num_kernels (ec-gpu-proxy/src/multiexp.rs:402). */
int msm_ctx_num_devices(const msm_ctx* ctx);
/* Cooperative cancel, polled between phases: replaces the maybe_abort callback of
 * MultiexpKernel::create_with_abort.  flag may be NULL to clear. */
int msm_set_abort_flag(msm_ctx* ctx, const volatile int* flag);
/* Text of the last error on this context (or of context creation when ctx == NULL). */
const char* msm_last_error(const msm_ctx* ctx);
int msm_last_timings(const msm_ctx* ctx, msm_timings* out);
/* Window-size override for experiments (0 = automatic).  Results never depend on it. */
int msm_set_window_bits(msm_ctx* ctx, uint32_t c);
/* Host-side view of the launch plan the engine makes for one multiple_multiexp call of the given shape: window size,
 * sub-batches of the pipelined scalar upload (parts of one MSM, or groups of whole tasks of a many-task row), slice
 * length of the bucket kernel and the waves it fills.  Pure host arithmetic: no device work, callable without a GPU
 * (the kernel's blocks per SM then default to 4 x 148 SMs).  table_window_bits != 0: the bases are a window table of
 * that window size; sub_batches / growth: what the call would pass (1 and 2.0 for device-resident scalars).
 * A test and diagnostics aid with no counterpart in the reference. */
typedef struct {
  uint32_t window_bits, num_windows;
  uint32_t buckets;          /* of the whole call */
  uint32_t sub_batches;      /* after the plan's own limits (one line of bases, >= 2 tasks per group) */
  uint32_t by_task;          /* 1: the sub-batches are groups of whole tasks with bucket ranges of their own */
  uint32_t sub_first[9];     /* scalar index where sub-batch k starts; [sub_batches] = scalars used */
  uint32_t slice_len;        /* sorted digits per thread of the bucket kernel (longest sub-batch) */
  uint32_t slices;           /* threads of the bucket kernel for the whole row */
  uint32_t wave_slices;      /* threads of one wave of the bucket kernel */
  uint32_t waves;            /* waves of the longest sub-batch; 0: the slice length was imposed */
  uint32_t sort_mode;        /* 0 single-level sort, 2 binned sort */
  uint32_t reduce_q;         /* buckets per reduction thread */
  uint64_t digits_max;       /* upper bound of the digits sorted */
  uint64_t scratch_bytes;    /* per-call device scratch */
} msm_plan_info;
int msm_plan_describe(int curve, size_t L, uint32_t n_lines, uint32_t num_chunks, uint32_t table_window_bits,
                      uint32_t sub_batches, double growth, msm_plan_info* out);
/* How the engine pipelines the upload of the host scalars of one multiple_multiexp call: the number of sub-batches
 * (parts of one MSM, or groups of whole tasks of a many-task row) and the factor their sizes grow by.  h2d_gbs and
 * device_ms are what a workspace has measured on earlier calls of the shape (upload rate in GB/s, shortest device time
 * in ms); 0 = nothing measured yet.  Pure host arithmetic, like msm_plan_describe. */
int msm_pipeline_shape(size_t L, uint32_t n_lines, uint32_t num_chunks, float h2d_gbs, float device_ms,
                       uint32_t* sub_batches, double* growth);
/* Name of the field implementation behind this context ("bn254/u29", "bn254/sat32", "bls12-381/sat32"). */
const char* msm_field_impl(const msm_ctx* ctx);
/* Run device 0's work on a caller-owned cudaStream_t (e.g. the framework's current stream), so that
 * the caller's own stream-ordered work and events see it; NULL restores a private stream.  The
 * reference creates a fresh NON_BLOCKING stream per kernel (ag-cuda-proxy/src/module.rs:56-62). */
int msm_set_stream(msm_ctx* ctx, void* cuda_stream);

/* ---- resident bases -------------------------------------------------------------------- */
/* Lifetime: an msm_bases keeps its context alive.  msm_ctx_destroy on a context that still has live
 * msm_bases only closes it (every later call on it returns MSM_ERR_INVALID); the device resources go
 * when the last msm_bases_free has run.  So Drop order in the Rust shim (a thread-local workspace
 * dying before a DeviceData that was sent to another thread) is harmless.
 *
 * upload_multiexp_bases (ag-cuda-ec/src/multiexp.rs:12-19): copy n_points {x,y} Montgomery
 * points from host memory to device 0 of the context and keep them resident.
 *
 * Window tables by policy: the reference API has no "precompute" step, so the engine decides by itself.
 * Default MSM_TABLE_LAZY: the second msm_multiple_multiexp[_device] call with the same (L, num_chunks)
 * shape on the same msm_bases builds the window table described at msm_bases_precompute below (once;
 * 0.2 - 0.9 s for 2^24 points), provided it fits the budget: table bytes <= MSM_B200_TABLE_BUDGET_GB
 * (default: half of the device's memory) and <= 80 % of the memory free at that moment; every later
 * call of that shape uses it.  A one-off call therefore never pays for a table, a prover that reuses
 * its SRS gets the fast path without calling anything outside the reference's API.  A different
 * shape seen twice in a row replaces the table.  Environment MSM_B200_TABLE=off|lazy|eager sets the
 * default policy of new msm_bases. */
int msm_bases_upload(msm_ctx* ctx, const void* xy_mont, size_t n_points, msm_bases** out);
/* Same, but split contiguously over all devices of the context, ceil(n/devices) points each:
 * the partition MultiexpKernel::parallel_multiexp uses (ec-gpu-proxy/src/multiexp.rs:329-337). */
int msm_bases_upload_sharded(msm_ctx* ctx, const void* xy_mont, size_t n_points, msm_bases** out);
/* Same as msm_bases_upload with the source points already in device memory of device 0 (same
 * {x,y} Montgomery layout); the engine keeps its own resident copy, the source may be freed. */
int msm_bases_from_device(msm_ctx* ctx, const void* d_xy_mont, size_t n_points, msm_bases** out);
/* Optional: build the window table T[w][i] = 2^(c w) P_i next to the resident copy
 * (W = ceil((bits+1)/c) times its size; window_bits == 0 lets the engine choose c for one MSM over
 * the whole shard; 8 <= c <= 24 otherwise).  With the table every window of a task lands in ONE
 * bucket set -- no per-window bucket arrays, no Horner doublings.  Used by calls that cover a whole
 * shard as one MSM, and by chunked / multi-line calls (the per-segment commitment and AMT shapes,
 * ag-cuda-ec/benches/{multiexp,amt}.rs) whenever the engine's cost model says the folded form is
 * cheaper than the plain one; every other call keeps using the plain resident copy.
 * Pure optimisation: results are unchanged.  The reference has no counterpart (its kernel walks
 * the windows of each scalar in separate threads, ag-build/cl/multiexp.cl:95-119). */
int msm_bases_precompute(msm_ctx* ctx, msm_bases* b, uint32_t window_bits);
/* Same with the window size chosen for tasks of chunk_len points each (multiple_multiexp with
 * num_chunks = L / chunk_len), e.g. 4096 for ag-cuda-ec/benches/multiexp.rs:19-22. */
int msm_bases_precompute_chunked(msm_ctx* ctx, msm_bases* b, size_t chunk_len);
/* Window size of the table (0: none). */
uint32_t msm_bases_table_window(const msm_bases* b);
/* Policy for this handle (overrides MSM_B200_TABLE): MSM_TABLE_OFF drops nothing but never builds;
 * MSM_TABLE_LAZY as described above; MSM_TABLE_EAGER builds the whole-shard table now. */
typedef enum { MSM_TABLE_OFF = 0, MSM_TABLE_LAZY = 1, MSM_TABLE_EAGER = 2 } msm_table_policy;
int msm_bases_set_table_policy(msm_ctx* ctx, msm_bases* b, int policy);
/* DeviceData::size (ag-cuda-proxy/src/params.rs:209): bytes. */
size_t msm_bases_size_bytes(const msm_bases* b);
size_t msm_bases_num_points(const msm_bases* b);
int msm_bases_free(msm_bases* b);

/* ---- the hot path ---------------------------------------------------------------------- */
/* multiple_multiexp (ag-cuda-ec/src/multiexp.rs:22-81).  `scalars` = one row of L scalars in
 * host memory; bases = num_lines * L resident points (num_lines = points / L).  Each line is cut
 * into num_chunks chunks of L / num_chunks points; out_jacobian receives num_lines * num_chunks
 * points, task (line, chunk) at index line * num_chunks + chunk (ag-build/cl/multiexp.cl:262):
 *     out[line*num_chunks + chunk] = sum_{i in chunk} scalars[i] * bases[line*L + i].
 * window_hint / neg_is_cheap are accepted for signature compatibility; the engine chooses its
 * own signed-digit window and the result does not depend on them (reference test
 * ag-cuda-ec/src/multiexp.rs:115-143).  Synchronous: the result is in out_jacobian on return. */
int msm_multiple_multiexp(msm_ctx* ctx, const msm_bases* bases, const void* scalars, size_t L,
                          uint32_t num_chunks, uint32_t window_hint, int neg_is_cheap,
                          void* out_jacobian);
/* Same with the scalar row and the result buffer in device memory of device 0 (no host copies;
 * used for device-resident timing).  Returns after the result is complete. */
int msm_multiple_multiexp_device(msm_ctx* ctx, const msm_bases* bases, const void* d_scalars,
                                 size_t L, uint32_t num_chunks, void* d_out_jacobian);

/* Exponents in Montgomery form (arkworks' in-memory Fr, 4 x u64 little-endian) -> canonical
 * integers on the device: replaces the host pass PrimeFieldRepr::to_bigint
 * (ag-types/src/impls.rs:7-18) that the reference's benches time separately
 * (ag-cuda-ec/benches/multiexp.rs:28-36).  d_in == d_out is allowed. */
int msm_scalars_from_montgomery_device(msm_ctx* ctx, const void* d_scalars_mont, size_t n, void* d_scalars_out);
/* msm_multiple_multiexp for a host row of exponents still in Montgomery form. */
int msm_multiple_multiexp_montgomery(msm_ctx* ctx, const msm_bases* bases, const void* scalars_mont, size_t L,
                                     uint32_t num_chunks, void* out_jacobian);

/* The same call repeated `repeats` times back to back with CUDA events recorded on the launching
 * stream around the whole sequence (total_ms) and around every bucket-accumulation kernel (their
 * sum in accumulate_ms).  Measurement aid for bench.py; results as above. */
int msm_multiple_multiexp_device_timed(msm_ctx* ctx, const msm_bases* bases, const void* d_scalars,
                                       size_t L, uint32_t num_chunks, void* d_out_jacobian,
                                       uint32_t repeats, float* total_ms, float* accumulate_ms);

/* MultiexpKernel::multiexp (ec-gpu-proxy/src/multiexp.rs:372-400): one MSM over n host points
 * and n host scalars (the Rust shim applies `skip` as a pointer offset), split over all devices
 * of the context in contiguous chunks of ceil(n/devices); per-device partial points are gathered
 * on device 0 over NVLink peer copies and summed there.  out_jacobian = one point. */
int msm_multiexp(msm_ctx* ctx, const void* bases_xy_mont, const void* scalars, size_t n,
                 void* out_jacobian);
/* Same with bases resident from msm_bases_upload_sharded (or msm_bases_upload on a one-device
 * context); uses bases[skip .. skip + n). */
int msm_multiexp_resident(msm_ctx* ctx, const msm_bases* bases, size_t skip, const void* scalars,
                          size_t n, void* out_jacobian);

/* ---- EC-FFT ---------------------------------------------------- */
/* radix_ec_fft (ag-cuda-ec/src/ec_fft.rs:13-99) / SingleEcFftKernel::radix_ec_fft
 * (ec-gpu-proxy/src/ec_fft.rs:53-160): in-place discrete Fourier transform of n = 2^log_n G1
 * points, out[k] = sum_j omega^(j k) in[j], natural order in and out -- what
 * Radix2EvaluationDomain::fft returns (ag-cuda-ec/src/ec_fft.rs:131-137).  Points are Jacobian
 * {x, y, z} Montgomery (Vec<Curve>), infinity <=> z == 0; omegas_mont[i] = omega^(2^i) in arkworks'
 * in-memory Fr layout (4 x u64 little-endian, Montgomery form), n_omegas >= log_n entries (the
 * reference passes 32, ag-cuda-ec/src/ec_fft.rs:120-124).  Passing the powers of omega^-1 gives
 * the unscaled inverse transform (ag-cuda-ec/benches/ec_fft.rs:88-106).  Synchronous.
 * Precondition on G1: the twiddle multiplications use the GLV endomorphism (k P = k1 P + k2 phi(P)), which equals
 * k P only for P in the prime-order subgroup.  BN254 G1 has cofactor 1 (every curve point qualifies); BLS12-381 G1 has
 * a cofactor, so inputs must be subgroup points -- what arkworks' checked deserialisation and every KZG / AMT setup
 * guarantee.  For on-curve points outside the subgroup set MSM_B200_ECFFT_GLV=0 (plain double-and-add, the
 * reference's POINT_mul semantics for any point, about 1.8x slower).  G2 never uses the split. */
int msm_ec_fft(msm_ctx* ctx, void* jacobian_inout, uint32_t log_n, const void* omegas_mont, uint32_t n_omegas);
/* Same with the point array in device memory of device 0 (omegas stay a host array). */
int msm_ec_fft_device(msm_ctx* ctx, void* d_jacobian_inout, uint32_t log_n, const void* omegas_mont,
                      uint32_t n_omegas);

/* ---- scalar-field FFT ---------------------------------------------- */
/* SingleFftKernel::radix_fft (ec-gpu-proxy/src/fft.rs:50-136; KERNEL FIELD_radix_fft,
 * ag-build/cl/fft.cl:4-66): in-place transform of n = 2^log_n elements of the scalar field Fr of the
 * context's curve, out[k] = sum_j omega^(j k) in[j], natural order in and out (serial_fft,
 * ec-gpu-proxy/src/fft_cpu.rs:10-52).  Elements and omega are in arkworks' in-memory Fr layout
 * (4 x u64 little-endian, Montgomery form, canonical).  Synchronous. */
int msm_scalar_fft(msm_ctx* ctx, void* fr_inout, uint32_t log_n, const void* omega_mont);
/* Same with the elements in device memory of device 0 (omega stays a host value). */
int msm_scalar_fft_device(msm_ctx* ctx, void* d_fr_inout, uint32_t log_n, const void* omega_mont);

/* ---- small device-side helpers used by callers, tests and the bench ----------------------- */
/* Sum `count` Jacobian points that live in device memory of device 0 into one Jacobian point
 * (device memory): the on-device replacement of the host loop acc.add_assign(&r)
 * (ec-gpu-proxy/src/multiexp.rs:394-397) after an NVLink / NCCL gather of per-GPU partials. */
int msm_sum_points_device(msm_ctx* ctx, const void* d_jacobian, size_t count, void* d_out_jacobian);
/* Jacobian -> affine {x,y} (Montgomery when mont_out != 0, else canonical) on the device, host
 * buffers in/out; infinity -> (0,0) and out_is_inf[i] = 1.  Curve::into_affine. */
int msm_to_affine(msm_ctx* ctx, const void* jacobian, size_t count, int mont_out, void* out_xy,
                  uint8_t* out_is_inf);
/* Deterministic synthetic inputs generated on the device (the counterpart of
 * ag-cuda-ec/src/test_tools.rs:4-15 random_input).  Outputs are device pointers on device 0.
 * points: P_i = (a + (start+i) b) G with 64-bit a, b derived from seed; scalars uniform < r. */
int msm_synth_points_device(msm_ctx* ctx, uint64_t seed, size_t start, size_t n, void* d_xy_mont);
int msm_synth_scalars_device(msm_ctx* ctx, uint64_t seed, size_t start, size_t n, void* d_scalars);
/* Per-primitive known-answer kernels, the counterpart of ag-build/cl/test.cl:1-35.  Host buffers.
 * fq op: 0 add, 1 sub, 2 mul, 3 sqr, 4 double, 5 to_mont, 6 from_mont, 7 inverse, 8 neg.
 * ec op: 0 add(Jac a, Jac b), 1 mixed add(Jac a, Aff b), 2 double(Jac a). */
int msm_test_fq_op(msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count);
int msm_test_ec_op(msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count);

/* Device memory helpers so that non-CUDA hosts (ctypes, Rust) can stage buffers on device 0. */
int msm_device_alloc(msm_ctx* ctx, size_t bytes, void** d_ptr);
int msm_device_free(msm_ctx* ctx, void* d_ptr);
int msm_memcpy_h2d(msm_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
int msm_memcpy_d2h(msm_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);
/* Pin / unpin a host range so that scalar uploads run at full PCIe rate. */
int msm_host_register(void* h_ptr, size_t bytes);
int msm_host_unregister(void* h_ptr);

const char* msm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MSM_B200_H */
