// engine.cu -- the extern "C" surface declared in include/msm_b200.h: contexts, resident bases,
// dispatch to the per-field engines (inst_*.cu through FieldOps).
//
// Replaces, for the MSM path only: ag-cuda-proxy (CudaWorkspace / Kernel / Param / DeviceData,
// ag-cuda-proxy/src/{module,context,kernel,params}.rs) and the dispatcher of
// ec_gpu_proxy::MultiexpKernel (ec-gpu-proxy/src/multiexp.rs:256-403).
// There is no CPU fallback: without a CUDA device every entry point fails with MSM_ERR_NO_DEVICE.
#include "engine_common.h"
#include "kernels.cuh"
#include "ntt.cuh"

using namespace msm;

namespace msm {
thread_local std::string g_create_error;
}

namespace {

struct CtxLock {
  msm_ctx* ctx;
  bool ok;
  explicit CtxLock(msm_ctx* c) : ctx(c) {
    int expected = 0;
    ok = c->in_use.compare_exchange_strong(expected, 1);
  }
  ~CtxLock() {
    if (ok) ctx->in_use.store(0);
  }
};

ScalarField scalar_field(int curve) {
  ScalarField f;
  if (curve_is_bn254(curve)) {
    const uint32_t r[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                           0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    memcpy(f.r, r, sizeof(r));
    f.bits = 254;
  } else {
    const uint32_t r[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                           0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
    memcpy(f.r, r, sizeof(r));
    f.bits = 255;
  }
  return f;
}

const FieldOps* pick_ops(int curve) {
  // Default: saturated 32-bit limbs with values kept in [0, 2p) ("sat32-lazy").  MSM_B200_FIELD
  // selects the alternatives that were built and measured:
  //   sat32  canonical values, conditional subtraction after every product
  //   u29    BN254 only: nine 29-bit limbs, carry-free IMAD.WIDE (slower on B200)
  if (curve == MSM_CURVE_BN254_G2) return field_ops_bn254_g2();
  if (curve == MSM_CURVE_BLS12_381_G2) return field_ops_bls381_g2();
#ifdef MSM_EXPERIMENTAL  // make MSM_EXPERIMENTAL=1: the measured-slower variants are not in the product build
  const char* env = getenv("MSM_B200_FIELD");
  const bool want_sat = env && (strcmp(env, "sat32") == 0 || strcmp(env, "sat") == 0);
  if (curve == MSM_CURVE_BLS12_381_G1 && want_sat) return field_ops_bls381_sat();
  if (curve == MSM_CURVE_BN254_G1 && env && strcmp(env, "u29") == 0) return field_ops_bn254_u29();
  if (curve == MSM_CURVE_BN254_G1 && want_sat) return field_ops_bn254_sat();
#endif
  return curve == MSM_CURVE_BLS12_381_G1 ? field_ops_bls381_lazy() : field_ops_bn254_lazy();
}

#define LOCK_OR_BUSY(ctx)                                   \
  if (!(ctx) || (ctx)->closed.load()) return MSM_ERR_INVALID; \
  CtxLock _lock(ctx);                                       \
  if (!_lock.ok) return MSM_ERR_BUSY

int default_table_policy() {
  const char* env = getenv("MSM_B200_TABLE");
  if (!env) return MSM_TABLE_LAZY;
  if (strcmp(env, "off") == 0 || strcmp(env, "0") == 0) return MSM_TABLE_OFF;
  if (strcmp(env, "eager") == 0) return MSM_TABLE_EAGER;
  return MSM_TABLE_LAZY;
}

// frees the device side of a context; called when the last reference (handle or msm_bases) goes
void ctx_teardown(msm_ctx* ctx) {
  for (auto& dc : ctx->devs) {
    cudaSetDevice(dc.dev);
    if (dc.stream) cudaStreamSynchronize(dc.stream);
    dc.arena.release();
    dc.io.release();
    if (dc.small) cudaFree(dc.small);
    if (dc.copy_stream) cudaStreamDestroy(dc.copy_stream);
    for (auto& e : dc.ev_copy)
      if (e) cudaEventDestroy(e);
    for (auto& e : dc.ev_h2d)
      if (e) cudaEventDestroy(e);
    for (auto& e : dc.ev_sorted)
      if (e) cudaEventDestroy(e);
    for (auto& e : dc.ev_acc)
      if (e) cudaEventDestroy(e);
    if (dc.ev_fork) cudaEventDestroy(dc.ev_fork);
    if (dc.sort_stream) cudaStreamDestroy(dc.sort_stream);
    for (auto& e : dc.ev)
      if (e) cudaEventDestroy(e);
    if (dc.stream && dc.owns_stream) cudaStreamDestroy(dc.stream);
  }
  delete ctx;
}
void ctx_unref(msm_ctx* ctx) {
  if (ctx->refs.fetch_sub(1) == 1) ctx_teardown(ctx);
}

}  // namespace

// radix_fft (ec-gpu-proxy/src/fft.rs:50-136) over the scalar field of the context's curve.
template <class PR>
int scalar_fft_impl(msm_ctx* ctx, void* data, uint32_t log_n, const void* omega_mont, bool device_io) {
  if (log_n > 27) {
    set_error(ctx, "scalar fft: at most 2^27 elements");
    return MSM_ERR_TOO_LARGE;
  }
  if (log_n == 0) return MSM_OK;
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const size_t n = (size_t)1 << log_n, bytes = n * 32;
  const uint32_t half = (uint32_t)(n / 2), n_hi = half > 1024 ? half >> 10 : 0;
  if (!device_io) CU_TRY(ctx, dc.io.ensure(Arena::padded(bytes)));
  uint32_t* d_x = device_io ? static_cast<uint32_t*>(data) : dc.io.take<uint32_t>(n * 8);
  CU_TRY(ctx, dc.arena.ensure(Arena::padded(bytes) + Arena::padded((size_t)half * 32) + Arena::padded((size_t)(n_hi + 1) * 32)));
  uint32_t* d_y = dc.arena.take<uint32_t>(n * 8);
  uint32_t* d_tw = dc.arena.take<uint32_t>((size_t)half * 8);
  uint32_t* d_hi = dc.arena.take<uint32_t>((size_t)(n_hi + 1) * 8);
  Fp<PR> omega;
  memcpy(omega.v, omega_mont, 32);
  cudaStream_t st = dc.stream;
  CU_TRY(ctx, cudaEventRecord(dc.ev[0], st));
  if (!device_io) CU_TRY(ctx, cudaMemcpyAsync(d_x, data, bytes, cudaMemcpyHostToDevice, st));
  CU_TRY(ctx, cudaEventRecord(dc.ev[1], st));
  // twiddles: tw[j] = omega^j, j < n/2 -- the first 1024 by the binary method, the rest with one product each
  const uint32_t low = half < 1024 ? half : 1024;
  k_ntt_pow_table<PR><<<(low + 127) / 128, 128, 0, st>>>(omega, 0u, low, d_tw);
  dc.launches += 1;
  if (n_hi) {
    k_ntt_pow_table<PR><<<(n_hi + 127) / 128, 128, 0, st>>>(omega, 10u, n_hi, d_hi);
    k_ntt_expand_table<PR><<<(half - 1024 + 255) / 256, 256, 0, st>>>(d_hi, half, d_tw);
    dc.launches += 2;
  }
  const uint32_t R = log_n < 10 ? log_n : 10;
  CU_TRY(ctx, cudaFuncSetAttribute(k_ntt_first<PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 << 10));
  CU_TRY(ctx, cudaFuncSetAttribute(k_ntt_pass<PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 << 10));
  k_ntt_first<PR><<<(uint32_t)(n >> R), NTT_BLOCK, (size_t)48 << R, st>>>(d_x, d_y, log_n, R, d_tw);  // 2^R elements + 2^(R-1) twiddles
  dc.launches += 1;
  for (uint32_t s0 = R; s0 < log_n;) {
    if (aborted(ctx)) {  // SingleFftKernel polls maybe_abort once per pass (ec-gpu-proxy/src/fft.rs:86-90)
      cudaStreamSynchronize(st);
      return MSM_ERR_ABORTED;
    }
    const uint32_t r = log_n - s0 < 5 ? log_n - s0 : 5;
    k_ntt_pass<PR><<<(uint32_t)(n >> (5 + r)), NTT_BLOCK, ((size_t)1024 << r) + (((size_t)1 << r) - 1) * 1024, st>>>(d_y, log_n, s0, r, d_tw);  // tile + its twiddles
    dc.launches += 1;
    s0 += r;
  }
  CU_TRY(ctx, cudaEventRecord(dc.ev[4], st));
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaMemcpyAsync(data, d_y, bytes, device_io ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, cudaStreamSynchronize(st));
  msm_timings& t = ctx->tm;
  memset(&t, 0, sizeof(t));
  cudaEventElapsedTime(&t.h2d_ms, dc.ev[0], dc.ev[1]);
  cudaEventElapsedTime(&t.total_ms, dc.ev[1], dc.ev[4]);
  t.kernel_launches = dc.launches;
  return MSM_OK;
}

// =================================================================================================
extern "C" {

const char* msm_version(void) { return "msm_b200 0.3 (sm_100a; BN254 / BLS12-381 G1 and G2, 32-bit-limb Montgomery fields)"; }

int msm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int msm_ctx_create(int curve, const int* device_ids, int n_devices, msm_ctx** out) {
  if (!out || curve < MSM_CURVE_BN254_G1 || curve > MSM_CURVE_BLS12_381_G2 || n_devices < 0) {
    set_error(nullptr, "msm_ctx_create: invalid argument");
    return MSM_ERR_INVALID;
  }
  const int visible = msm_device_count();
  if (visible <= 0) {
    set_error(nullptr, "No working GPUs found!");
    return MSM_ERR_NO_DEVICE;
  }
  if (n_devices == 0) n_devices = visible;
  msm_ctx* ctx = new msm_ctx();
  ctx->curve = curve;
  ctx->ops = pick_ops(curve);
  for (int i = 0; i < n_devices; i++) {
    const int dev = device_ids ? device_ids[i] : i;
    if (dev < 0 || dev >= visible) {
      set_error(nullptr, "msm_ctx_create: device id out of range");
      msm_ctx_destroy(ctx);
      return MSM_ERR_INVALID;
    }
    DeviceCtx dc;
    dc.dev = dev;
    cudaError_t e = cudaSetDevice(dev);
    if (const char* env = getenv("MSM_B200_L2_FETCH")) {
      // experiment: L2 fetch granularity for the 64-byte point gathers (32 / 64 / 128 bytes)
      cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(env));
      cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&dc.stream, cudaStreamNonBlocking);
    for (int k = 0; k < 5 && e == cudaSuccess; k++) e = cudaEventCreate(&dc.ev[k]);
    if (e == cudaSuccess) e = cudaMalloc(&dc.small, 65536);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&dc.copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < 8 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&dc.ev_copy[k], cudaEventDisableTiming);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaEventCreate(&dc.ev_h2d[k]);
    if (e == cudaSuccess) {
      int prio_lo = 0, prio_hi = 0;
      cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);  // numerically lowest = highest priority
      e = cudaStreamCreateWithPriority(&dc.sort_stream, cudaStreamNonBlocking, prio_hi);
    }
    for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&dc.ev_sorted[k], cudaEventDisableTiming);
    for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(&dc.ev_acc[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&dc.ev_fork, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      // a device whose kernel cannot be initialised is skipped (ec-gpu-proxy/src/multiexp.rs:288-303)
      set_error(nullptr, std::string("device init failed: ") + cudaGetErrorString(e));
      cudaGetLastError();
      continue;
    }
    ctx->devs.push_back(dc);
  }
  if (ctx->devs.empty()) {
    delete ctx;
    set_error(nullptr, "No working GPUs found!");
    return MSM_ERR_NO_DEVICE;
  }
  // peer access towards device 0 for the partial-point gather
  for (size_t i = 1; i < ctx->devs.size(); i++) {
    int can = 0;
    cudaDeviceCanAccessPeer(&can, ctx->devs[0].dev, ctx->devs[i].dev);
    if (can) {
      cudaSetDevice(ctx->devs[0].dev);
      cudaDeviceEnablePeerAccess(ctx->devs[i].dev, 0);
      cudaSetDevice(ctx->devs[i].dev);
      cudaDeviceEnablePeerAccess(ctx->devs[0].dev, 0);
      cudaGetLastError();  // already-enabled is fine
    }
  }
  cudaSetDevice(ctx->devs[0].dev);
  *out = ctx;
  return MSM_OK;
}

int msm_ctx_destroy(msm_ctx* ctx) {
  if (!ctx || ctx->closed.exchange(1)) return MSM_ERR_INVALID;
  // scratch goes now; streams and the context object stay while resident bases still refer to them
  if (ctx->refs.load() > 1)
    for (auto& dc : ctx->devs) {
      cudaSetDevice(dc.dev);
      if (dc.stream) cudaStreamSynchronize(dc.stream);
      dc.arena.release();
      dc.io.release();
    }
  ctx_unref(ctx);
  return MSM_OK;
}

int msm_ctx_num_devices(const msm_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int msm_set_abort_flag(msm_ctx* ctx, const volatile int* flag) {
  if (!ctx) return MSM_ERR_INVALID;
  ctx->abort_flag = flag;
  return MSM_OK;
}

const char* msm_last_error(const msm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int msm_last_timings(const msm_ctx* ctx, msm_timings* out) {
  if (!ctx || !out) return MSM_ERR_INVALID;
  *out = ctx->tm;
  return MSM_OK;
}

int msm_plan_describe(int curve, size_t L, uint32_t n_lines, uint32_t num_chunks, uint32_t table_window_bits,
                      uint32_t sub_batches, double growth, msm_plan_info* out) {
  if (!out || curve < MSM_CURVE_BN254_G1 || curve > MSM_CURVE_BLS12_381_G2 || L == 0 || L >= (1ull << 31) ||
      sub_batches > 8 || table_window_bits > 24)
    return MSM_ERR_INVALID;
  msm_ctx shadow;  // no devices: make_plan only reads the curve, the window override and the scalar form
  shadow.curve = curve;
  shadow.ops = pick_ops(curve);
  return shadow.ops->describe_plan(&shadow, (uint32_t)L, n_lines, num_chunks, table_window_bits, sub_batches, growth, out);
}

int msm_pipeline_shape(size_t L, uint32_t n_lines, uint32_t num_chunks, float h2d_gbs, float device_ms,
                       uint32_t* sub_batches, double* growth) {
  if (!sub_batches || !growth || L == 0 || n_lines == 0 || num_chunks == 0) return MSM_ERR_INVALID;
  pipeline_shape(L, num_chunks, n_lines, h2d_gbs, device_ms, sub_batches, growth);
  return MSM_OK;
}

int msm_set_window_bits(msm_ctx* ctx, uint32_t c) {
  if (!ctx || (c != 0 && (c < 2 || c > 24))) return MSM_ERR_INVALID;
  ctx->window_override = c;
  return MSM_OK;
}

// Copies (host source) or reads (device source) API-layout points and leaves the engine's resident
// copy -- converted to the field's packed layout -- on the device(s).
static int make_resident(msm_ctx* ctx, const void* xy, size_t n, bool sharded, bool src_on_device, msm_bases** out) {
  if (!ctx || !out || (!xy && n)) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  const FieldOps* ops = ctx->ops;
  msm_bases* b = new msm_bases();
  b->ctx = ctx;
  b->n = n;
  b->table_policy = default_table_policy();
  const size_t n_dev = sharded ? ctx->devs.size() : 1;
  const size_t chunk = n_dev ? (n + n_dev - 1) / n_dev : 0;
  auto fail = [&](const std::string& what) {
    set_error(ctx, what);
    for (auto& s : b->shards)
      if (s.owned && s.ptr) {
        cudaSetDevice(ctx->devs[s.dev_idx].dev);
        cudaFree(s.ptr);
      }
    delete b;
    return MSM_ERR_CUDA;
  };
  for (size_t d = 0; d < n_dev; d++) {
    const size_t start = d * chunk;
    if (start >= n && d != 0) break;
    const size_t cnt = start < n ? std::min(chunk, n - start) : 0;
    msm_bases::Shard sh{(int)d, nullptr, start, cnt, true};
    if (cnt) {
      DeviceCtx& dc = ctx->devs[d];
      cudaError_t e = cudaSetDevice(dc.dev);
      if (e == cudaSuccess) e = cudaMalloc(&sh.ptr, cnt * ops->packed_point_bytes);
      if (e != cudaSuccess) return fail(std::string("msm_bases_upload: ") + cudaGetErrorString(e));
      b->shards.push_back(sh);
      const char* src = static_cast<const char*>(xy) + start * ops->api_point_bytes;
      const void* d_api = src;
      if (!src_on_device || d != 0) {
        e = dc.io.ensure(Arena::padded(cnt * ops->api_point_bytes));
        if (e != cudaSuccess) return fail(std::string("msm_bases_upload staging: ") + cudaGetErrorString(e));
        void* stage = dc.io.take<char>(cnt * ops->api_point_bytes);
        e = cudaMemcpyAsync(stage, src, cnt * ops->api_point_bytes, cudaMemcpyDefault, dc.stream);
        if (e != cudaSuccess) return fail(std::string("msm_bases_upload copy: ") + cudaGetErrorString(e));
        d_api = stage;
      }
      int rc = ops->convert_bases(ctx, dc, d_api, cnt, sh.ptr);
      // the reference returns without synchronising (ag-cuda-proxy/src/params.rs:186-201); the
      // source slice may be freed by the caller right after, so this implementation waits.
      if (rc == MSM_OK && cudaStreamSynchronize(dc.stream) != cudaSuccess) rc = MSM_ERR_CUDA;
      if (rc != MSM_OK) return fail("msm_bases_upload: conversion failed: " + ctx->err);
    } else {
      b->shards.push_back(sh);
    }
  }
  ctx->refs.fetch_add(1);
  if (b->table_policy == MSM_TABLE_EAGER && !sharded) {
    // best effort: a table that does not fit leaves the plain resident copy in charge
    for (auto& sh : b->shards)
      if (ops->build_table(ctx, sh, 0, 0) == MSM_OK) {
        b->table_L = sh.n;
        b->table_chunks = 1;
      }
  }
  *out = b;
  return MSM_OK;
}

int msm_bases_upload(msm_ctx* ctx, const void* xy, size_t n, msm_bases** out) {
  return make_resident(ctx, xy, n, false, false, out);
}
int msm_bases_upload_sharded(msm_ctx* ctx, const void* xy, size_t n, msm_bases** out) {
  return make_resident(ctx, xy, n, true, false, out);
}
int msm_bases_from_device(msm_ctx* ctx, const void* d_xy, size_t n, msm_bases** out) {
  return make_resident(ctx, d_xy, n, false, true, out);
}
static int precompute_tables(msm_ctx* ctx, msm_bases* b, uint32_t window_bits, size_t chunk_len) {
  if (!ctx || !b || b->ctx != ctx) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  for (auto& sh : b->shards) {
    int rc = ctx->ops->build_table(ctx, sh, window_bits, chunk_len);
    if (rc != MSM_OK) return rc;
  }
  b->table_explicit = true;
  return MSM_OK;
}
int msm_bases_set_table_policy(msm_ctx* ctx, msm_bases* b, int policy) {
  if (!ctx || !b || b->ctx != ctx || policy < MSM_TABLE_OFF || policy > MSM_TABLE_EAGER) return MSM_ERR_INVALID;
  b->table_policy = policy;
  if (policy == MSM_TABLE_EAGER && !msm_bases_table_window(b)) {
    int rc = precompute_tables(ctx, b, 0, 0);
    b->table_explicit = false;
    if (rc == MSM_OK && !b->shards.empty()) {
      b->table_L = b->shards[0].n;
      b->table_chunks = 1;
    }
    return rc;
  }
  return MSM_OK;
}
int msm_bases_precompute(msm_ctx* ctx, msm_bases* b, uint32_t window_bits) {
  return precompute_tables(ctx, b, window_bits, 0);
}
int msm_bases_precompute_chunked(msm_ctx* ctx, msm_bases* b, size_t chunk_len) {
  if (chunk_len == 0) return MSM_ERR_INVALID;
  return precompute_tables(ctx, b, 0, chunk_len);
}
uint32_t msm_bases_table_window(const msm_bases* b) { return (b && !b->shards.empty()) ? b->shards[0].table_c : 0; }
size_t msm_bases_size_bytes(const msm_bases* b) { return b ? b->n * b->ctx->ops->api_point_bytes : 0; }
size_t msm_bases_num_points(const msm_bases* b) { return b ? b->n : 0; }
int msm_bases_free(msm_bases* b) {
  if (!b) return MSM_ERR_INVALID;
  for (auto& s : b->shards) {
    if ((s.owned && s.ptr) || s.table) cudaSetDevice(b->ctx->devs[s.dev_idx].dev);
    if (s.owned && s.ptr) cudaFree(s.ptr);
    if (s.table) cudaFree(s.table);
  }
  msm_ctx* ctx = b->ctx;
  delete b;
  ctx_unref(ctx);
  return MSM_OK;
}

int msm_multiple_multiexp(msm_ctx* ctx, const msm_bases* bases, const void* scalars, size_t L,
                          uint32_t num_chunks, uint32_t window_hint, int neg_is_cheap, void* out) {
  (void)window_hint;
  (void)neg_is_cheap;
  if (!ctx || !bases || !scalars || !out || bases->ctx != ctx) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  return ctx->ops->multiple_multiexp(ctx, bases, scalars, L, num_chunks, out, false);
}

int msm_multiple_multiexp_device(msm_ctx* ctx, const msm_bases* bases, const void* d_scalars, size_t L,
                                 uint32_t num_chunks, void* d_out) {
  if (!ctx || !bases || !d_scalars || !d_out || bases->ctx != ctx) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  return ctx->ops->multiple_multiexp(ctx, bases, d_scalars, L, num_chunks, d_out, true);
}

int msm_multiexp(msm_ctx* ctx, const void* bases_xy, const void* scalars, size_t n, void* out) {
  if (!ctx || !out || (n && (!bases_xy || !scalars))) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  return ctx->ops->multiexp(ctx, bases_xy, nullptr, 0, scalars, n, out);
}

int msm_multiexp_resident(msm_ctx* ctx, const msm_bases* bases, size_t skip, const void* scalars, size_t n,
                          void* out) {
  if (!ctx || !bases || !out || (n && !scalars) || bases->ctx != ctx) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  return ctx->ops->multiexp(ctx, nullptr, bases, skip, scalars, n, out);
}

int msm_sum_points_device(msm_ctx* ctx, const void* d_jac, size_t count, void* d_out) {
  if (!ctx || !d_jac || !d_out) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  return ctx->ops->sum_points(ctx, d_jac, count, d_out);
}

int msm_to_affine(msm_ctx* ctx, const void* jac, size_t count, int mont_out, void* out_xy, uint8_t* out_inf) {
  if (!ctx || !jac || !out_xy) return MSM_ERR_INVALID;
  if (count == 0) return MSM_OK;
  LOCK_OR_BUSY(ctx);
  return ctx->ops->to_affine(ctx, jac, count, mont_out, out_xy, out_inf);
}

int msm_synth_points_device(msm_ctx* ctx, uint64_t seed, size_t start, size_t n, void* d_xy) {
  if (!ctx || (!d_xy && n)) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  return ctx->ops->synth_points(ctx, seed, start, n, d_xy);
}

int msm_synth_scalars_device(msm_ctx* ctx, uint64_t seed, size_t start, size_t n, void* d_scalars) {
  if (!ctx || (!d_scalars && n)) return MSM_ERR_INVALID;
  if (n == 0) return MSM_OK;
  if (n >= (1ull << 32)) return MSM_ERR_TOO_LARGE;
  LOCK_OR_BUSY(ctx);
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  k_synth_scalars<<<(uint32_t)((n + 255) / 256), 256, 0, dc.stream>>>(scalar_field(ctx->curve), seed, (uint64_t)start,
                                                                      (uint32_t)n, static_cast<uint32_t*>(d_scalars));
  dc.launches += 1;
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

int msm_scalars_from_montgomery_device(msm_ctx* ctx, const void* d_in, size_t n, void* d_out) {
  if (!ctx || (n && (!d_in || !d_out))) return MSM_ERR_INVALID;
  if (n == 0) return MSM_OK;
  if (n >= (1ull << 32)) return MSM_ERR_TOO_LARGE;
  LOCK_OR_BUSY(ctx);
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const uint32_t grid = (uint32_t)((n + 255) / 256);
  if (curve_is_bn254(ctx->curve))
    k_scalars_unmont<Bn254Fr><<<grid, 256, 0, dc.stream>>>(static_cast<const uint32_t*>(d_in), (uint32_t)n, static_cast<uint32_t*>(d_out));
  else
    k_scalars_unmont<Bls381Fr><<<grid, 256, 0, dc.stream>>>(static_cast<const uint32_t*>(d_in), (uint32_t)n, static_cast<uint32_t*>(d_out));
  dc.launches += 1;
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

int msm_multiple_multiexp_montgomery(msm_ctx* ctx, const msm_bases* bases, const void* scalars_mont, size_t L,
                                     uint32_t num_chunks, void* out) {
  if (!ctx || !bases || !scalars_mont || !out || bases->ctx != ctx || L == 0 || L >= (1ull << 31)) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  // the conversion is fused into the digit decomposition (sort.cu: load_scalar_geo): same call path as
  // msm_multiple_multiexp, pipelined upload included
  ctx->scalars_mont = true;
  const int rc = ctx->ops->multiple_multiexp(ctx, bases, scalars_mont, L, num_chunks, out, false);
  ctx->scalars_mont = false;
  return rc;
}

int msm_ec_fft(msm_ctx* ctx, void* jacobian_inout, uint32_t log_n, const void* omegas_mont, uint32_t n_omegas) {
  if (!ctx || !omegas_mont || (!jacobian_inout && log_n)) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  return ctx->ops->ec_fft(ctx, jacobian_inout, log_n, omegas_mont, n_omegas, false);
}
int msm_ec_fft_device(msm_ctx* ctx, void* d_jacobian_inout, uint32_t log_n, const void* omegas_mont, uint32_t n_omegas) {
  if (!ctx || !omegas_mont || (!d_jacobian_inout && log_n)) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  return ctx->ops->ec_fft(ctx, d_jacobian_inout, log_n, omegas_mont, n_omegas, true);
}

int msm_scalar_fft(msm_ctx* ctx, void* fr_inout, uint32_t log_n, const void* omega_mont) {
  if (!ctx || !omega_mont || (!fr_inout && log_n)) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  return curve_is_bn254(ctx->curve) ? scalar_fft_impl<Bn254Fr>(ctx, fr_inout, log_n, omega_mont, false)
                                          : scalar_fft_impl<Bls381Fr>(ctx, fr_inout, log_n, omega_mont, false);
}
int msm_scalar_fft_device(msm_ctx* ctx, void* d_fr_inout, uint32_t log_n, const void* omega_mont) {
  if (!ctx || !omega_mont || (!d_fr_inout && log_n)) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  return curve_is_bn254(ctx->curve) ? scalar_fft_impl<Bn254Fr>(ctx, d_fr_inout, log_n, omega_mont, true)
                                          : scalar_fft_impl<Bls381Fr>(ctx, d_fr_inout, log_n, omega_mont, true);
}

int msm_test_fq_op(msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
  if (!ctx || !a || !out || op < 0 || op > 8) return MSM_ERR_INVALID;
  if (count == 0) return MSM_OK;
  LOCK_OR_BUSY(ctx);
  return ctx->ops->test_fq(ctx, op, a, b, out, count);
}

int msm_test_ec_op(msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
  if (!ctx || !a || !out || op < 0 || op > 2 || (op < 2 && !b)) return MSM_ERR_INVALID;
  if (count == 0) return MSM_OK;
  LOCK_OR_BUSY(ctx);
  return ctx->ops->test_ec(ctx, op, a, b, out, count);
}

const char* msm_field_impl(const msm_ctx* ctx) { return ctx ? ctx->ops->name : ""; }

int msm_set_stream(msm_ctx* ctx, void* cuda_stream) {
  if (!ctx) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  if (dc.owns_stream && dc.stream) cudaStreamDestroy(dc.stream);
  if (cuda_stream) {
    dc.stream = static_cast<cudaStream_t>(cuda_stream);
    dc.owns_stream = false;
  } else {
    CU_TRY(ctx, cudaStreamCreateWithFlags(&dc.stream, cudaStreamNonBlocking));
    dc.owns_stream = true;
  }
  return MSM_OK;
}

int msm_multiple_multiexp_device_timed(msm_ctx* ctx, const msm_bases* bases, const void* d_scalars, size_t L,
                                       uint32_t num_chunks, void* d_out, uint32_t repeats, float* total_ms,
                                       float* accumulate_ms) {
  if (!ctx || !bases || !d_scalars || !d_out || bases->ctx != ctx || repeats == 0) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  cudaEvent_t e0, e1;
  CU_TRY(ctx, cudaEventCreate(&e0));
  CU_TRY(ctx, cudaEventCreate(&e1));
  CU_TRY(ctx, cudaEventRecord(e0, dc.stream));
  float acc = 0.f;
  int rc = MSM_OK;
  for (uint32_t r = 0; r < repeats && rc == MSM_OK; r++) {
    // the per-call synchronise only waits for the tiny result; the GPU timeline between e0 and e1
    // is what is reported
    rc = ctx->ops->multiple_multiexp(ctx, bases, d_scalars, L, num_chunks, d_out, true);
    acc += ctx->tm.accumulate_ms;
  }
  cudaEventRecord(e1, dc.stream);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (total_ms) *total_ms = ms;
  if (accumulate_ms) *accumulate_ms = acc;
  return rc;
}

int msm_device_alloc(msm_ctx* ctx, size_t bytes, void** d_ptr) {
  if (!ctx || !d_ptr) return MSM_ERR_INVALID;
  CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
  CU_TRY(ctx, cudaMalloc(d_ptr, bytes ? bytes : 1));
  return MSM_OK;
}
int msm_device_free(msm_ctx* ctx, void* d_ptr) {
  if (!ctx) return MSM_ERR_INVALID;
  CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
  CU_TRY(ctx, cudaFree(d_ptr));
  return MSM_OK;
}
int msm_memcpy_h2d(msm_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  if (!ctx) return MSM_ERR_INVALID;
  CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
  CU_TRY(ctx, cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
  return MSM_OK;
}
int msm_memcpy_d2h(msm_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  if (!ctx) return MSM_ERR_INVALID;
  CU_TRY(ctx, cudaSetDevice(ctx->devs[0].dev));
  CU_TRY(ctx, cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
  return MSM_OK;
}
int msm_host_register(void* h_ptr, size_t bytes) {
  return cudaHostRegister(h_ptr, bytes, cudaHostRegisterDefault) == cudaSuccess ? MSM_OK : MSM_ERR_CUDA;
}
int msm_host_unregister(void* h_ptr) {
  return cudaHostUnregister(h_ptr) == cudaSuccess ? MSM_OK : MSM_ERR_CUDA;
}

}  // extern "C"
