// engine_common.h -- host-side structures shared by engine.cu (C ABI) and the per-field
// instantiation units (inst_*.cu).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/msm_b200.h"

namespace msm {

// Grow-only device arena: replaces the per-call cuMemAlloc of the reference runtime
// (ag-cuda-proxy/src/params.rs:58-78, ag-cuda-ec/src/multiexp.rs:42-44).
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  cudaError_t ensure(size_t bytes) {
    off = 0;
    if (bytes <= cap) return cudaSuccess;
    if (base) cudaFree(base);
    base = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc((void**)&base, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc((void**)&base, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  template <class T> T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
  static size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }
  void release() {
    if (base) cudaFree(base);
    base = nullptr;
    cap = off = 0;
  }
};

struct DeviceCtx {
  int dev = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = true;
  Arena arena;        // per-call scratch (sorted entries, buckets, partials)
  Arena io;           // scalars / bases staged from the host + result points
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t launches = 0;
  void* small = nullptr;  // 64 KiB persistent (partial-point gather)
  cudaStream_t copy_stream = nullptr;  // host->device scalar chunks of pipelined calls
  cudaEvent_t ev_copy[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // Host -> device rate of the last scalar upload on this device (GB/s; 0 = none yet), from ev_h2d[0..1] around the
  // copies.  With the best device time of the shape (msm_bases::shape_best_ms) it sets how fast the sub-batches of
  // the next pipelined call may grow (multiple_multiexp_impl).
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
  float h2d_gbs = 0.f;
  // Sub-batch k+1 of a pipelined call is sorted on this (high-priority) stream while sub-batch k is still being
  // accumulated on `stream`: the sort lives on the load/store and shared-memory pipes, the bucket kernel on the
  // multiplier pipe.  ev_sorted[p] / ev_acc[p]: sort output set p is ready / no longer read; ev_fork orders the
  // sort stream behind whatever preceded the call on `stream`.
  cudaStream_t sort_stream = nullptr;
  cudaEvent_t ev_sorted[2] = {nullptr, nullptr}, ev_acc[2] = {nullptr, nullptr}, ev_fork = nullptr;
};

struct FieldOps;

}  // namespace msm

struct msm_ctx {
  int curve = 0;
  const msm::FieldOps* ops = nullptr;
  std::vector<msm::DeviceCtx> devs;
  std::atomic<int> in_use{0};
  const volatile int* abort_flag = nullptr;
  std::string err;
  msm_timings tm{};
  uint32_t window_override = 0;
  bool scalars_mont = false;  // set for the duration of a *_montgomery call: the scalar row is Fr in Montgomery form
  // 1 for the caller's handle + 1 per live msm_bases: msm_ctx_destroy only marks the context closed while
  // resident bases still point at it; the last msm_bases_free (or msm_ctx_destroy) tears it down
  std::atomic<int> refs{1};
  std::atomic<int> closed{0};
};

struct msm_bases {
  msm_ctx* ctx = nullptr;
  size_t n = 0;
  struct Shard {
    int dev_idx;
    void* ptr;  // resident copy in the field's packed layout
    size_t start, n;
    bool owned;
    void* table = nullptr;     // optional window table T[w][i] = 2^(c w) P_i (msm_bases_precompute)
    uint32_t table_c = 0, table_W = 0;
  };
  std::vector<Shard> shards;
  // Window-table policy (msm_bases_set_table_policy): MSM_TABLE_OFF never, MSM_TABLE_LAZY on the second
  // multiple_multiexp call of the same (L, num_chunks) shape, MSM_TABLE_EAGER at upload (whole-shard shape).
  int table_policy = 1;
  bool table_explicit = false;   // built by msm_bases_precompute[_chunked]: never rebuilt behind the caller's back
  bool table_failed = false;     // the lazy build for the current shape did not fit: stop trying
  size_t shape_L = 0, table_L = 0;
  uint32_t shape_chunks = 0, shape_calls = 0, table_chunks = 0;
  // shortest first-kernel-to-result time seen for the current shape (ms; 0 = none), and whether that was on the table
  float shape_best_ms = 0.f;
  bool shape_best_table = false;
};

namespace msm {

inline void set_error(msm_ctx* ctx, const std::string& s);
extern thread_local std::string g_create_error;
inline void set_error(msm_ctx* ctx, const std::string& s) {
  if (ctx) ctx->err = s;
  else g_create_error = s;
}
inline bool aborted(msm_ctx* ctx) { return ctx->abort_flag && *ctx->abort_flag; }
inline bool curve_is_bn254(int curve) { return curve == MSM_CURVE_BN254_G1 || curve == MSM_CURVE_BN254_G2; }
inline uint32_t scalar_bits(int curve) { return curve_is_bn254(curve) ? 254 : 255; }

#define CU_TRY(ctx, call)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::msm::set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(_e));            \
      return MSM_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

// Per-field entry points (one instantiation unit each, see engine_impl.cuh).
// How the host scalars of one call are pipelined: number of sub-batches and the factor their sizes grow by.
//   * One MSM (num_chunks == 1, one line of bases) from 2^20 scalars: parts of the row that continue one bucket array.
//     Nothing measured yet (first call of a shape): 2 parts, 4 from 2^23 scalars, sizes doubling -- four doubling parts
//     expose the same first upload as fifteen equal ones.  From the second call of a shape on, both speeds are known
//     (h2d_gbs: upload rate of the last call on this device, device_ms: shortest device time of the shape): the parts
//     grow by 0.85 x device time / upload time, clamped to 1.5 .. 3.  Every extra part costs ~2.3 % of the call (its
//     buckets are merged into the running ones, more slices are cut), so when the link feeds 3-fold growth, three parts
//     (1/13, 3/13, 9/13) beat four growing 2-fold (1/15 ... 8/15): 2^24 scalars on one B200 35.55 against 36.74 ms,
//     2^23 19.20 against 19.66 (job r2_run23); a slower link keeps the growth below what it can feed.
//   * Many independent tasks in one row (the reference's bench geometry, 1024 x 2^12; one line, >= 16 tasks, from 2^20
//     scalars): 3 groups of whole tasks, sizes doubling.  Every group is sorted and accumulated into its own range of
//     the bucket array as soon as it has landed, and ONE reduction + combine runs over all tasks at the end (a
//     reduction per group was measured first: each is a latency-bound chain of ~50 dependent additions, 4 groups
//     18.8 ms = no gain, job r2_run30).  1024 x 2^12 end to end: 18.77 ms in one piece, 17.62 / 17.49 / 17.69 / 18.23 /
//     18.78 in 2 / 3 / 4 / 6 / 8 groups (each group still costs ~0.35 ms in short sorts and partial waves).
//   * Everything else (several lines of bases, short rows): one piece.
inline void pipeline_shape(size_t L, uint32_t num_chunks, uint32_t n_lines, float h2d_gbs, float device_ms,
                           uint32_t* n_sub, double* growth) {
  *n_sub = 1;
  *growth = 2.0;
  if (n_lines != 1 || L < (1u << 20)) return;
  if (num_chunks == 1) {
    *n_sub = L >= (1u << 23) ? 4 : 2;
    if (h2d_gbs > 0.f && device_ms > 0.f) {
      const double copy_ms = (double)L * 32.0 / ((double)h2d_gbs * 1e6);
      const double r = 0.85 * (double)device_ms / copy_ms;
      *growth = r < 1.5 ? 1.5 : (r > 3.0 ? 3.0 : r);
      if (L >= (1u << 23)) *n_sub = *growth >= 2.5 ? 3 : 4;
    }
  } else if (num_chunks >= 16) {
    *n_sub = 3;
  }
}

struct FieldOps {
  const char* name;
  size_t api_point_bytes;     // {x,y} at the API boundary
  size_t packed_point_bytes;  // resident copy
  int (*multiple_multiexp)(msm_ctx*, const msm_bases*, const void* scalars, size_t L, uint32_t num_chunks,
                           void* out, bool device_io);
  int (*multiexp)(msm_ctx*, const void* host_bases, const msm_bases* resident, size_t skip, const void* scalars,
                  size_t n, void* out);
  // d_api (device, API layout) -> d_packed (device, resident layout); enqueued on dc.stream
  int (*convert_bases)(msm_ctx*, DeviceCtx&, const void* d_api, size_t n, void* d_packed);
  // window table for one shard (allocates sh.table); c == 0: the engine's choice for shard-sized MSMs
  // (chunk_len == 0) or for tasks of chunk_len points
  int (*build_table)(msm_ctx*, msm_bases::Shard& sh, uint32_t c, size_t chunk_len);
  int (*synth_points)(msm_ctx*, uint64_t seed, size_t start, size_t n, void* d_out);
  int (*test_fq)(msm_ctx*, int op, const void* a, const void* b, void* out, size_t count);
  int (*test_ec)(msm_ctx*, int op, const void* a, const void* b, void* out, size_t count);
  int (*to_affine)(msm_ctx*, const void* jac, size_t count, int mont_out, void* out_xy, uint8_t* out_inf);
  int (*sum_points)(msm_ctx*, const void* d_in, size_t count, void* d_out);
  int (*ec_fft)(msm_ctx*, void* jac, uint32_t log_n, const void* omegas_mont, uint32_t n_omegas, bool device_io);
  // make_plan without a call behind it (msm_plan_describe): host arithmetic only
  int (*describe_plan)(msm_ctx*, uint32_t L, uint32_t n_lines, uint32_t num_chunks, uint32_t table_c, uint32_t n_sub,
                       double growth, msm_plan_info* out);
};
const FieldOps* field_ops_bn254_u29();
const FieldOps* field_ops_bn254_sat();
const FieldOps* field_ops_bls381_sat();
const FieldOps* field_ops_bn254_lazy();
const FieldOps* field_ops_bls381_lazy();
const FieldOps* field_ops_bn254_g2();
const FieldOps* field_ops_bls381_g2();
void fill_field_ops_aux_bn254_g2(FieldOps& o);   // inst_bn254_g2_aux.cu
void fill_field_ops_aux_bls381_g2(FieldOps& o);  // inst_bls381_g2_aux.cu

}  // namespace msm
