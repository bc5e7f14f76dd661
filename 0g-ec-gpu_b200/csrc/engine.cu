// engine.cu -- host side of libmsm_b200.so: contexts, workspace arena, launch sequencing, the
// multi-GPU split and the extern "C" surface declared in include/msm_b200.h.
//
// Replaces, for the MSM path only: ag-cuda-proxy (CudaWorkspace / Kernel / Param / DeviceData,
// ag-cuda-proxy/src/{module,context,kernel,params}.rs), the host wrapper
// ag_cuda_ec::multiple_multiexp (ag-cuda-ec/src/multiexp.rs:22-81) and the legacy dispatcher
// ec_gpu_proxy::MultiexpKernel (ec-gpu-proxy/src/multiexp.rs:256-403).
//
// Differences from the reference runtime that are the point of the rewrite:
//   * no per-call cuMemAlloc (params.rs:58-78, multiexp.rs:42-44): one grow-only arena per device;
//   * no per-call module/function lookup or stream creation: kernels are linked in, one stream
//     per device lives with the context;
//   * the window combine and the cross-device sum run on the device; one point comes back.
// There is no CPU fallback: without a CUDA device every entry point fails with MSM_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/msm_b200.h"
#include "kernels.cuh"

using namespace msm;

namespace {

thread_local std::string g_create_error;

#define CU_TRY(ctx, call)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error(ctx, std::string(#call) + ": " + cudaGetErrorString(_e));                   \
      return MSM_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

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
  Arena arena;        // per-call scratch (sorted entries, buckets, partials)
  Arena io;           // scalars staged from the host + result points
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint64_t launches = 0;
  void* small = nullptr;  // 64 KiB persistent (partial-point gather)
};

}  // namespace

struct msm_ctx {
  int curve = 0;
  std::vector<DeviceCtx> devs;
  std::atomic<int> in_use{0};
  const volatile int* abort_flag = nullptr;
  std::string err;
  msm_timings tm{};
  uint32_t window_override = 0;
};

struct msm_bases {
  msm_ctx* ctx = nullptr;
  size_t n = 0;
  struct Shard {
    int dev_idx;
    void* ptr;
    size_t start, n;
    bool owned;
  };
  std::vector<Shard> shards;
};

namespace {

void set_error(msm_ctx* ctx, const std::string& s) {
  if (ctx) ctx->err = s;
  else g_create_error = s;
}

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

inline size_t fq_bytes(int curve) { return curve == MSM_CURVE_BN254_G1 ? 32 : 48; }
inline uint32_t scalar_bits(int curve) { return curve == MSM_CURVE_BN254_G1 ? 254 : 255; }

bool aborted(msm_ctx* ctx) { return ctx->abort_flag && *ctx->abort_flag; }

// ---------------------------------------------------------------------------------------------
// Window choice: minimise  W * (chunk_len * M_madd + 2^(c-1) * 2 * M_add)  in field multiplies
// (M_madd = 10, M_add = 14), subject to the bucket array staying modest.  The reference leaves
// this to the caller (window_size argument) or to calc_window_size (ec-gpu-proxy/src/multiexp.rs
// :245-252); results never depend on it.
// ---------------------------------------------------------------------------------------------
uint32_t choose_window(uint32_t chunk_len, uint32_t bits, uint64_t n_tasks_lines, size_t xyzz_bytes) {
  double best = 1e300;
  uint32_t best_c = 2;
  for (uint32_t c = 2; c <= 22; c++) {
    const uint32_t W = (bits + 1 + c - 1) / c;
    const double B = (double)(1u << (c - 1));
    const double bucket_bytes = (double)n_tasks_lines * W * B * (double)xyzz_bytes;
    if (bucket_bytes > 6e9) break;
    const double cost = (double)W * ((double)chunk_len * 10.0 + B * (2.0 * 14.0 + 6.0));
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

struct Plan {
  Geometry geo;
  uint32_t n_lines, S, n_slices, Q, RW, PG, n_tasks;
  uint64_t E_max;
  size_t scratch_bytes;
};

template <class P>
int make_plan(msm_ctx* ctx, uint32_t L, uint32_t n_lines, uint32_t num_chunks, Plan& pl) {
  if (L == 0 || num_chunks == 0 || n_lines == 0 || num_chunks > L) return MSM_ERR_INVALID;
  Geometry& g = pl.geo;
  g.num_chunks = num_chunks;
  g.chunk_len = L / num_chunks;  // tail dropped, as ag-build/cl/multiexp.cl:235
  g.L = g.chunk_len * num_chunks;
  g.scalar_bits = scalar_bits(ctx->curve);
  uint32_t c = ctx->window_override;
  if (const char* env = getenv("MSM_B200_WINDOW")) {
    if (!c) c = (uint32_t)atoi(env);
  }
  if (c < 2 || c > 24) c = choose_window(g.chunk_len, g.scalar_bits, (uint64_t)num_chunks * n_lines, sizeof(Xyzz<P>));
  g.c = c;
  g.W = (g.scalar_bits + 1 + c - 1) / c;
  g.B = 1u << (c - 1);
  const uint64_t NB = (uint64_t)num_chunks * g.W * g.B;
  if (NB >= (1ull << 31)) return MSM_ERR_TOO_LARGE;
  g.NB = (uint32_t)NB;
  pl.n_lines = n_lines;
  pl.n_tasks = n_lines * num_chunks;
  pl.E_max = (uint64_t)g.L * g.W;
  if (pl.E_max >= (1ull << 32) || (uint64_t)L * n_lines >= (1ull << 31)) return MSM_ERR_TOO_LARGE;
  // slice length: enough slices to fill the machine several times over, few cut buckets
  uint32_t S = (uint32_t)(pl.E_max / (148ull * 512 * 8));
  if (const char* env = getenv("MSM_B200_SLICE")) S = (uint32_t)atoi(env);
  S = S < 8 ? 8 : (S > 1024 ? 1024 : S);
  pl.S = S;
  pl.n_slices = (uint32_t)((pl.E_max + S - 1) / S);
  pl.Q = g.B < 8 ? g.B : 8;
  const uint32_t TG = g.B / pl.Q;
  pl.RW = TG < 128 ? TG : 128;
  pl.PG = TG / pl.RW;
  const uint32_t n_tiles = (g.NB + SCAN_TILE - 1) / SCAN_TILE;
  size_t b = 0;
  b += Arena::padded((size_t)(g.NB + 1) * 4) * 3;                      // counts, bucket_start, cursor
  b += Arena::padded((size_t)(n_tiles + 1) * 4);                       // tile sums + grand total
  b += Arena::padded(pl.E_max * 4);                                    // entries
  b += Arena::padded((size_t)g.NB * n_lines * sizeof(Xyzz<P>));        // bucket accumulators
  b += Arena::padded((size_t)2 * pl.n_slices * n_lines * sizeof(Xyzz<P>));  // slice partials
  b += Arena::padded((size_t)pl.n_tasks * g.W * pl.PG * sizeof(Xyzz<P>));   // group partials
  pl.scratch_bytes = b;
  return MSM_OK;
}

// Enqueue one whole MSM batch on dc.stream.  d_scalars / d_out are device pointers.  No sync.
template <class P>
int enqueue_msm(msm_ctx* ctx, DeviceCtx& dc, const Plan& pl, const Affine<P>* d_bases,
                uint32_t line_stride, const uint32_t* d_scalars, Jacobian<P>* d_out, bool timed) {
  const Geometry& g = pl.geo;
  CU_TRY(ctx, dc.arena.ensure(pl.scratch_bytes));
  const uint32_t n_tiles = (g.NB + SCAN_TILE - 1) / SCAN_TILE;
  uint32_t* counts = dc.arena.take<uint32_t>(g.NB + 1);
  uint32_t* bucket_start = dc.arena.take<uint32_t>(g.NB + 1);
  uint32_t* cursor = dc.arena.take<uint32_t>(g.NB + 1);
  uint32_t* tile_sums = dc.arena.take<uint32_t>(n_tiles + 1);
  uint32_t* entries = dc.arena.take<uint32_t>(pl.E_max);
  Xyzz<P>* bucket_acc = dc.arena.take<Xyzz<P>>((size_t)g.NB * pl.n_lines);
  Xyzz<P>* partials = dc.arena.take<Xyzz<P>>((size_t)2 * pl.n_slices * pl.n_lines);
  Xyzz<P>* group_partials = dc.arena.take<Xyzz<P>>((size_t)pl.n_tasks * g.W * pl.PG);
  cudaStream_t st = dc.stream;

  if (timed) CU_TRY(ctx, cudaEventRecord(dc.ev[1], st));
  // --- sort: histogram, scan, scatter
  CU_TRY(ctx, cudaMemsetAsync(counts, 0, (size_t)(g.NB + 1) * 4, st));
  const uint32_t db = 256, dg = (g.L + db - 1) / db;
  k_digits<false><<<dg, db, 0, st>>>(d_scalars, g, counts, nullptr);
  k_scan_tiles<<<n_tiles, SCAN_BLOCK, 0, st>>>(counts, g.NB, bucket_start, tile_sums);
  k_scan_tile_sums<<<1, SCAN_BLOCK, 0, st>>>(tile_sums, n_tiles, tile_sums + n_tiles);
  k_scan_finish<<<(g.NB + 1 + 255) / 256, 256, 0, st>>>(bucket_start, g.NB, tile_sums, tile_sums + n_tiles, cursor);
  k_digits<true><<<dg, db, 0, st>>>(d_scalars, g, cursor, entries);
  if (timed) CU_TRY(ctx, cudaEventRecord(dc.ev[2], st));
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  // --- accumulate
  {
    const uint32_t tb = 128;
    dim3 grid((pl.n_slices + tb - 1) / tb, pl.n_lines);
    k_accumulate<P><<<grid, tb, 0, st>>>(d_bases, line_stride, entries, bucket_start, g.NB,
                                         bucket_start + g.NB, pl.S, pl.n_slices, bucket_acc, partials);
    dim3 fgrid((g.NB + tb - 1) / tb, pl.n_lines);
    k_fixup<P><<<fgrid, tb, 0, st>>>(bucket_start, g.NB, pl.S, pl.n_slices, bucket_acc, partials);
  }
  if (timed) CU_TRY(ctx, cudaEventRecord(dc.ev[3], st));
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  // --- reduce + combine
  {
    const uint32_t tb = 128;
    const uint64_t n_threads = (uint64_t)g.NB * pl.n_lines / pl.Q;
    k_bucket_reduce<P><<<(uint32_t)((n_threads + tb - 1) / tb), tb, tb * sizeof(Xyzz<P>), st>>>(
        bucket_acc, (uint32_t)n_threads, g.B, pl.Q, pl.RW, group_partials);
    const uint32_t wt = g.W < 32 ? 32 : (g.W > 256 ? 256 : ((g.W + 31) / 32) * 32);
    k_window_combine<P><<<pl.n_tasks, wt, (size_t)g.W * sizeof(Xyzz<P>), st>>>(group_partials, g.W, pl.PG, g.c, d_out);
  }
  if (timed) CU_TRY(ctx, cudaEventRecord(dc.ev[4], st));
  dc.launches += 10;
  CU_TRY(ctx, cudaGetLastError());
  return MSM_OK;
}

void collect_timings(msm_ctx* ctx, DeviceCtx& dc, const Plan& pl, bool have_h2d) {
  msm_timings& t = ctx->tm;
  memset(&t, 0, sizeof(t));
  if (have_h2d) cudaEventElapsedTime(&t.h2d_ms, dc.ev[0], dc.ev[1]);
  cudaEventElapsedTime(&t.sort_ms, dc.ev[1], dc.ev[2]);
  cudaEventElapsedTime(&t.accumulate_ms, dc.ev[2], dc.ev[3]);
  cudaEventElapsedTime(&t.reduce_ms, dc.ev[3], dc.ev[4]);
  cudaEventElapsedTime(&t.total_ms, dc.ev[1], dc.ev[4]);
  t.window_bits = pl.geo.c;
  t.num_windows = pl.geo.W;
  t.num_entries = pl.E_max;
  t.kernel_launches = dc.launches;
}

template <class P>
int multiple_multiexp_impl(msm_ctx* ctx, const msm_bases* bases, const void* scalars, size_t L,
                           uint32_t num_chunks, void* out, bool device_io) {
  if (bases->shards.size() != 1 || bases->shards[0].dev_idx != 0) {
    set_error(ctx, "multiple_multiexp needs bases resident on device 0 (msm_bases_upload)");
    return MSM_ERR_INVALID;
  }
  if (L == 0 || L > bases->n || L >= (1ull << 31)) return MSM_ERR_INVALID;
  const uint32_t n_lines = (uint32_t)(bases->n / L);  // ag-cuda-ec/src/multiexp.rs:28-30
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  Plan pl;
  int rc = make_plan<P>(ctx, (uint32_t)L, n_lines, num_chunks, pl);
  if (rc) return rc;
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  const uint32_t* d_scalars;
  Jacobian<P>* d_out;
  const size_t out_bytes = (size_t)pl.n_tasks * sizeof(Jacobian<P>);
  if (device_io) {
    d_scalars = static_cast<const uint32_t*>(scalars);
    d_out = static_cast<Jacobian<P>*>(out);
  } else {
    CU_TRY(ctx, dc.io.ensure(Arena::padded(L * 32) + Arena::padded(out_bytes)));
    uint32_t* ds = dc.io.take<uint32_t>(L * 8);
    d_out = dc.io.take<Jacobian<P>>(pl.n_tasks);
    CU_TRY(ctx, cudaEventRecord(dc.ev[0], dc.stream));
    CU_TRY(ctx, cudaMemcpyAsync(ds, scalars, L * 32, cudaMemcpyHostToDevice, dc.stream));
    d_scalars = ds;
  }
  rc = enqueue_msm<P>(ctx, dc, pl, static_cast<const Affine<P>*>(bases->shards[0].ptr), (uint32_t)L,
                      d_scalars, d_out, true);
  if (rc) {
    cudaStreamSynchronize(dc.stream);
    return rc;
  }
  if (!device_io) CU_TRY(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, dc.stream));
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  collect_timings(ctx, dc, pl, !device_io);
  return MSM_OK;
}

// One MSM split over all devices of the context (MultiexpKernel::multiexp).  Either host bases
// (uploaded per call, like the reference does) or resident sharded bases.
template <class P>
int multiexp_impl(msm_ctx* ctx, const void* host_bases, const msm_bases* resident, size_t skip,
                  const void* scalars, size_t n, void* out) {
  const size_t n_dev = ctx->devs.size();
  if (n >= (1ull << 31)) return MSM_ERR_TOO_LARGE;
  if (n == 0) {
    Jacobian<P> inf;
    inf.x = fp_zero<P>();
    inf.y = fp_one<P>();
    inf.z = fp_zero<P>();
    memcpy(out, &inf, sizeof(inf));
    return MSM_OK;
  }
  // work list: (device, bases pointer or host pointer, scalar offset, count)
  struct Job {
    size_t dev_idx;
    const void* d_bases;   // device pointer when resident
    const char* h_bases;   // host pointer when not
    size_t s_off, cnt;
  };
  std::vector<Job> jobs;
  if (resident) {
    if (skip + n > resident->n) return MSM_ERR_INVALID;
    for (const auto& sh : resident->shards) {
      const size_t lo = std::max(skip, sh.start), hi = std::min(skip + n, sh.start + sh.n);
      if (lo >= hi) continue;
      jobs.push_back({(size_t)sh.dev_idx,
                      static_cast<const char*>(sh.ptr) + (lo - sh.start) * sizeof(Affine<P>), nullptr,
                      lo - skip, hi - lo});
    }
  } else {
    const size_t chunk = (n + n_dev - 1) / n_dev;  // ec-gpu-proxy/src/multiexp.rs:329-337
    for (size_t d = 0; d * chunk < n; d++) {
      const size_t cnt = std::min(chunk, n - d * chunk);
      jobs.push_back({d, nullptr, static_cast<const char*>(host_bases) + d * chunk * sizeof(Affine<P>),
                      d * chunk, cnt});
    }
  }
  std::vector<int> rcs(jobs.size(), MSM_OK);
  std::vector<std::string> errs(jobs.size());
  std::vector<Plan> plans(jobs.size());
  std::vector<Jacobian<P>*> d_partials(jobs.size(), nullptr);
  auto run_job = [&](size_t j) {
    const Job& job = jobs[j];
    DeviceCtx& dc = ctx->devs[job.dev_idx];
    auto fail = [&](cudaError_t e, const char* what) {
      errs[j] = std::string(what) + ": " + cudaGetErrorString(e);
      rcs[j] = MSM_ERR_CUDA;
    };
    cudaError_t e = cudaSetDevice(dc.dev);
    if (e != cudaSuccess) return fail(e, "cudaSetDevice");
    msm_ctx shadow;  // per-thread error sink (ctx->err is not thread-safe)
    shadow.curve = ctx->curve;
    shadow.window_override = ctx->window_override;
    shadow.abort_flag = ctx->abort_flag;
    int rc = make_plan<P>(&shadow, (uint32_t)job.cnt, 1, 1, plans[j]);
    if (rc) { rcs[j] = rc; return; }
    const size_t bases_bytes = job.h_bases ? job.cnt * sizeof(Affine<P>) : 0;
    e = dc.io.ensure(Arena::padded(job.cnt * 32) + Arena::padded(sizeof(Jacobian<P>)) + Arena::padded(bases_bytes));
    if (e != cudaSuccess) return fail(e, "io arena");
    uint32_t* ds = dc.io.take<uint32_t>(job.cnt * 8);
    Jacobian<P>* d_out = dc.io.take<Jacobian<P>>(1);
    const Affine<P>* d_bases = static_cast<const Affine<P>*>(job.d_bases);
    cudaEventRecord(dc.ev[0], dc.stream);
    if (job.h_bases) {
      Affine<P>* db = dc.io.take<Affine<P>>(job.cnt);
      e = cudaMemcpyAsync(db, job.h_bases, bases_bytes, cudaMemcpyHostToDevice, dc.stream);
      if (e != cudaSuccess) return fail(e, "bases H2D");
      d_bases = db;
    }
    e = cudaMemcpyAsync(ds, static_cast<const char*>(scalars) + job.s_off * 32, job.cnt * 32,
                        cudaMemcpyHostToDevice, dc.stream);
    if (e != cudaSuccess) return fail(e, "scalars H2D");
    rc = enqueue_msm<P>(&shadow, dc, plans[j], d_bases, (uint32_t)job.cnt, ds, d_out, true);
    if (rc) { rcs[j] = rc; errs[j] = shadow.err; cudaStreamSynchronize(dc.stream); return; }
    d_partials[j] = d_out;
  };
  if (jobs.size() == 1) {
    run_job(0);
  } else {
    // one host thread per device, as parallel_multiexp does (ec-gpu-proxy/src/multiexp.rs:346)
    std::vector<std::thread> th;
    for (size_t j = 0; j < jobs.size(); j++) th.emplace_back(run_job, j);
    for (auto& t : th) t.join();
  }
  for (size_t j = 0; j < jobs.size(); j++) {
    if (rcs[j] != MSM_OK) {  // first error wins (ec-gpu-proxy/src/multiexp.rs:351-364)
      for (auto& job : jobs) {
        cudaSetDevice(ctx->devs[job.dev_idx].dev);
        cudaStreamSynchronize(ctx->devs[job.dev_idx].stream);
      }
      set_error(ctx, errs[j]);
      return rcs[j];
    }
  }
  // gather the per-device partial points on device 0 over peer copies, sum there
  DeviceCtx& d0 = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(d0.dev));
  Jacobian<P>* gather = nullptr;
  Jacobian<P>* d_result = nullptr;
  {
    // wait for every device first
    for (size_t j = 0; j < jobs.size(); j++) {
      DeviceCtx& dc = ctx->devs[jobs[j].dev_idx];
      CU_TRY(ctx, cudaSetDevice(dc.dev));
      CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
    }
    CU_TRY(ctx, cudaSetDevice(d0.dev));
    gather = reinterpret_cast<Jacobian<P>*>(d0.small);
    d_result = gather + jobs.size();
    for (size_t j = 0; j < jobs.size(); j++) {
      DeviceCtx& dc = ctx->devs[jobs[j].dev_idx];
      if (dc.dev == d0.dev) {
        CU_TRY(ctx, cudaMemcpyAsync(gather + j, d_partials[j], sizeof(Jacobian<P>), cudaMemcpyDeviceToDevice, d0.stream));
      } else {
        CU_TRY(ctx, cudaMemcpyPeerAsync(gather + j, d0.dev, d_partials[j], dc.dev, sizeof(Jacobian<P>), d0.stream));
      }
    }
    k_sum_points<P><<<1, 32, 0, d0.stream>>>(gather, (uint32_t)jobs.size(), d_result);
    d0.launches += 1;
    CU_TRY(ctx, cudaMemcpyAsync(out, d_result, sizeof(Jacobian<P>), cudaMemcpyDeviceToHost, d0.stream));
    CU_TRY(ctx, cudaStreamSynchronize(d0.stream));
  }
  // timings of device 0's share
  for (size_t j = 0; j < jobs.size(); j++)
    if (jobs[j].dev_idx == 0) collect_timings(ctx, d0, plans[j], true);
  return MSM_OK;
}

template <class P> Affine<P> host_generator() {
  // canonical generator -> Montgomery with the host build of fp.cuh
  Affine<P> g;
  g.x = fp_zero<P>();
  g.y = fp_zero<P>();
  if (P::N == 8) {
    g.x.v[0] = 1;
    g.y.v[0] = 2;
  } else {
    const uint32_t gx[12] = {0xdb22c6bbu, 0xfb3af00au, 0xf97a1aefu, 0x6c55e83fu, 0x171bac58u, 0xa14e3a3fu,
                             0x9774b905u, 0xc3688c4fu, 0x4fa9ac0fu, 0x2695638cu, 0x3197d794u, 0x17f1d3a7u};
    const uint32_t gy[12] = {0x46c5e7e1u, 0x0caa2329u, 0xa2888ae4u, 0xd03cc744u, 0x2c04b3edu, 0x00db18cbu,
                             0xd5d00af6u, 0xfcf5e095u, 0x741d8ae4u, 0xa09e30edu, 0xe3aaa0f1u, 0x08b3f481u};
    for (int i = 0; i < 12; i++) {
      g.x.v[i] = gx[i];
      g.y.v[i] = gy[i];
    }
  }
  g.x = fp_to_mont<P>(g.x);
  g.y = fp_to_mont<P>(g.y);
  return g;
}

template <class P> int synth_points_impl(msm_ctx* ctx, uint64_t seed, size_t start, size_t n, void* d_out) {
  if (n == 0) return MSM_OK;
  if (n >= (1ull << 32)) return MSM_ERR_TOO_LARGE;
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const uint64_t a = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ull, 0);
  const uint64_t b = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ull, 1) | 1;
  // D = b*G on the host (64 doublings; setup only)
  Affine<P> gen = host_generator<P>();
  Xyzz<P> acc = xyzz_inf<P>();
  for (int bit = 63; bit >= 0; bit--) {
    acc = xyzz_dbl<P>(acc);
    if ((b >> bit) & 1) xyzz_madd<P>(acc, gen);
  }
  Affine<P> d = xyzz_to_affine<P>(acc);
  const uint32_t threads = (uint32_t)((n + SYNTH_RUN - 1) / SYNTH_RUN);
  k_synth_points<P><<<(threads + 63) / 64, 64, 0, dc.stream>>>(gen, d, a, b, (uint64_t)start, (uint32_t)n,
                                                              static_cast<Affine<P>*>(d_out));
  dc.launches += 1;
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

ScalarField scalar_field(int curve) {
  ScalarField f;
  if (curve == MSM_CURVE_BN254_G1) {
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

template <class P>
int test_fq_impl(msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const size_t bytes = count * sizeof(Fp<P>);
  CU_TRY(ctx, dc.io.ensure(3 * Arena::padded(bytes)));
  Fp<P>* da = dc.io.take<Fp<P>>(count);
  Fp<P>* db = dc.io.take<Fp<P>>(count);
  Fp<P>* dout = dc.io.take<Fp<P>>(count);
  CU_TRY(ctx, cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, dc.stream));
  if (b) CU_TRY(ctx, cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, dc.stream));
  k_test_fq<P><<<(uint32_t)((count + 127) / 128), 128, 0, dc.stream>>>(op, da, b ? db : nullptr, dout, (uint32_t)count);
  dc.launches += 1;
  CU_TRY(ctx, cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, dc.stream));
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

template <class P>
int test_ec_impl(msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const size_t ja = count * sizeof(Jacobian<P>);
  const size_t bb = op == 0 ? ja : (op == 1 ? count * sizeof(Affine<P>) : 0);
  CU_TRY(ctx, dc.io.ensure(3 * Arena::padded(ja)));
  Jacobian<P>* da = dc.io.take<Jacobian<P>>(count);
  void* db = dc.io.take<Jacobian<P>>(count);
  Jacobian<P>* dout = dc.io.take<Jacobian<P>>(count);
  CU_TRY(ctx, cudaMemcpyAsync(da, a, ja, cudaMemcpyHostToDevice, dc.stream));
  if (bb) CU_TRY(ctx, cudaMemcpyAsync(db, b, bb, cudaMemcpyHostToDevice, dc.stream));
  k_test_ec<P><<<(uint32_t)((count + 63) / 64), 64, 0, dc.stream>>>(op, da, db, dout, (uint32_t)count);
  dc.launches += 1;
  CU_TRY(ctx, cudaMemcpyAsync(out, dout, ja, cudaMemcpyDeviceToHost, dc.stream));
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

template <class P>
int to_affine_impl(msm_ctx* ctx, const void* jac, size_t count, int mont_out, void* out_xy, uint8_t* out_inf) {
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const size_t jb = count * sizeof(Jacobian<P>), ab = count * sizeof(Affine<P>);
  CU_TRY(ctx, dc.io.ensure(Arena::padded(jb) + Arena::padded(ab) + Arena::padded(count)));
  Jacobian<P>* dj = dc.io.take<Jacobian<P>>(count);
  Affine<P>* da = dc.io.take<Affine<P>>(count);
  uint8_t* di = dc.io.take<uint8_t>(count);
  CU_TRY(ctx, cudaMemcpyAsync(dj, jac, jb, cudaMemcpyHostToDevice, dc.stream));
  k_to_affine<P><<<(uint32_t)((count + 63) / 64), 64, 0, dc.stream>>>(dj, (uint32_t)count, mont_out, da, di);
  dc.launches += 1;
  CU_TRY(ctx, cudaMemcpyAsync(out_xy, da, ab, cudaMemcpyDeviceToHost, dc.stream));
  if (out_inf) CU_TRY(ctx, cudaMemcpyAsync(out_inf, di, count, cudaMemcpyDeviceToHost, dc.stream));
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

template <class P> int sum_points_impl(msm_ctx* ctx, const void* d_in, size_t count, void* d_out) {
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  k_sum_points<P><<<1, 32, 0, dc.stream>>>(static_cast<const Jacobian<P>*>(d_in), (uint32_t)count,
                                          static_cast<Jacobian<P>*>(d_out));
  dc.launches += 1;
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

#define CURVE_DISPATCH(ctx, EXPR_BN, EXPR_BLS) \
  ((ctx)->curve == MSM_CURVE_BN254_G1 ? (EXPR_BN) : (EXPR_BLS))

#define LOCK_OR_BUSY(ctx)            \
  if (!(ctx)) return MSM_ERR_INVALID; \
  CtxLock _lock(ctx);                \
  if (!_lock.ok) return MSM_ERR_BUSY

}  // namespace

// =================================================================================================
extern "C" {

const char* msm_version(void) { return "msm_b200 0.1 (sm_100a)"; }

int msm_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int msm_ctx_create(int curve, const int* device_ids, int n_devices, msm_ctx** out) {
  if (!out || (curve != MSM_CURVE_BN254_G1 && curve != MSM_CURVE_BLS12_381_G1) || n_devices < 0) {
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
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&dc.stream, cudaStreamNonBlocking);
    for (int k = 0; k < 5 && e == cudaSuccess; k++) e = cudaEventCreate(&dc.ev[k]);
    if (e == cudaSuccess) e = cudaMalloc(&dc.small, 65536);
    if (e != cudaSuccess) {
      // a device whose kernel cannot be initialised is skipped (ec-gpu-proxy/src/multiexp.rs:288-303)
      set_error(nullptr, std::string("device init failed: ") + cudaGetErrorString(e));
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
  if (!ctx) return MSM_ERR_INVALID;
  for (auto& dc : ctx->devs) {
    cudaSetDevice(dc.dev);
    if (dc.stream) cudaStreamSynchronize(dc.stream);
    dc.arena.release();
    dc.io.release();
    if (dc.small) cudaFree(dc.small);
    for (auto& e : dc.ev)
      if (e) cudaEventDestroy(e);
    if (dc.stream) cudaStreamDestroy(dc.stream);
  }
  delete ctx;
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

int msm_set_window_bits(msm_ctx* ctx, uint32_t c) {
  if (!ctx || (c != 0 && (c < 2 || c > 24))) return MSM_ERR_INVALID;
  ctx->window_override = c;
  return MSM_OK;
}

static int upload_common(msm_ctx* ctx, const void* xy, size_t n, bool sharded, bool wrap, msm_bases** out) {
  if (!ctx || !out || (!xy && n)) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  const size_t pt = 2 * fq_bytes(ctx->curve);
  msm_bases* b = new msm_bases();
  b->ctx = ctx;
  b->n = n;
  const size_t n_dev = sharded ? ctx->devs.size() : 1;
  const size_t chunk = n_dev ? (n + n_dev - 1) / n_dev : 0;
  for (size_t d = 0; d < n_dev; d++) {
    const size_t start = d * chunk;
    if (start >= n && !(d == 0)) break;
    const size_t cnt = start < n ? std::min(chunk, n - start) : 0;
    msm_bases::Shard sh{(int)d, nullptr, start, cnt, !wrap};
    if (wrap) {
      sh.ptr = const_cast<void*>(xy);
    } else if (cnt) {
      DeviceCtx& dc = ctx->devs[d];
      cudaError_t e = cudaSetDevice(dc.dev);
      if (e == cudaSuccess) e = cudaMalloc(&sh.ptr, cnt * pt);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(sh.ptr, static_cast<const char*>(xy) + start * pt, cnt * pt, cudaMemcpyHostToDevice, dc.stream);
      // the reference returns without synchronising (ag-cuda-proxy/src/params.rs:186-201); the
      // source slice may be freed by the caller right after, so this implementation waits.
      if (e == cudaSuccess) e = cudaStreamSynchronize(dc.stream);
      if (e != cudaSuccess) {
        set_error(ctx, std::string("msm_bases_upload: ") + cudaGetErrorString(e));
        b->shards.push_back(sh);
        for (auto& s : b->shards)
          if (s.owned && s.ptr) cudaFree(s.ptr);
        delete b;
        return MSM_ERR_CUDA;
      }
    }
    b->shards.push_back(sh);
  }
  *out = b;
  return MSM_OK;
}

int msm_bases_upload(msm_ctx* ctx, const void* xy, size_t n, msm_bases** out) {
  return upload_common(ctx, xy, n, false, false, out);
}
int msm_bases_upload_sharded(msm_ctx* ctx, const void* xy, size_t n, msm_bases** out) {
  return upload_common(ctx, xy, n, true, false, out);
}
int msm_bases_wrap_device(msm_ctx* ctx, const void* d_xy, size_t n, msm_bases** out) {
  return upload_common(ctx, d_xy, n, false, true, out);
}
size_t msm_bases_size_bytes(const msm_bases* b) { return b ? b->n * 2 * fq_bytes(b->ctx->curve) : 0; }
size_t msm_bases_num_points(const msm_bases* b) { return b ? b->n : 0; }
int msm_bases_free(msm_bases* b) {
  if (!b) return MSM_ERR_INVALID;
  for (auto& s : b->shards) {
    if (s.owned && s.ptr) {
      cudaSetDevice(b->ctx->devs[s.dev_idx].dev);
      cudaFree(s.ptr);
    }
  }
  delete b;
  return MSM_OK;
}

int msm_multiple_multiexp(msm_ctx* ctx, const msm_bases* bases, const void* scalars, size_t L,
                          uint32_t num_chunks, uint32_t window_hint, int neg_is_cheap, void* out) {
  (void)window_hint;
  (void)neg_is_cheap;
  if (!ctx || !bases || !scalars || !out || bases->ctx != ctx) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  return CURVE_DISPATCH(ctx, multiple_multiexp_impl<Bn254Fq>(ctx, bases, scalars, L, num_chunks, out, false),
                        multiple_multiexp_impl<Bls381Fq>(ctx, bases, scalars, L, num_chunks, out, false));
}

int msm_multiple_multiexp_device(msm_ctx* ctx, const msm_bases* bases, const void* d_scalars, size_t L,
                                 uint32_t num_chunks, void* d_out) {
  if (!ctx || !bases || !d_scalars || !d_out || bases->ctx != ctx) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  return CURVE_DISPATCH(ctx, multiple_multiexp_impl<Bn254Fq>(ctx, bases, d_scalars, L, num_chunks, d_out, true),
                        multiple_multiexp_impl<Bls381Fq>(ctx, bases, d_scalars, L, num_chunks, d_out, true));
}

int msm_multiexp(msm_ctx* ctx, const void* bases_xy, const void* scalars, size_t n, void* out) {
  if (!ctx || !out || (n && (!bases_xy || !scalars))) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  return CURVE_DISPATCH(ctx, multiexp_impl<Bn254Fq>(ctx, bases_xy, nullptr, 0, scalars, n, out),
                        multiexp_impl<Bls381Fq>(ctx, bases_xy, nullptr, 0, scalars, n, out));
}

int msm_multiexp_resident(msm_ctx* ctx, const msm_bases* bases, size_t skip, const void* scalars, size_t n,
                          void* out) {
  if (!ctx || !bases || !out || (n && !scalars) || bases->ctx != ctx) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  return CURVE_DISPATCH(ctx, multiexp_impl<Bn254Fq>(ctx, nullptr, bases, skip, scalars, n, out),
                        multiexp_impl<Bls381Fq>(ctx, nullptr, bases, skip, scalars, n, out));
}

int msm_sum_points_device(msm_ctx* ctx, const void* d_jac, size_t count, void* d_out) {
  if (!ctx || !d_jac || !d_out) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  return CURVE_DISPATCH(ctx, sum_points_impl<Bn254Fq>(ctx, d_jac, count, d_out),
                        sum_points_impl<Bls381Fq>(ctx, d_jac, count, d_out));
}

int msm_to_affine(msm_ctx* ctx, const void* jac, size_t count, int mont_out, void* out_xy, uint8_t* out_inf) {
  if (!ctx || !jac || !out_xy) return MSM_ERR_INVALID;
  if (count == 0) return MSM_OK;
  LOCK_OR_BUSY(ctx);
  return CURVE_DISPATCH(ctx, to_affine_impl<Bn254Fq>(ctx, jac, count, mont_out, out_xy, out_inf),
                        to_affine_impl<Bls381Fq>(ctx, jac, count, mont_out, out_xy, out_inf));
}

int msm_synth_points_device(msm_ctx* ctx, uint64_t seed, size_t start, size_t n, void* d_xy) {
  if (!ctx || (!d_xy && n)) return MSM_ERR_INVALID;
  LOCK_OR_BUSY(ctx);
  return CURVE_DISPATCH(ctx, synth_points_impl<Bn254Fq>(ctx, seed, start, n, d_xy),
                        synth_points_impl<Bls381Fq>(ctx, seed, start, n, d_xy));
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

int msm_test_fq_op(msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
  if (!ctx || !a || !out || op < 0 || op > 8) return MSM_ERR_INVALID;
  if (count == 0) return MSM_OK;
  LOCK_OR_BUSY(ctx);
  return CURVE_DISPATCH(ctx, test_fq_impl<Bn254Fq>(ctx, op, a, b, out, count),
                        test_fq_impl<Bls381Fq>(ctx, op, a, b, out, count));
}

int msm_test_ec_op(msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
  if (!ctx || !a || !out || op < 0 || op > 2 || (op < 2 && !b)) return MSM_ERR_INVALID;
  if (count == 0) return MSM_OK;
  LOCK_OR_BUSY(ctx);
  return CURVE_DISPATCH(ctx, test_ec_impl<Bn254Fq>(ctx, op, a, b, out, count),
                        test_ec_impl<Bls381Fq>(ctx, op, a, b, out, count));
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
