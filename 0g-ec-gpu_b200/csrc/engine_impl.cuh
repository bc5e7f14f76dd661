// engine_impl.cuh -- launch sequencing of the MSM pipeline, templated on the field class.
// Instantiated once per (curve, field implementation) in inst_*.cu and reached from the C ABI
// (engine.cu) through the FieldOps table.
//
// Replaces the host wrapper ag_cuda_ec::multiple_multiexp (ag-cuda-ec/src/multiexp.rs:22-81) and
// the per-device part of ec_gpu_proxy::MultiexpKernel (ec-gpu-proxy/src/multiexp.rs:135-253,
// 324-400).  Differences that are the point of the rewrite: no per-call cuMemAlloc, no per-call
// function lookup or stream creation, window combine and cross-device sum on the device.
#pragma once
#include <type_traits>

#include "engine_common.h"
#include "kernels.cuh"
#include "bucket_affine.cuh"
#include "ecfft.cuh"

namespace msm {

// ---------------------------------------------------------------------------------------------
// Window choice: minimise  W * (chunk_len * M_madd + 2^(c-1) * (2 * M_add + overhead))  in field
// multiplies (M_madd = 10, M_add = 14), subject to the bucket array staying modest.  The
// reference leaves this to the caller (window_size argument) or to calc_window_size
// (ec-gpu-proxy/src/multiexp.rs:245-252); results never depend on it.
// ---------------------------------------------------------------------------------------------
// cost of one bucket in the reduction, in field products (measured): with very many (task, window)
// groups one thread owns a whole group and pays the two additions of the running sum; with few
// groups the per-thread fix-up, the shared-memory tree and low occupancy triple that
// (round 2, job r2_run13: 1024 .. 20480 groups reduce a bucket in 0.56 - 0.73 ns = 37 - 49 products' worth of time)
inline double reduce_cost_per_bucket(double n_groups) { return n_groups >= 32768.0 ? 34.0 : (n_groups >= 512.0 ? 40.0 : 90.0); }
// (A latency floor for the whole reduction -- "below ~2.4e7 bucket-products a smaller window buys nothing" -- was tried
// and is wrong: the reduction keeps growing with the bucket count in that range too.  BN254 2^20 on tables of
// c = 17 / 18 / 19 / 20: reduce 0.35 / 0.41 / 0.51 / 0.64 ms, call 2.96 / 3.15 / 3.07 / 2.93 ms; BLS12-381 2^19 on
// c = 16 ... 20: reduce 0.78 / 0.81 / 1.01 / 1.36 / 1.58 ms, call 4.00 / 4.28 / 4.41 / 4.48 / 4.40 ms
// (profiles/r02_window_sweep.log): the windows this model picks, 17 and 16, are within 1 % of the best.)

inline uint32_t choose_window(uint32_t chunk_len, uint32_t bits, uint64_t n_tasks_lines, size_t xyzz_bytes,
                              double* cost_out = nullptr) {
  double best = 1e300;
  uint32_t best_c = 2;
  for (uint32_t c = 2; c <= 22; c++) {
    const uint32_t W = (bits + 1 + c - 1) / c;
    const double B = (double)(1u << (c - 1));
    const double bucket_bytes = (double)n_tasks_lines * W * B * (double)xyzz_bytes;
    if (bucket_bytes > 6e9) break;
    const double per_bucket = reduce_cost_per_bucket((double)n_tasks_lines * W);
    const double cost = (double)W * ((double)chunk_len * 10.0 + B * per_bucket);
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  if (cost_out) *cost_out = best;
  return best_c;
}

// Same figure for a folded window table of window size c: the W digits of a point all land in the
// task's single bucket set, so the per-window reduction disappears.
inline double fold_cost(uint32_t c, uint32_t chunk_len, uint32_t bits, uint64_t n_tasks_lines) {
  const uint32_t W = (bits + 1 + c - 1) / c;
  return (double)W * (double)chunk_len * 10.0 + (double)(1u << (c - 1)) * reduce_cost_per_bucket((double)n_tasks_lines);
}
// Window size of a table meant for tasks of chunk_len points (msm_bases_precompute_chunked).
inline uint32_t choose_table_window(uint32_t chunk_len, uint32_t bits, uint64_t n_tasks_lines, uint64_t n_points) {
  double best = 1e300;
  uint32_t best_c = 0;
  for (uint32_t c = 8; c <= 24; c++) {
    const uint32_t W = (bits + 1 + c - 1) / c;
    if ((uint64_t)W * n_points >= (1ull << 31) || W > (uint32_t)TABLE_MAX_W) continue;
    if ((double)n_tasks_lines * (double)(1u << (c - 1)) * 128.0 > 12e9) break;
    const double cost = fold_cost(c, chunk_len, bits, n_tasks_lines);
    if (cost < best) {
      best = cost;
      best_c = c;
    }
  }
  return best_c;
}

// which curve a field class belongs to, from its word count: 8 / 12 = BN254 / BLS12-381 Fq, 16 / 24 = their Fq2
template <class F> constexpr bool is_bn254() { return F::API_WORDS == 8 || F::API_WORDS == 16; }
template <class F> constexpr bool is_ext2() { return F::API_WORDS == 16 || F::API_WORDS == 24; }
template <class F> using BaseParams = typename std::conditional<is_bn254<F>(), Bn254Fq, Bls381Fq>::type;

// G1 fields on 32-bit limbs only (the Fq2 instantiations have no registers to spare)
template <class F> constexpr bool ba_supported() { return F::N % 4 == 0 && F::N <= 12 && F::API_WORDS == F::N; }

// table_c != 0: the bases are a window table built for window size table_c covering exactly L points
template <class F>
int make_plan(msm_ctx* ctx, uint32_t L, uint32_t n_lines, uint32_t num_chunks, Plan& pl, uint32_t table_c = 0,
              uint32_t n_sub = 1, uint32_t table_stride = 0, double sub_ratio = 2.0) {
  if (L == 0 || num_chunks == 0 || n_lines == 0 || num_chunks > L) return MSM_ERR_INVALID;
  if (n_lines > 65535) {  // the line index is gridDim.y of the bucket kernels
    set_error(ctx, "multiple_multiexp: more than 65535 base lines per call");
    return MSM_ERR_TOO_LARGE;
  }
  Geometry& g = pl.geo;
  g.fold = table_c ? 1 : 0;
  g.table_stride = table_stride ? table_stride : L;
  g.point_offset = 0;
  g.mont = ctx->scalars_mont ? (curve_is_bn254(ctx->curve) ? 1u : 2u) : 0u;
  if (n_lines != 1 || n_sub < 1 || (num_chunks > 1 && num_chunks < 2 * n_sub)) n_sub = 1;
  pl.n_sub = n_sub;
  pl.by_task = num_chunks > 1 && n_sub > 1;  // sub-batches are groups of whole tasks with bucket ranges of their own
  g.num_chunks = num_chunks;
  g.chunk_len = L / num_chunks;  // tail dropped, as ag-build/cl/multiexp.cl:235
  g.L = g.chunk_len * num_chunks;
  g.scalar_bits = scalar_bits(ctx->curve);
  {
    // Sub-batch boundaries.  Host scalars arrive back to back on the copy stream; sub-batch k can start when its
    // last byte has landed and sub-batch k-1 is done.  With sizes growing by a factor q the first (small) upload
    // is all that is exposed as long as q stays below the upload : compute speed ratio (~3.4 on one B200 with
    // its own PCIe link, 1.8 with eight GPUs pulling from one host: tools/numa_h2d_probe.py).  q = 2 until the
    // caller has measured both speeds (multiple_multiexp_impl).  MSM_B200_PIPELINE_RATIO=1 gives equal sizes.
    double q = sub_ratio;
    if (const char* env = getenv("MSM_B200_PIPELINE_RATIO")) q = atof(env);
    if (q < 1.0) q = 1.0;
    double total = 0, w = 1;
    for (uint32_t k = 0; k < n_sub; k++, w *= q) total += w;
    double acc = 0;
    w = 1;
    pl.sub_first[0] = 0;
    pl.sub_max = 0;
    // boundaries in scalars, or in whole tasks (at least one per group) when the sub-batches are task groups
    const uint32_t units = pl.by_task ? num_chunks : g.L, unit_len = pl.by_task ? g.chunk_len : 1;
    for (uint32_t k = 0; k < n_sub; k++, w *= q) {
      acc += w;
      uint32_t end = k + 1 == n_sub ? units : (uint32_t)((double)units * (acc / total));
      const uint32_t prev = pl.sub_first[k] / unit_len;
      if (end < prev) end = prev;
      if (pl.by_task) {
        if (end <= prev) end = prev + 1;
        if (end > units - (n_sub - 1 - k)) end = units - (n_sub - 1 - k);
      }
      if (end > units) end = units;
      pl.sub_first[k + 1] = end * unit_len;
      pl.sub_max = std::max(pl.sub_max, pl.sub_first[k + 1] - pl.sub_first[k]);
    }
    for (uint32_t k = n_sub + 1; k < 9; k++) pl.sub_first[k] = g.L;
  }
  uint32_t c = ctx->window_override;
  if (const char* env = getenv("MSM_B200_WINDOW")) {
    if (!c) c = (uint32_t)atoi(env);
  }
  if (c < 2 || c > 24) c = choose_window(g.chunk_len, g.scalar_bits, (uint64_t)num_chunks * n_lines, sizeof(Xyzz<F>));
  if (table_c) c = table_c;
  g.c = c;
  g.W = (g.scalar_bits + 1 + c - 1) / c;
  g.B = 1u << (c - 1);
  pl.W_sets = g.fold ? 1 : g.W;
  const uint64_t NB = (uint64_t)num_chunks * pl.W_sets * g.B;
  if (NB >= (1ull << 31)) return MSM_ERR_TOO_LARGE;
  g.NB = (uint32_t)NB;
  pl.n_lines = n_lines;
  pl.n_tasks = n_lines * num_chunks;
  pl.E_max = (uint64_t)g.L * g.W;
  if (pl.E_max >= (1ull << 31) || (uint64_t)L * n_lines >= (1ull << 31)) return MSM_ERR_TOO_LARGE;
  // Slice length: enough slices to fill the machine several times over, few cut buckets -- and a grid that is a
  // whole number of waves, rounded DOWN.  Every slice is the same amount of work, so the blocks of k_accumulate
  // finish wave by wave: 2^24 points at S = 332 made 4738 blocks = 8 waves of 592 and two blocks that ran a ninth
  // wave on an empty machine (29.41 ms); S = 333, 4724 blocks, is 28.71 ms (job r2_run24; S = 380 and 443, just
  // under 7 and 6 waves, measure the same, S = 300, 8.9 waves, 29.0).
  const uint64_t E_sub = (uint64_t)pl.sub_max * g.W;  // digits of the longest sub-batch
  static const int acc_bps = [] {
    int bps = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_accumulate<F>, 128, 0) != cudaSuccess || bps < 1) bps = 4;
    cudaGetLastError();
    return bps;
  }();
  static const uint64_t wave_slices = [] {
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
    cudaGetLastError();
    return (uint64_t)sms * acc_bps * 128;  // slices (threads) of one wave
  }();
  pl.acc_blocks_per_sm = (uint32_t)acc_bps;
  uint64_t waves = (148ull * 512 * 8 + wave_slices / 2) / wave_slices;
  if (waves < 1) waves = 1;
  {
    // Fewer, longer slices -- down to 5/8 of the waves -- while that gets them to 72 digits (short rows) or to 4/3 of
    // the average bucket: a slice shorter than a bucket cuts every bucket it touches, and every cut is a full addition
    // in k_fixup_cut.  BN254 2^21 points, c = 20 (52 digits per bucket): S = 45 in 8 waves 4.08 ms of accumulation,
    // S = 72 in 5 waves 3.99; BLS12-381 2^22, c = 20 (104 per bucket, k_fixup_cut 0.61 ms): S = 88 in 11 waves 21.11 ms
    // per call, S = 138 in 7 waves 20.82 (jobs r2_run25, r2_run40).
    const uint64_t avg_bucket = (pl.by_task ? pl.E_max : E_sub) / (NB ? NB : 1);  // a task group has its share of both
    const uint64_t s_target = std::max<uint64_t>(72, avg_bucket * 4 / 3);
    const uint64_t w_want = (E_sub + wave_slices * s_target / 2) / (wave_slices * s_target);  // nearest
    const uint64_t w_min = (waves * 5 + 7) / 8;
    if (w_want < waves) waves = w_want < w_min ? w_min : w_want;
  }
  uint32_t S = (uint32_t)((E_sub + wave_slices * waves - 1) / (wave_slices * waves));
  pl.wave_slices = (uint32_t)wave_slices;
  pl.waves = (uint32_t)waves;
  if (const char* env = getenv("MSM_B200_SLICE")) {
    S = (uint32_t)atoi(env);
    pl.waves = 0;
  }
  if (S < 8 || S > 1024) pl.waves = 0;
  S = S < 8 ? 8 : (S > 1024 ? 1024 : S);
  {
    // Few digits over few buckets (BLS12-381 shards of 2^19 points at c = 16: 256 digits per bucket, S = 13):
    // every bucket was cut into ~20 slices and went through the one-warp-per-bucket fix-up, which is built for a
    // FEW very heavy buckets -- 2.8 ms of a 6.7 ms call (profiles/r02_launches_bls_2p19.csv).  Slices of at least
    // an eighth of the average bucket keep a bucket's partial slots within what one thread adds up serially.
    const uint64_t avg_bucket = E_sub / (g.NB ? g.NB : 1);
    const uint32_t s_min = (uint32_t)(avg_bucket / 8 < 64 ? avg_bucket / 8 : 64);
    if (S < s_min && !getenv("MSM_B200_SLICE")) {
      S = s_min;
      pl.waves = 0;
    }
  }
  pl.S = S;
  pl.n_slices = (uint32_t)((pl.E_max + S - 1) / S);
  pl.slices_cap = (uint32_t)((E_sub + S - 1) / S) + g.W / S + 3;
  // buckets per reduction thread: the per-thread fix-up (first_weight * plain sum, a ~c-bit
  // double-and-add) is amortised over Q buckets; keep about one wave of threads on the machine
  {
    const uint64_t total_buckets = (uint64_t)g.NB * n_lines;
    uint32_t Q = 8;
    // measured: 2^21 buckets -> Q = 64, 2^18 -> Q = 8; beyond 2^21 buckets (many-task shapes) more threads beat
    // longer serial chains (AMT shape, 10.5 M buckets: Q = 64 5.6 ms, Q = 256 7.4 ms)
    while (Q < 64 && total_buckets / Q > 32768) Q <<= 1;
    if (const char* env = getenv("MSM_B200_REDUCE_Q")) Q = (uint32_t)atoi(env);
    while (Q > g.B) Q >>= 1;
    pl.Q = Q < 1 ? 1 : Q;
    // (A block tree carrying (sum_k (k+1) S, sum S, sum_j j P_j) instead of the per-thread first_weight x sum fix-up was
    // built and measured in round 2: 2.95 ms against 1.53 ms at 2^21 buckets -- the tree's partially filled warps cost
    // more issue slots than the fix-ups they replace.  An even split over exactly two blocks per SM instead of the
    // power-of-two Q -- 1024 warps on 592 schedulers -- changed nothing either: 1.58 ms, job r2_run17.)
  }
  const uint32_t TG = g.B / pl.Q;
  pl.RW = TG < 128 ? TG : 128;
  pl.PG = TG / pl.RW;
  const uint32_t n_tiles = (g.NB + SCAN_TILE - 1) / SCAN_TILE;
  size_t b = 0;
  // the sub-batches of a pipelined call run one after the other on the stream and share this scratch;
  // sizes are upper bounds for every sub-batch (n_sub == 1: the whole call)
  const size_t sort_sets = n_sub > 1 ? 2 : 1;  // sub-batch k+1 is sorted while k is accumulated from the other set
  b += sort_sets * Arena::padded((size_t)(g.NB + 1) * 4) * 3;                // counts, bucket_start, cursor
  b += sort_sets * Arena::padded((size_t)(n_tiles + 1) * 4);                 // tile sums + grand total
  b += sort_sets * Arena::padded((E_sub + g.W) * 4);                         // entries
  // measured slower than the bucket-range passes on B200 (2^24, c = 22: 6.8 vs 5.2 ms): off unless asked for
  pl.partition = false;
  if (const char* env = getenv("MSM_B200_PARTITION")) pl.partition = atoi(env) != 0;
  // binned sort for large calls: at most 1024 bins of at most 8192 buckets
  pl.sort_mode = 0;
  pl.bin_shift = 0;
  {
    uint32_t nb_log = 0;
    while ((1ull << nb_log) < g.NB) nb_log++;
    const uint32_t shift = nb_log > 10 ? (nb_log - 10 < 11 ? 11 : nb_log - 10) : 11;
    const bool fits = shift <= 13 && (((uint64_t)g.NB + (1ull << shift) - 1) >> shift) <= 1024;
    bool want = E_sub >= (1ull << 22);  // measured break-even against the single-level sort: ~2^21 digits
    if (const char* env = getenv("MSM_B200_SORT")) want = strcmp(env, "binned") == 0;
    if (want && fits && !pl.partition) {
      pl.sort_mode = 2;
      pl.bin_shift = shift;
    }
  }
  if (pl.sort_mode == 2) b += 2 * Arena::padded((E_sub + g.W) * 4) + 4 * Arena::padded(1025 * 4);
  if (pl.partition) b += (2 * Arena::padded((E_sub + g.W) * 4) + Arena::padded(4096 * 4));  // (bucket, entry) pairs, bin cursors
  b += Arena::padded((size_t)g.NB * n_lines * sizeof(Xyzz<F>));              // bucket accumulators (shared by the sub-batches)
  b += Arena::padded((size_t)2 * pl.slices_cap * n_lines * sizeof(Xyzz<F>));  // slice partials
  b += 2 * Arena::padded((size_t)pl.n_tasks * pl.W_sets * pl.PG * sizeof(Xyzz<F>));  // group partials (ping-pong)
  {
    // cut-bucket, heavy-bucket and heavy-chunk work lists (FixupLists in enqueue_msm).  Every slice
    // appends at most one cut bucket; a heavy bucket spans > HEAVY_SPAN slices and spans overlap by
    // at most one slot, so there are at most slices/HEAVY_SPAN + 1 of them and at most
    // (slices + heavy)/HEAVY_CHUNK + heavy chunk items.
    const size_t heavy_cap = pl.slices_cap / HEAVY_SPAN + 1, chunk_cap = 2 * heavy_cap + pl.slices_cap / HEAVY_CHUNK + 2;
    b += Arena::padded((size_t)3 * n_lines * 4) + Arena::padded((size_t)n_lines * pl.slices_cap * 4) +
         2 * Arena::padded(n_lines * heavy_cap * 4) + Arena::padded(n_lines * chunk_cap * 4) +
         Arena::padded(n_lines * chunk_cap * sizeof(Xyzz<F>));
  }
  // Affine halving rounds (bucket_affine.cuh).  MEASURED SLOWER than the XYZZ slice kernel on B200 (2^24 points:
  // 32.6 ms against 30.5 ms of accumulation; per addition 189 ps in the gather round and 127 - 159 ps in the
  // plane rounds against 150 ps for XYZZ; DESIGN.md section 3b, profiles/r02_affine_rounds.md), so they are OFF
  // unless asked for: MSM_B200_BA=1 applies the heuristic below (rounds while every resident thread still gets
  // a batch that amortises the block's inversion), MSM_B200_BA_ROUNDS=k forces k rounds.
  pl.ba_rounds = 0;
  if (ba_supported<F>() && n_lines == 1 && pl.E_max < (1ull << 30)) {
    uint32_t bps = F::N <= 8 ? 4 : 2;  // resident blocks per SM (launch bounds of k_affine_round)
    if (const char* env = getenv("MSM_B200_BA_BPS")) bps = (uint32_t)atoi(env);
    if (F::N > 8) bps = bps >= 3 ? 3 : 2;
    else bps = bps >= 4 ? 4 : 3;
    pl.ba_bps = bps;
    pl.ba_threads = 148u * bps * BA_BLOCK;
    pl.ba_batch = 1024;
    if (const char* env = getenv("MSM_B200_BA_BATCH")) pl.ba_batch = (uint32_t)std::max(16, atoi(env));
    uint32_t min_per_thread = 96;
    if (const char* env = getenv("MSM_B200_BA_MIN_BATCH")) min_per_thread = (uint32_t)std::max(1, atoi(env));
    int force = 0;
    if (const char* env = getenv("MSM_B200_BA")) {
      if (atoi(env) != 0) force = -1;
    }
    if (const char* env = getenv("MSM_B200_BA_ROUNDS")) force = atoi(env);
    uint64_t eb = E_sub + g.W;  // entries of one sub-batch, upper bound
    uint32_t r = 0;
    while (r < 8) {
      const bool want = force >= 0 ? (int)r < force
                                   : (eb / 2 >= (uint64_t)pl.ba_threads * min_per_thread && eb >= 3ull * g.NB);
      if (!want) break;
      eb = (eb + g.NB) / 2 + 1;
      if (r < 2) pl.ba_cap[r] = eb;
      else pl.ba_cap[r & 1] = std::max(pl.ba_cap[r & 1], eb);
      r++;
    }
    pl.ba_rounds = r;
    if (r) {
      uint32_t St = (uint32_t)(eb / (148ull * 512 * 8));
      St = St < 8 ? 8 : (St > 1024 ? 1024 : St);
      if ((eb + St - 1) / St > pl.slices_cap) St = (uint32_t)((eb + pl.slices_cap - 1) / pl.slices_cap);
      pl.S_tail = St;
      pl.n_slices_tail = (uint32_t)((eb + St - 1) / St);
      b += Arena::padded(pl.ba_cap[0] * sizeof(PackedAffine<F>)) + Arena::padded(pl.ba_cap[1] * sizeof(PackedAffine<F>));
      b += 2 * Arena::padded((size_t)(g.NB + 1) * 4);                                   // offsets, ping-pong
      b += Arena::padded((size_t)pl.ba_threads * pl.ba_batch * F::N * 4);               // prefix products
      b += Arena::padded((size_t)pl.ba_threads * pl.ba_batch * 4);                      // input index | kind
    }
  }
  pl.scratch_bytes = b;
  return MSM_OK;
}

// Enqueue one whole MSM batch on dc.stream.  d_scalars / d_out are device pointers.  No sync.
// pl.n_sub > 1 (single-task calls only): the scalar row is processed in n_sub contiguous
// sub-batches, each sorted and accumulated as soon as sub_ready[k] has fired (its scalars have
// arrived from the host); sub-batch k > 0 continues the buckets sub-batch k-1 left (carry_in), so
// the per-bucket work of the reduction is done once.
template <class F>
int enqueue_msm(msm_ctx* ctx, DeviceCtx& dc, const Plan& pl, const PackedAffine<F>* d_bases,
                uint32_t line_stride, const uint32_t* d_scalars, ApiJacobian<F>* d_out, int timed,
                cudaEvent_t* sub_ready = nullptr) {
  // timed: 0 no events; 1 all phase events; 2 only the closing ones (a later task group of one call: the call's
  // start and first-sort events stay where the first group recorded them)
  const Geometry& g = pl.geo;
  CU_TRY(ctx, dc.arena.ensure(pl.scratch_bytes));
  const uint32_t n_tiles = (g.NB + SCAN_TILE - 1) / SCAN_TILE;
  const uint32_t n_sub = pl.n_sub;
  const uint32_t tb = 128;
  cudaStream_t st = dc.stream;
  Xyzz<F>* bucket_acc = dc.arena.take<Xyzz<F>>((size_t)g.NB * pl.n_lines);
  Xyzz<F>* group_partials = dc.arena.take<Xyzz<F>>((size_t)pl.n_tasks * pl.W_sets * pl.PG);
  Xyzz<F>* group_partials2 = dc.arena.take<Xyzz<F>>((size_t)pl.n_tasks * pl.W_sets * pl.PG);

  // Sort outputs: two sets when the sub-batches of a pipelined call overlap -- sub-batch k+1 is sorted on the sort
  // stream while sub-batch k is accumulated from the other set (MSM_B200_SORT_OVERLAP=0: one stream, one set).
  // Only where the bucket kernel keeps 4 blocks on an SM (BN254 G1): a sort block takes the place of two of them, and
  // the two that remain still keep the multiplier pipe busy.  BLS12-381 has 3 (168 registers): one left per scheduler
  // is not enough -- its 2^22-point row split in two takes 22.1 ms on one stream and 23.2 ms overlapped (job r2_run39).
  bool overlap = n_sub > 1 && dc.sort_stream != nullptr && pl.acc_blocks_per_sm >= 4;
  if (const char* env = getenv("MSM_B200_SORT_OVERLAP")) overlap = n_sub > 1 && dc.sort_stream != nullptr && atoi(env) != 0;
  uint32_t *counts_p[2], *bucket_start_p[2], *cursor_p[2], *tile_sums_p[2], *entries_p[2];
  for (int p = 0; p < (n_sub > 1 ? 2 : 1); p++) {
    counts_p[p] = dc.arena.take<uint32_t>(g.NB + 1);
    bucket_start_p[p] = dc.arena.take<uint32_t>(g.NB + 1);
    cursor_p[p] = dc.arena.take<uint32_t>(g.NB + 1);
    tile_sums_p[p] = dc.arena.take<uint32_t>(n_tiles + 1);
    entries_p[p] = dc.arena.take<uint32_t>((size_t)pl.sub_max * g.W + g.W);
  }
  cudaStream_t sort_st = overlap ? dc.sort_stream : st;
  if (overlap) {
    CU_TRY(ctx, cudaEventRecord(dc.ev_fork, st));
    CU_TRY(ctx, cudaStreamWaitEvent(sort_st, dc.ev_fork, 0));
  }

  if (timed == 1) CU_TRY(ctx, cudaEventRecord(dc.ev[1], st));
  const size_t arena_mark = dc.arena.off;
  for (uint32_t sb = 0; sb < n_sub; sb++) {
    dc.arena.off = arena_mark;  // sub-batches are stream-ordered: they reuse the same scratch
    Geometry sg = g;
    uint32_t S = pl.S, n_slices = pl.n_slices;
    uint64_t E_max = pl.E_max;
    const uint32_t first = pl.sub_first[sb];
    uint32_t bucket0 = 0;  // first bucket of this sub-batch in the call's bucket array (task groups)
    if (n_sub > 1) {
      sg.L = pl.sub_first[sb + 1] - first;
      if (pl.by_task) {
        sg.num_chunks = sg.L / g.chunk_len;
        sg.NB = sg.num_chunks * pl.W_sets * g.B;
        bucket0 = (first / g.chunk_len) * pl.W_sets * g.B;
      } else {
        sg.chunk_len = sg.L ? sg.L : 1;
      }
      sg.point_offset = g.fold ? first : 0;
      E_max = (uint64_t)sg.L * g.W;
      if (pl.waves && sg.L < pl.sub_max) {
        // a shorter sub-batch gets its own whole number of waves (rounded down, at least one): with the longest
        // one's slice length, 3/13 of the row next to 9/13 is 2.67 waves -- a third of a wave on an empty machine
        uint64_t w = (uint64_t)pl.waves * sg.L / pl.sub_max;
        if (w < 1) w = 1;
        const uint64_t cap = (uint64_t)pl.wave_slices * w;
        const uint64_t s_own = (E_max + cap - 1) / cap;
        S = (uint32_t)(s_own < 8 ? 8 : (s_own > 1024 ? 1024 : s_own));
      }
      n_slices = (uint32_t)((E_max + S - 1) / S);
      if (n_slices == 0) n_slices = 1;
      if (n_slices > pl.slices_cap) {  // cannot happen (w <= waves); the plan's slice length always fits
        S = pl.S;
        n_slices = (uint32_t)((E_max + S - 1) / S);
      }
    }
    const int par = overlap ? (int)(sb & 1) : 0;
    uint32_t *counts = counts_p[par], *bucket_start = bucket_start_p[par], *cursor = cursor_p[par];
    uint32_t *tile_sums = tile_sums_p[par], *entries = entries_p[par];
    Xyzz<F>* partials = dc.arena.take<Xyzz<F>>((size_t)2 * pl.slices_cap * pl.n_lines);
    FixupLists fl;
    fl.n_lines = pl.n_lines;
    fl.cut_cap = pl.slices_cap;
    fl.heavy_cap = pl.slices_cap / HEAVY_SPAN + 1;
    fl.chunk_cap = 2 * fl.heavy_cap + pl.slices_cap / HEAVY_CHUNK + 2;
    fl.counts = dc.arena.take<uint32_t>((size_t)3 * pl.n_lines);
    fl.cut_list = dc.arena.take<uint32_t>((size_t)pl.n_lines * fl.cut_cap);
    fl.heavy_list = dc.arena.take<uint32_t>((size_t)pl.n_lines * fl.heavy_cap);
    fl.heavy_chunk0 = dc.arena.take<uint32_t>((size_t)pl.n_lines * fl.heavy_cap);
    fl.chunk_list = dc.arena.take<uint32_t>((size_t)pl.n_lines * fl.chunk_cap);
    Xyzz<F>* chunk_out = dc.arena.take<Xyzz<F>>((size_t)pl.n_lines * fl.chunk_cap);
    // sub-batches of one MSM continue the same buckets (carry_in); task groups own disjoint bucket ranges
    Xyzz<F>* acc_sb = bucket_acc + bucket0;
    const uint32_t carry_in = (sb > 0 && !pl.by_task) ? 1u : 0u;
    const uint32_t* sc_sb = d_scalars + (size_t)first * 8;
    const PackedAffine<F>* bases_sb = (n_sub > 1 && !g.fold) ? d_bases + first : d_bases;

    if (sub_ready) CU_TRY(ctx, cudaStreamWaitEvent(sort_st, sub_ready[sb], 0));
    // the set this sort writes was last read by the bucket kernels of sub-batch sb - 2
    if (overlap && sb >= 2) CU_TRY(ctx, cudaStreamWaitEvent(sort_st, dc.ev_acc[par], 0));
    // --- sort: bucket_start[NB+1] and the entries in bucket order
    const SortBuffers sbuf{counts, bucket_start, cursor, tile_sums, entries};
    int src = pl.sort_mode == 2 ? enqueue_sort_binned(ctx, dc, pl, sg, E_max, sc_sb, sbuf, sort_st)
              : pl.partition   ? enqueue_sort_partition(ctx, dc, pl, sg, E_max, sc_sb, sbuf, sort_st)
                               : enqueue_sort_atomic(ctx, dc, pl, sg, E_max, sc_sb, sbuf, sort_st);
    if (src != MSM_OK) return src;
    if (overlap) {
      CU_TRY(ctx, cudaEventRecord(dc.ev_sorted[par], sort_st));
      CU_TRY(ctx, cudaStreamWaitEvent(st, dc.ev_sorted[par], 0));
    }
    if (timed == 1 && sb == 0) CU_TRY(ctx, cudaEventRecord(dc.ev[2], st));
    if (aborted(ctx)) return MSM_ERR_ABORTED;
    // --- accumulate
    dim3 grid((n_slices + tb - 1) / tb, pl.n_lines);
    CU_TRY(ctx, cudaMemsetAsync(fl.counts, 0, (size_t)3 * pl.n_lines * 4, st));
    if (!carry_in) CU_TRY(ctx, cudaMemsetAsync(acc_sb, 0, (size_t)sg.NB * pl.n_lines * sizeof(Xyzz<F>), st));  // all infinity
    const uint32_t* acc_starts = bucket_start;  // bucket offsets of what the XYZZ kernel walks
    bool ba_done = false;
    if constexpr (ba_supported<F>()) {
      if (pl.ba_rounds) {
        // affine halving rounds (bucket_affine.cuh): entries -> planes[0] -> planes[1] -> planes[0] ...
        constexpr size_t NW = F::N;
        uint32_t* planes[2] = {dc.arena.take<uint32_t>(2 * pl.ba_cap[0] * NW), dc.arena.take<uint32_t>(2 * pl.ba_cap[1] * NW)};
        uint32_t* offs[2] = {dc.arena.take<uint32_t>(sg.NB + 1), dc.arena.take<uint32_t>(sg.NB + 1)};
        uint32_t* sc_prefix = dc.arena.take<uint32_t>((size_t)pl.ba_threads * pl.ba_batch * F::N);
        uint32_t* sc_idx = dc.arena.take<uint32_t>((size_t)pl.ba_threads * pl.ba_batch);
        const uint32_t* off_in = bucket_start;
        const uint32_t ba_grid = pl.ba_threads / BA_BLOCK;
        BaPoints<F> pin{bases_sb, entries, nullptr, nullptr};
        for (uint32_t r = 0; r < pl.ba_rounds; r++) {
          uint32_t* off_out = offs[r & 1];
          const SortBuffers hb{counts, off_out, cursor, tile_sums, nullptr};
          enqueue_halve_scan(st, sg, off_in, hb);
          BaPoints<F> pout{nullptr, nullptr, planes[r & 1], planes[r & 1] + pl.ba_cap[r & 1] * NW};
          constexpr int BPS_LO = F::N <= 8 ? 3 : 2, BPS_HI = BPS_LO + 1;
          if (r == 0 && pl.ba_bps == BPS_HI)
            k_affine_round<F, true, BPS_HI><<<ba_grid, BA_BLOCK, 0, st>>>(pin, off_in, off_out, sg.NB, pl.ba_batch, pout, sc_prefix, sc_idx);
          else if (r == 0)
            k_affine_round<F, true, BPS_LO><<<ba_grid, BA_BLOCK, 0, st>>>(pin, off_in, off_out, sg.NB, pl.ba_batch, pout, sc_prefix, sc_idx);
          else if (pl.ba_bps == BPS_HI)
            k_affine_round<F, false, BPS_HI><<<ba_grid, BA_BLOCK, 0, st>>>(pin, off_in, off_out, sg.NB, pl.ba_batch, pout, sc_prefix, sc_idx);
          else
            k_affine_round<F, false, BPS_LO><<<ba_grid, BA_BLOCK, 0, st>>>(pin, off_in, off_out, sg.NB, pl.ba_batch, pout, sc_prefix, sc_idx);
          pin = pout;
          off_in = off_out;
          dc.launches += 5;
        }
        acc_starts = off_in;
        S = pl.S_tail;
        n_slices = pl.n_slices_tail;
        grid = dim3((n_slices + tb - 1) / tb, 1);
        k_accumulate<F><<<grid, tb, 0, st>>>(nullptr, 0u, nullptr, acc_starts, sg.NB, acc_starts + sg.NB, S, n_slices, acc_sb,
                                             partials, carry_in, fl.counts, fl.cut_list, fl.cut_cap, pin.x, pin.y);
        ba_done = true;
      }
    }
    if (!ba_done)
      k_accumulate<F><<<grid, tb, 0, st>>>(bases_sb, line_stride, entries, bucket_start, sg.NB, bucket_start + sg.NB,
                                           S, n_slices, acc_sb, partials, carry_in, fl.counts, fl.cut_list, fl.cut_cap,
                                           nullptr, nullptr);
    // at most one cut bucket per slice
    k_fixup_cut<F><<<grid, tb, 0, st>>>(acc_starts, sg.NB, S, n_slices, acc_sb, partials, fl);
    const uint32_t hblocks = (fl.chunk_cap + 3) / 4;
    dim3 hgrid(hblocks < 148 * 8 ? hblocks : 148 * 8, pl.n_lines);
    k_fixup_heavy<F><<<hgrid, tb, tb * sizeof(Xyzz<F>), st>>>(acc_starts, sg.NB, S, n_slices, acc_sb, partials, fl,
                                                             chunk_out);
    dim3 fgrid2(hblocks < 148 ? hblocks : 148, pl.n_lines);
    k_fixup_heavy_final<F><<<fgrid2, tb, tb * sizeof(Xyzz<F>), st>>>(acc_starts, sg.NB, S, acc_sb, fl, chunk_out);
    if (overlap) CU_TRY(ctx, cudaEventRecord(dc.ev_acc[par], st));
    dc.launches += 9;
  }
  if (timed) CU_TRY(ctx, cudaEventRecord(dc.ev[3], st));
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  // --- reduce + combine
  {
    const uint64_t n_threads = (uint64_t)g.NB * pl.n_lines / pl.Q;
    k_bucket_reduce<F><<<(uint32_t)((n_threads + tb - 1) / tb), tb, tb * sizeof(Xyzz<F>), st>>>(
        bucket_acc, (uint32_t)n_threads, g.B, pl.Q, pl.RW, group_partials);
    const uint32_t Ws = pl.W_sets;
    // the PG partials of every (task, window) group shrink 1024-fold per pass
    uint32_t pg = pl.PG;
    Xyzz<F>*src = group_partials, *dst = group_partials2;
    while (pg > 8) {
      const uint32_t out_pg = (pg + 1023) / 1024;
      dim3 rgrid(out_pg, pl.n_tasks * Ws);
      k_reduce_points<F><<<rgrid, tb, tb * sizeof(Xyzz<F>), st>>>(src, pg, out_pg, dst);
      dc.launches += 1;
      std::swap(src, dst);
      pg = out_pg;
    }
    if (pl.n_tasks >= 512) {
      k_window_combine_batched<F><<<(pl.n_tasks + tb - 1) / tb, tb, 0, st>>>(src, pl.n_tasks, Ws, pg, g.c, d_out);
    } else {
      const uint32_t wt = Ws < 32 ? 32 : (Ws > 256 ? 256 : ((Ws + 31) / 32) * 32);
      k_window_combine<F><<<pl.n_tasks, wt, (size_t)Ws * sizeof(Xyzz<F>), st>>>(src, Ws, pg, g.c, d_out);
    }
  }
  if (timed) CU_TRY(ctx, cudaEventRecord(dc.ev[4], st));
  dc.launches += 2;
  CU_TRY(ctx, cudaGetLastError());
  return MSM_OK;
}

inline void collect_timings(msm_ctx* ctx, DeviceCtx& dc, const Plan& pl, bool have_h2d) {
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
  t.scatter_passes = pl.scatter_passes;
  t.sub_batches = pl.n_sub;
}

template <class F> int build_table_impl(msm_ctx* ctx, msm_bases::Shard& sh, uint32_t c, size_t chunk_len, bool budgeted = false);

template <class F>
int multiple_multiexp_impl(msm_ctx* ctx, const msm_bases* bases, const void* scalars, size_t L,
                           uint32_t num_chunks, void* out, bool device_io) {
  if (bases->shards.size() != 1 || bases->shards[0].dev_idx != 0) {
    set_error(ctx, "multiple_multiexp needs bases resident on device 0 (msm_bases_upload)");
    return MSM_ERR_INVALID;
  }
  if (L == 0 || L > bases->n || L >= (1ull << 31)) {
    set_error(ctx, "multiple_multiexp: exponent count must be in 1..=number of bases");
    return MSM_ERR_INVALID;
  }
  const uint32_t n_lines = (uint32_t)(bases->n / L);  // ag-cuda-ec/src/multiexp.rs:28-30
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  // Lazy window table (include/msm_b200.h, "Window tables by policy"): the second call of one shape
  // builds the table that shape wants, if the cost model wants one and it fits the memory budget.
  {
    msm_bases* mb = const_cast<msm_bases*>(bases);
    if (mb->shape_L == L && mb->shape_chunks == num_chunks) {
      mb->shape_calls++;
    } else {
      mb->shape_L = L;
      mb->shape_chunks = num_chunks;
      mb->shape_calls = 1;
      mb->table_failed = false;
      mb->shape_best_ms = 0.f;
    }
    msm_bases::Shard& shm = mb->shards[0];
    const bool fits_shape = shm.table && mb->table_L == L && mb->table_chunks == num_chunks;
    if (mb->table_policy != MSM_TABLE_OFF && !mb->table_explicit && !fits_shape && !mb->table_failed &&
        mb->shape_calls >= 2 && !ctx->window_override && !getenv("MSM_B200_WINDOW") && num_chunks >= 1 && num_chunks <= L) {
      const bool whole = num_chunks == 1 && n_lines == 1 && L == shm.n;
      const uint32_t chunk_len = (uint32_t)(L / num_chunks);
      const uint32_t bits = scalar_bits(ctx->curve);
      bool want = whole;
      if (!whole && (size_t)L * n_lines == shm.n) {
        const uint64_t groups = (uint64_t)num_chunks * n_lines;
        const uint32_t tc = choose_table_window(chunk_len, bits, groups, shm.n);
        double plain_cost = 0;
        choose_window(chunk_len, bits, groups, sizeof(Xyzz<F>), &plain_cost);
        want = tc != 0 && fold_cost(tc, chunk_len, bits, groups) < plain_cost;
      }
      if (want) {
        int trc = build_table_impl<F>(ctx, shm, 0, whole ? 0 : chunk_len, /*budgeted=*/true);
        if (trc == MSM_OK) {
          mb->table_L = L;
          mb->table_chunks = num_chunks;
        } else {
          mb->table_failed = true;  // over budget or out of memory: the plain resident copy keeps serving
          cudaGetLastError();
        }
      } else {
        mb->table_failed = true;
      }
    }
  }
  const msm_bases::Shard& sh0 = bases->shards[0];
  // A window table serves (i) the call it was built for -- one MSM over the whole shard -- and (ii) any
  // chunked / multi-line call for which folding all windows of a task into one bucket set is cheaper
  // than the plain per-window bucket sets at the window size the plain path would pick.
  bool use_table = false;
  if (sh0.table && !ctx->window_override && !getenv("MSM_B200_WINDOW") && num_chunks >= 1 && num_chunks <= L) {
    if (num_chunks == 1 && n_lines == 1 && L == sh0.n) {
      use_table = true;
    } else if (!getenv("MSM_B200_NO_CHUNK_TABLE")) {
      const uint32_t chunk_len = (uint32_t)(L / num_chunks);
      const uint64_t groups = (uint64_t)num_chunks * n_lines;
      const double bucket_bytes = (double)groups * (double)(1u << (sh0.table_c - 1)) * (double)sizeof(Xyzz<F>);
      double plain_cost = 0;
      choose_window(chunk_len, scalar_bits(ctx->curve), groups, sizeof(Xyzz<F>), &plain_cost);
      use_table = bucket_bytes <= 12e9 && (uint64_t)num_chunks << (sh0.table_c - 1) < (1ull << 31) &&
                  fold_cost(sh0.table_c, chunk_len, scalar_bits(ctx->curve), groups) < plain_cost;
    }
  }
  // Host scalars of a large call arrive in sub-batches on a copy stream while the previous sub-batch is already being
  // sorted and accumulated (pipeline_shape above; the 32 B/scalar upload is ~20 % of the call otherwise).
  uint32_t n_sub = 1;
  double sub_ratio = 2.0;
  if (!device_io) {
    const msm_bases* mb = bases;
    const bool measured = mb->shape_best_table == use_table && !getenv("MSM_B200_PIPELINE_STATIC");
    pipeline_shape(L, num_chunks, n_lines, measured ? dc.h2d_gbs : 0.f, measured ? mb->shape_best_ms : 0.f, &n_sub, &sub_ratio);
  }
  if (const char* env = getenv("MSM_B200_PIPELINE")) {
    const int v = atoi(env);
    // MSM_B200_PIPELINE_DEVICE: also split device-resident rows (measurement of the split's own cost)
    if (v >= 1 && v <= 8 && n_lines == 1) {
      if (num_chunks == 1 && (!device_io || getenv("MSM_B200_PIPELINE_DEVICE"))) n_sub = (uint32_t)v;
      if (num_chunks > 1 && !device_io && num_chunks >= 2 * (uint32_t)v) n_sub = (uint32_t)v;
    }
  }
  Plan pl;
  int rc = make_plan<F>(ctx, (uint32_t)L, n_lines, num_chunks, pl, use_table ? sh0.table_c : 0, n_sub,
                        use_table ? (uint32_t)sh0.n : 0, sub_ratio);
  if (rc) return rc;
  if (aborted(ctx)) return MSM_ERR_ABORTED;
  const uint32_t* d_scalars;
  ApiJacobian<F>* d_out;
  const size_t out_bytes = (size_t)pl.n_tasks * sizeof(ApiJacobian<F>);
  // Once the pipelined sub-batch copies are queued they read the caller's buffer asynchronously (truly so
  // for pinned memory): every exit before the final synchronise -- error or abort -- drains both streams
  // first, so the caller may free / unregister the buffer as soon as the call returns.
  struct Drain {
    cudaStream_t a = nullptr, b = nullptr, c = nullptr;
    ~Drain() {
      if (a) cudaStreamSynchronize(a);
      if (c) cudaStreamSynchronize(c);
      if (b) cudaStreamSynchronize(b);
    }
  } drain;
  if (device_io) {
    d_scalars = static_cast<const uint32_t*>(scalars);
    d_out = static_cast<ApiJacobian<F>*>(out);
  } else {
    CU_TRY(ctx, dc.io.ensure(Arena::padded(L * 32) + Arena::padded(out_bytes)));
    uint32_t* ds = dc.io.take<uint32_t>(L * 8);
    d_out = dc.io.take<ApiJacobian<F>>(pl.n_tasks);
    CU_TRY(ctx, cudaEventRecord(dc.ev[0], dc.stream));
    if (pl.n_sub > 1) {
      // the copy stream must not overwrite the staging area while an earlier call still reads it
      CU_TRY(ctx, cudaStreamWaitEvent(dc.copy_stream, dc.ev[0], 0));
      drain.a = dc.copy_stream;
      drain.b = dc.stream;
      drain.c = dc.sort_stream;
      CU_TRY(ctx, cudaEventRecord(dc.ev_h2d[0], dc.copy_stream));
      for (uint32_t sb = 0; sb < pl.n_sub; sb++) {
        const size_t first = pl.sub_first[sb], cnt = pl.sub_first[sb + 1] - first;
        if (cnt)
          CU_TRY(ctx, cudaMemcpyAsync(ds + first * 8, static_cast<const char*>(scalars) + first * 32, cnt * 32,
                                      cudaMemcpyHostToDevice, dc.copy_stream));
        CU_TRY(ctx, cudaEventRecord(dc.ev_copy[sb], dc.copy_stream));
      }
      CU_TRY(ctx, cudaEventRecord(dc.ev_h2d[1], dc.copy_stream));
    } else {
      CU_TRY(ctx, cudaMemcpyAsync(ds, scalars, L * 32, cudaMemcpyHostToDevice, dc.stream));
    }
    d_scalars = ds;
  }
  rc = enqueue_msm<F>(ctx, dc, pl, static_cast<const PackedAffine<F>*>(use_table ? sh0.table : sh0.ptr), (uint32_t)L,
                      d_scalars, d_out, true, (!device_io && pl.n_sub > 1) ? dc.ev_copy : nullptr);
  if (rc) {
    if (dc.sort_stream) cudaStreamSynchronize(dc.sort_stream);
    cudaStreamSynchronize(dc.stream);
    return rc;
  }
  if (!device_io) CU_TRY(ctx, cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, dc.stream));
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  drain.a = drain.b = drain.c = nullptr;  // every sub-batch copy was waited for by the kernels that just finished
  collect_timings(ctx, dc, pl, !device_io);
  {
    // what the next call of this shape plans its sub-batches with
    msm_bases* mb = const_cast<msm_bases*>(bases);
    const float total = ctx->tm.total_ms;
    if (total > 0.f && (mb->shape_best_ms <= 0.f || mb->shape_best_table != use_table || total < mb->shape_best_ms)) {
      mb->shape_best_ms = total;
      mb->shape_best_table = use_table;
    }
    if (!device_io && L >= (1u << 20)) {
      float ms = 0.f;
      if (pl.n_sub > 1) {
        // the copy stream recorded this event right after the last sub-batch's copy, which the kernels that just
        // finished had waited for: the wait below returns at once, and the elapsed-time query cannot come back
        // "not ready" (an error code the next cudaGetLastError would otherwise pick up)
        if (cudaEventSynchronize(dc.ev_h2d[1]) != cudaSuccess || cudaEventElapsedTime(&ms, dc.ev_h2d[0], dc.ev_h2d[1]) != cudaSuccess) {
          ms = 0.f;
          cudaGetLastError();
        }
      } else {
        ms = ctx->tm.h2d_ms;
      }
      if (ms > 0.f) dc.h2d_gbs = (float)((double)L * 32.0 / ((double)ms * 1e6));
    }
  }
  return MSM_OK;
}

template <class F>
int convert_bases_impl(msm_ctx* ctx, DeviceCtx& dc, const void* d_api, size_t n, void* d_packed) {
  if (n == 0) return MSM_OK;
  k_convert_bases<F><<<(uint32_t)((n + 127) / 128), 128, 0, dc.stream>>>(
      static_cast<const ApiAffine<F>*>(d_api), (uint32_t)n, static_cast<PackedAffine<F>*>(d_packed));
  dc.launches += 1;
  CU_TRY(ctx, cudaGetLastError());
  return MSM_OK;
}

// One MSM split over all devices of the context (MultiexpKernel::multiexp).  Either host bases
// (uploaded and converted per call, as the reference uploads per call) or resident sharded bases.
template <class F>
int multiexp_impl(msm_ctx* ctx, const void* host_bases, const msm_bases* resident, size_t skip,
                  const void* scalars, size_t n, void* out) {
  const size_t n_dev = ctx->devs.size();
  if (n >= (1ull << 31)) return MSM_ERR_TOO_LARGE;
  if (n == 0) {
    ApiJacobian<F> inf;
    F::to_api(F::zero(), inf.x);
    F::to_api(F::one(), inf.y);
    F::to_api(F::zero(), inf.z);
    memcpy(out, &inf, sizeof(inf));
    return MSM_OK;
  }
  struct Job {
    size_t dev_idx;
    const void* d_bases;  // resident (packed) device pointer, or the shard's window table
    const char* h_bases;  // host pointer (API layout) when not resident
    size_t s_off, cnt;
    uint32_t table_c;     // != 0: d_bases is a window table covering exactly cnt points
  };
  std::vector<Job> jobs;
  if (resident) {
    if (skip + n > resident->n) return MSM_ERR_INVALID;
    {
      // lazy window tables, as in multiple_multiexp (msm_b200.h, "Window tables by policy"): the second call
      // over the same range builds the table of every shard the range covers completely
      msm_bases* mb = const_cast<msm_bases*>(resident);
      if (mb->shape_L == n && mb->shape_chunks == (uint32_t)skip) {
        mb->shape_calls++;
      } else {
        mb->shape_L = n;
        mb->shape_chunks = (uint32_t)skip;
        mb->shape_calls = 1;
        mb->table_failed = false;
      }
      if (mb->table_policy != MSM_TABLE_OFF && !mb->table_explicit && !mb->table_failed && mb->shape_calls >= 2 &&
          !ctx->window_override && !getenv("MSM_B200_WINDOW")) {
        for (auto& sh : mb->shards) {
          if (sh.table || sh.n == 0 || sh.start < skip || sh.start + sh.n > skip + n) continue;
          if (build_table_impl<F>(ctx, sh, 0, 0, /*budgeted=*/true) != MSM_OK) {
            mb->table_failed = true;
            cudaGetLastError();
            break;
          }
        }
      }
    }
    for (const auto& sh : resident->shards) {
      const size_t lo = std::max(skip, sh.start), hi = std::min(skip + n, sh.start + sh.n);
      if (lo >= hi) continue;
      const bool whole = sh.table && lo == sh.start && hi == sh.start + sh.n && !ctx->window_override;
      if (whole) jobs.push_back({(size_t)sh.dev_idx, sh.table, nullptr, lo - skip, hi - lo, sh.table_c});
      else jobs.push_back({(size_t)sh.dev_idx,
                           static_cast<const char*>(sh.ptr) + (lo - sh.start) * sizeof(PackedAffine<F>), nullptr,
                           lo - skip, hi - lo, 0u});
    }
  } else {
    const size_t chunk = (n + n_dev - 1) / n_dev;  // ec-gpu-proxy/src/multiexp.rs:329-337
    for (size_t d = 0; d * chunk < n; d++) {
      const size_t cnt = std::min(chunk, n - d * chunk);
      jobs.push_back({d, nullptr, static_cast<const char*>(host_bases) + d * chunk * sizeof(ApiAffine<F>),
                      d * chunk, cnt, 0u});
    }
  }
  std::vector<int> rcs(jobs.size(), MSM_OK);
  std::vector<std::string> errs(jobs.size());
  std::vector<Plan> plans(jobs.size());
  std::vector<ApiJacobian<F>*> d_partials(jobs.size(), nullptr);
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
    // host scalars of a large shard arrive in pipelined sub-batches, as in multiple_multiexp
    const uint32_t n_sub = job.cnt >= (1u << 23) ? 4 : (job.cnt >= (1u << 20) ? 2 : 1);
    int rc = make_plan<F>(&shadow, (uint32_t)job.cnt, 1, 1, plans[j], job.table_c, n_sub);
    if (rc) {
      rcs[j] = rc;
      return;
    }
    const size_t api_bytes = job.h_bases ? job.cnt * sizeof(ApiAffine<F>) : 0;
    const size_t packed_bytes = job.h_bases ? job.cnt * sizeof(PackedAffine<F>) : 0;
    e = dc.io.ensure(Arena::padded(job.cnt * 32) + Arena::padded(sizeof(ApiJacobian<F>)) +
                     Arena::padded(api_bytes) + Arena::padded(packed_bytes));
    if (e != cudaSuccess) return fail(e, "io arena");
    uint32_t* ds = dc.io.take<uint32_t>(job.cnt * 8);
    ApiJacobian<F>* d_out = dc.io.take<ApiJacobian<F>>(1);
    const PackedAffine<F>* d_bases = static_cast<const PackedAffine<F>*>(job.d_bases);
    cudaEventRecord(dc.ev[0], dc.stream);
    if (job.h_bases) {
      ApiAffine<F>* da = dc.io.take<ApiAffine<F>>(job.cnt);
      PackedAffine<F>* dp = dc.io.take<PackedAffine<F>>(job.cnt);
      e = cudaMemcpyAsync(da, job.h_bases, api_bytes, cudaMemcpyHostToDevice, dc.stream);
      if (e != cudaSuccess) return fail(e, "bases H2D");
      rc = convert_bases_impl<F>(&shadow, dc, da, job.cnt, dp);
      if (rc) {
        rcs[j] = rc;
        errs[j] = shadow.err;
        return;
      }
      d_bases = dp;
    }
    const Plan& pl = plans[j];
    const char* h_sc = static_cast<const char*>(scalars) + job.s_off * 32;
    if (pl.n_sub > 1) {
      cudaStreamWaitEvent(dc.copy_stream, dc.ev[0], 0);
      for (uint32_t sb = 0; sb < pl.n_sub && e == cudaSuccess; sb++) {
        const size_t first = pl.sub_first[sb], cnt = pl.sub_first[sb + 1] - first;
        if (cnt) e = cudaMemcpyAsync(ds + first * 8, h_sc + first * 32, cnt * 32, cudaMemcpyHostToDevice, dc.copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(dc.ev_copy[sb], dc.copy_stream);
      }
    } else {
      e = cudaMemcpyAsync(ds, h_sc, job.cnt * 32, cudaMemcpyHostToDevice, dc.stream);
    }
    if (e != cudaSuccess) {
      cudaStreamSynchronize(dc.copy_stream);
      return fail(e, "scalars H2D");
    }
    rc = enqueue_msm<F>(&shadow, dc, pl, d_bases, (uint32_t)job.cnt, ds, d_out, true, pl.n_sub > 1 ? dc.ev_copy : nullptr);
    if (rc) {
      rcs[j] = rc;
      errs[j] = shadow.err;
      cudaStreamSynchronize(dc.copy_stream);
      if (dc.sort_stream) cudaStreamSynchronize(dc.sort_stream);
      cudaStreamSynchronize(dc.stream);
      return;
    }
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
      for (auto& job : jobs) {  // nothing may still read the caller's buffers when the call returns
        cudaSetDevice(ctx->devs[job.dev_idx].dev);
        cudaStreamSynchronize(ctx->devs[job.dev_idx].copy_stream);
        if (ctx->devs[job.dev_idx].sort_stream) cudaStreamSynchronize(ctx->devs[job.dev_idx].sort_stream);
        cudaStreamSynchronize(ctx->devs[job.dev_idx].stream);
      }
      set_error(ctx, errs[j]);
      return rcs[j];
    }
  }
  // gather the per-device partial points on device 0 over peer copies, sum there: the on-device
  // replacement of the host loop at ec-gpu-proxy/src/multiexp.rs:394-397
  DeviceCtx& d0 = ctx->devs[0];
  for (size_t j = 0; j < jobs.size(); j++) {
    DeviceCtx& dc = ctx->devs[jobs[j].dev_idx];
    CU_TRY(ctx, cudaSetDevice(dc.dev));
    CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  }
  CU_TRY(ctx, cudaSetDevice(d0.dev));
  if (jobs.size() == 1 && jobs[0].dev_idx == 0) {
    CU_TRY(ctx, cudaMemcpyAsync(out, d_partials[0], sizeof(ApiJacobian<F>), cudaMemcpyDeviceToHost, d0.stream));
    CU_TRY(ctx, cudaStreamSynchronize(d0.stream));
  } else {
    ApiJacobian<F>* gather = reinterpret_cast<ApiJacobian<F>*>(d0.small);
    ApiJacobian<F>* d_result = gather + jobs.size();
    for (size_t j = 0; j < jobs.size(); j++) {
      DeviceCtx& dc = ctx->devs[jobs[j].dev_idx];
      if (dc.dev == d0.dev) {
        CU_TRY(ctx, cudaMemcpyAsync(gather + j, d_partials[j], sizeof(ApiJacobian<F>), cudaMemcpyDeviceToDevice, d0.stream));
      } else {
        CU_TRY(ctx, cudaMemcpyPeerAsync(gather + j, d0.dev, d_partials[j], dc.dev, sizeof(ApiJacobian<F>), d0.stream));
      }
    }
    k_sum_points<F><<<1, 32, 0, d0.stream>>>(gather, (uint32_t)jobs.size(), d_result);
    d0.launches += 1;
    CU_TRY(ctx, cudaMemcpyAsync(out, d_result, sizeof(ApiJacobian<F>), cudaMemcpyDeviceToHost, d0.stream));
    CU_TRY(ctx, cudaStreamSynchronize(d0.stream));
  }
  for (size_t j = 0; j < jobs.size(); j++)
    if (jobs[j].dev_idx == 0) collect_timings(ctx, d0, plans[j], true);
  return MSM_OK;
}

// Window table for one resident shard (msm_bases_precompute).
template <class F> int build_table_impl(msm_ctx* ctx, msm_bases::Shard& sh, uint32_t c, size_t chunk_len, bool budgeted) {
  if (sh.n == 0) return MSM_OK;
  DeviceCtx& dc = ctx->devs[sh.dev_idx];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const uint32_t bits = scalar_bits(ctx->curve);
  if (c == 0 && chunk_len != 0 && chunk_len < sh.n) {
    // table for many tasks of chunk_len points each (msm_bases_precompute_chunked)
    c = choose_table_window((uint32_t)chunk_len, bits, sh.n / chunk_len, sh.n);
    if (c == 0) return MSM_ERR_TOO_LARGE;
  }
  if (c == 0) {
    // all windows share one bucket set: cost = W * n mixed adds + 2^(c-1) * 2 full adds
    double best = 1e300;
    for (uint32_t cc = 11; cc <= 24; cc++) {
      const uint32_t W = (bits + 1 + cc - 1) / cc;
      if ((uint64_t)W * sh.n >= (1ull << 31)) continue;
      // per-bucket figure from measurement (2^21 points, c = 19 -> 20: -0.41 ms accumulate, +0.14 ms reduce)
      // (below 2^21 points the slices get short and a saved window buys less: measured 2^20, c = 17 beats 19)
      const double cost = (double)W * (double)sh.n * 10.0 + (double)(1u << (cc - 1)) * (sh.n < (1u << 21) ? 60.0 : 45.0);
      if (cost < best) {
        best = cost;
        c = cc;
      }
    }
    if (c == 0) return MSM_ERR_TOO_LARGE;
  }
  const uint32_t W = (bits + 1 + c - 1) / c;
  if (c < 8 || c > 24 || W > (uint32_t)TABLE_MAX_W || (uint64_t)W * sh.n >= (1ull << 31)) {
    set_error(ctx, "msm_bases_precompute: window size out of range for this shard");
    return MSM_ERR_INVALID;
  }
  if (sh.table) {
    cudaFree(sh.table);
    sh.table = nullptr;
    sh.table_c = sh.table_W = 0;
  }
  const size_t table_bytes = (size_t)W * sh.n * sizeof(PackedAffine<F>);
  if (budgeted) {
    // tables built by policy (nobody asked for them) stay within a stated budget
    size_t free_b = 0, total_b = 0;
    CU_TRY(ctx, cudaMemGetInfo(&free_b, &total_b));
    double budget = 0.5 * (double)total_b;
    if (const char* env = getenv("MSM_B200_TABLE_BUDGET_GB")) budget = atof(env) * 1e9;
    if ((double)table_bytes > budget || (double)table_bytes > 0.8 * (double)free_b) {
      set_error(ctx, "window table over budget");
      return MSM_ERR_TOO_LARGE;
    }
  }
  CU_TRY(ctx, cudaMalloc(&sh.table, table_bytes));
  k_build_tables<F><<<(uint32_t)((sh.n + 63) / 64), 64, 0, dc.stream>>>(
      static_cast<const PackedAffine<F>*>(sh.ptr), (uint32_t)sh.n, c, W, static_cast<PackedAffine<F>*>(sh.table));
  dc.launches += 1;
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  sh.table_c = c;
  sh.table_W = W;
  return MSM_OK;
}

// canonical generator in the API layout (Montgomery), computed with the host build of the field
template <class F> ApiAffine<F> host_generator_api() {
  using BP = BaseParams<F>;
  constexpr int NB = BP::N;
  constexpr int COMPONENTS = F::API_WORDS / NB;  // 1: G1, 2: G2 (c0 | c1)
  uint32_t gx[F::API_WORDS] = {0}, gy[F::API_WORDS] = {0};
  if (is_ext2<F>()) {
    for (int c = 0; c < 2; c++)
      for (int i = 0; i < NB; i++) {
        gx[c * NB + i] = is_bn254<F>() ? Bn254G2Gen::W(c, i < 8 ? i : 0) : Bls381G2Gen::W(c, i < 12 ? i : 0);
        gy[c * NB + i] = is_bn254<F>() ? Bn254G2Gen::W(2 + c, i < 8 ? i : 0) : Bls381G2Gen::W(2 + c, i < 12 ? i : 0);
      }
  } else if (is_bn254<F>()) {  // BN254 G1: (1, 2)
    gx[0] = 1;
    gy[0] = 2;
  } else {  // BLS12-381 G1
    const uint32_t x[12] = {0xdb22c6bbu, 0xfb3af00au, 0xf97a1aefu, 0x6c55e83fu, 0x171bac58u, 0xa14e3a3fu,
                            0x9774b905u, 0xc3688c4fu, 0x4fa9ac0fu, 0x2695638cu, 0x3197d794u, 0x17f1d3a7u};
    const uint32_t y[12] = {0x46c5e7e1u, 0x0caa2329u, 0xa2888ae4u, 0xd03cc744u, 0x2c04b3edu, 0x00db18cbu,
                            0xd5d00af6u, 0xfcf5e095u, 0x741d8ae4u, 0xa09e30edu, 0xe3aaa0f1u, 0x08b3f481u};
    for (int i = 0; i < NB && i < 12; i++) {
      gx[i] = x[i];
      gy[i] = y[i];
    }
  }
  // integer -> API Montgomery form, component by component
  ApiAffine<F> g;
  for (int c = 0; c < COMPONENTS; c++) {
    Fp<BP> ex, ey;
    for (int i = 0; i < NB; i++) {
      ex.v[i] = gx[c * NB + i];
      ey.v[i] = gy[c * NB + i];
    }
    ex = fp_to_mont<BP>(ex);
    ey = fp_to_mont<BP>(ey);
    for (int i = 0; i < NB; i++) {
      g.x[c * NB + i] = ex.v[i];
      g.y[c * NB + i] = ey.v[i];
    }
  }
  return g;
}

template <class F> int synth_points_impl(msm_ctx* ctx, uint64_t seed, size_t start, size_t n, void* d_out) {
  if (n == 0) return MSM_OK;
  if (n >= (1ull << 32)) return MSM_ERR_TOO_LARGE;
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const uint64_t a = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ull, 0);
  const uint64_t b = splitmix64(seed ^ 0xA5A5A5A5A5A5A5A5ull, 1) | 1;
  // D = b*G on the host (64 doublings; setup only)
  const ApiAffine<F> gen_api = host_generator_api<F>();
  const Affine<F> gen = affine_from_api<F>(&gen_api);
  Xyzz<F> acc = xyzz_inf<F>();
  for (int bit = 63; bit >= 0; bit--) {
    acc = xyzz_dbl<F>(acc);
    if ((b >> bit) & 1) xyzz_madd<F>(acc, gen);
  }
  ApiAffine<F> d_api;
  xyzz_to_api_affine<F>(acc, true, &d_api);
  const uint32_t threads = (uint32_t)((n + SYNTH_RUN - 1) / SYNTH_RUN);
  k_synth_points<F><<<(threads + 63) / 64, 64, 0, dc.stream>>>(gen_api, d_api, a, b, (uint64_t)start, (uint32_t)n,
                                                              static_cast<ApiAffine<F>*>(d_out));
  dc.launches += 1;
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

template <class F>
int test_fq_impl(msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const size_t bytes = count * sizeof(ApiElem<F>);
  CU_TRY(ctx, dc.io.ensure(3 * Arena::padded(bytes)));
  ApiElem<F>* da = dc.io.take<ApiElem<F>>(count);
  ApiElem<F>* db = dc.io.take<ApiElem<F>>(count);
  ApiElem<F>* dout = dc.io.take<ApiElem<F>>(count);
  CU_TRY(ctx, cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, dc.stream));
  if (b) CU_TRY(ctx, cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, dc.stream));
  using SatP = BaseParams<F>;
  ApiElem<F> r2;  // R^2 mod p (Fq2: (R^2, 0))
  for (int i = 0; i < F::API_WORDS; i++) r2.w[i] = i < SatP::N ? SatP::R2(i < SatP::N ? i : 0) : 0u;
  k_test_fq<F><<<(uint32_t)((count + 127) / 128), 128, 0, dc.stream>>>(op, da, b ? db : nullptr, r2, dout, (uint32_t)count);
  dc.launches += 1;
  CU_TRY(ctx, cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, dc.stream));
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

template <class F>
int test_ec_impl(msm_ctx* ctx, int op, const void* a, const void* b, void* out, size_t count) {
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const size_t ja = count * sizeof(ApiJacobian<F>);
  const size_t bb = op == 0 ? ja : (op == 1 ? count * sizeof(ApiAffine<F>) : 0);
  CU_TRY(ctx, dc.io.ensure(3 * Arena::padded(ja)));
  ApiJacobian<F>* da = dc.io.take<ApiJacobian<F>>(count);
  void* db = dc.io.take<ApiJacobian<F>>(count);
  ApiJacobian<F>* dout = dc.io.take<ApiJacobian<F>>(count);
  CU_TRY(ctx, cudaMemcpyAsync(da, a, ja, cudaMemcpyHostToDevice, dc.stream));
  if (bb) CU_TRY(ctx, cudaMemcpyAsync(db, b, bb, cudaMemcpyHostToDevice, dc.stream));
  k_test_ec<F><<<(uint32_t)((count + 63) / 64), 64, 0, dc.stream>>>(op, da, db, dout, (uint32_t)count);
  dc.launches += 1;
  CU_TRY(ctx, cudaMemcpyAsync(out, dout, ja, cudaMemcpyDeviceToHost, dc.stream));
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

template <class F>
int to_affine_impl(msm_ctx* ctx, const void* jac, size_t count, int mont_out, void* out_xy, uint8_t* out_inf) {
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const size_t jb = count * sizeof(ApiJacobian<F>), ab = count * sizeof(ApiAffine<F>);
  CU_TRY(ctx, dc.io.ensure(Arena::padded(jb) + Arena::padded(ab) + Arena::padded(count)));
  ApiJacobian<F>* dj = dc.io.take<ApiJacobian<F>>(count);
  ApiAffine<F>* da = dc.io.take<ApiAffine<F>>(count);
  uint8_t* di = dc.io.take<uint8_t>(count);
  CU_TRY(ctx, cudaMemcpyAsync(dj, jac, jb, cudaMemcpyHostToDevice, dc.stream));
  k_to_affine<F><<<(uint32_t)((count + 63) / 64), 64, 0, dc.stream>>>(dj, (uint32_t)count, mont_out, da, di);
  dc.launches += 1;
  CU_TRY(ctx, cudaMemcpyAsync(out_xy, da, ab, cudaMemcpyDeviceToHost, dc.stream));
  if (out_inf) CU_TRY(ctx, cudaMemcpyAsync(out_inf, di, count, cudaMemcpyDeviceToHost, dc.stream));
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

template <class F> int sum_points_impl(msm_ctx* ctx, const void* d_in, size_t count, void* d_out) {
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  k_sum_points<F><<<1, 32, 0, dc.stream>>>(static_cast<const ApiJacobian<F>*>(d_in), (uint32_t)count,
                                          static_cast<ApiJacobian<F>*>(d_out));
  dc.launches += 1;
  CU_TRY(ctx, cudaGetLastError());
  CU_TRY(ctx, cudaStreamSynchronize(dc.stream));
  return MSM_OK;
}

// radix_ec_fft (ag-cuda-ec/src/ec_fft.rs:13-99): in-place DFT of 2^log_n Jacobian points.
// device_io: jac is a device pointer on device 0 (omegas are always a small host array).
template <class F>
int ec_fft_impl(msm_ctx* ctx, void* jac, uint32_t log_n, const void* omegas_mont, uint32_t n_omegas, bool device_io) {
  using PR = typename std::conditional<is_bn254<F>(), Bn254Fr, Bls381Fr>::type;
  if (log_n > 26 || n_omegas < log_n || n_omegas > 64) {
    set_error(ctx, "ec_fft: need 2^log_n <= 2^26 points and omegas[i] = omega^(2^i) for i < log_n");
    return MSM_ERR_INVALID;
  }
  if (log_n == 0) return MSM_OK;
  DeviceCtx& dc = ctx->devs[0];
  CU_TRY(ctx, cudaSetDevice(dc.dev));
  const uint32_t n = 1u << log_n;
  const size_t jb = (size_t)n * sizeof(ApiJacobian<F>);
  CU_TRY(ctx, dc.io.ensure((device_io ? 0 : Arena::padded(jb)) + Arena::padded((size_t)n_omegas * 32) +
                           Arena::padded((size_t)(n / 2) * 32) + Arena::padded(n / 2)));
  ApiJacobian<F>* d_jac = device_io ? static_cast<ApiJacobian<F>*>(jac) : dc.io.take<ApiJacobian<F>>(n);
  uint32_t* d_omegas = dc.io.take<uint32_t>((size_t)n_omegas * 8);
  uint32_t* d_tw = dc.io.take<uint32_t>((size_t)(n / 2) * 8);
  // G1: twiddles are stored split by the GLV endomorphism (ecfft.cuh); MSM_B200_ECFFT_GLV=0 keeps them whole
  const char* glv_env = getenv("MSM_B200_ECFFT_GLV");
  const bool use_glv = !is_ext2<F>() && !(glv_env && atoi(glv_env) == 0);
  uint8_t* d_tw_sign = use_glv ? dc.io.take<uint8_t>(n / 2) : nullptr;
  const GlvParams gp = glv_params(is_bn254<F>());
  ApiElem<F> beta_api;
  for (int i = 0; i < F::API_WORDS; i++) beta_api.w[i] = i < 12 ? gp.beta[i < 12 ? i : 0] : 0u;
  CU_TRY(ctx, dc.arena.ensure(Arena::padded((size_t)n * sizeof(Xyzz<F>))));
  Xyzz<F>* x = dc.arena.take<Xyzz<F>>(n);
  cudaStream_t st = dc.stream;
  CU_TRY(ctx, cudaEventRecord(dc.ev[0], st));
  if (!device_io) CU_TRY(ctx, cudaMemcpyAsync(d_jac, jac, jb, cudaMemcpyHostToDevice, st));
  CU_TRY(ctx, cudaMemcpyAsync(d_omegas, omegas_mont, (size_t)n_omegas * 32, cudaMemcpyHostToDevice, st));
  CU_TRY(ctx, cudaEventRecord(dc.ev[1], st));
  k_fft_twiddles<PR><<<(n / 2 + 127) / 128, 128, 0, st>>>(d_omegas, n / 2, d_tw, gp, d_tw_sign);
  k_fft_load<F><<<(n + 127) / 128, 128, 0, st>>>(d_jac, log_n, x);
  dc.launches += 2;
  for (uint32_t s = 0; s < log_n; s++) {
    if (aborted(ctx)) {  // SingleEcFftKernel polls maybe_abort once per round (ec-gpu-proxy/src/ec_fft.rs:104-108)
      cudaStreamSynchronize(st);
      return MSM_ERR_ABORTED;
    }
    const uint32_t m = 1u << s;
    k_fft_round<F><<<(n / 2 + 63) / 64, 64, 0, st>>>(x, n, m, n / (2 * m), d_tw, d_tw_sign, beta_api);
    dc.launches += 1;
  }
  k_fft_store<F><<<(n + 127) / 128, 128, 0, st>>>(x, n, d_jac);
  dc.launches += 1;
  CU_TRY(ctx, cudaEventRecord(dc.ev[4], st));
  CU_TRY(ctx, cudaGetLastError());
  if (!device_io) CU_TRY(ctx, cudaMemcpyAsync(jac, d_jac, jb, cudaMemcpyDeviceToHost, st));
  CU_TRY(ctx, cudaStreamSynchronize(st));
  msm_timings& t = ctx->tm;
  memset(&t, 0, sizeof(t));
  cudaEventElapsedTime(&t.h2d_ms, dc.ev[0], dc.ev[1]);
  cudaEventElapsedTime(&t.total_ms, dc.ev[1], dc.ev[4]);
  t.kernel_launches = dc.launches;
  return MSM_OK;
}

// The entry points of one field class in two halves, so that the slow-to-compile Fq2 instantiations can be split
// over two translation units (inst_*_g2.cu: the MSM path; inst_*_g2_aux.cu: EC-FFT, helpers, test kernels).
template <class F>
int describe_plan_impl(msm_ctx* ctx, uint32_t L, uint32_t n_lines, uint32_t num_chunks, uint32_t table_c, uint32_t n_sub,
                       double growth, msm_plan_info* out) {
  Plan pl;
  const int rc = make_plan<F>(ctx, L, n_lines, num_chunks, pl, table_c, n_sub ? n_sub : 1, 0, growth > 0 ? growth : 2.0);
  if (rc) return rc;
  memset(out, 0, sizeof(*out));
  out->window_bits = pl.geo.c;
  out->num_windows = pl.geo.W;
  out->buckets = pl.geo.NB;
  out->sub_batches = pl.n_sub;
  out->by_task = pl.by_task ? 1u : 0u;
  for (int k = 0; k < 9; k++) out->sub_first[k] = pl.sub_first[k];
  out->slice_len = pl.S;
  out->slices = pl.n_slices;
  out->wave_slices = pl.wave_slices;
  out->waves = pl.waves;
  out->sort_mode = pl.sort_mode;
  out->reduce_q = pl.Q;
  out->digits_max = pl.E_max;
  out->scratch_bytes = pl.scratch_bytes;
  return MSM_OK;
}

template <class F> void fill_field_ops_msm(FieldOps& o, const char* name) {
  o.name = name;
  o.api_point_bytes = sizeof(ApiAffine<F>);
  o.packed_point_bytes = sizeof(PackedAffine<F>);
  o.multiple_multiexp = &multiple_multiexp_impl<F>;
  o.multiexp = &multiexp_impl<F>;
  o.convert_bases = &convert_bases_impl<F>;
  o.build_table = [](msm_ctx* c, msm_bases::Shard& sh, uint32_t w, size_t cl) { return build_table_impl<F>(c, sh, w, cl, false); };
  o.describe_plan = &describe_plan_impl<F>;
}
template <class F> void fill_field_ops_aux(FieldOps& o) {
  o.synth_points = &synth_points_impl<F>;
  o.test_fq = &test_fq_impl<F>;
  o.test_ec = &test_ec_impl<F>;
  o.to_affine = &to_affine_impl<F>;
  o.sum_points = &sum_points_impl<F>;
  o.ec_fft = &ec_fft_impl<F>;
}
template <class F> FieldOps make_field_ops(const char* name) {
  FieldOps o;
  fill_field_ops_msm<F>(o, name);
  fill_field_ops_aux<F>(o);
  return o;
}

}  // namespace msm
