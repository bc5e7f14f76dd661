// sort.h -- host-side interface of the digit sort (sort.cu).  The sort kernels do not depend on the field, so
// they are compiled once instead of once per (curve, field) instantiation unit.
#pragma once
#include "engine_common.h"
#include "ptx.cuh"

namespace msm {

struct Geometry {
  uint32_t L;          // scalars per row actually used (= num_chunks * chunk_len)
  uint32_t chunk_len;  // points per task
  uint32_t num_chunks; // scalar-side tasks
  uint32_t c;          // window bits
  uint32_t W;          // windows
  uint32_t B;          // buckets per (task, window) = 2^(c-1)
  uint32_t NB;         // num_chunks * W * B   (folded: num_chunks * B)
  uint32_t scalar_bits;
  uint32_t fold;       // 1: bases are a window table T[w][i] = 2^(c w) P_i, all windows of a task share one bucket set
  uint32_t table_stride;  // points per window in the table (= points of the resident shard)
  uint32_t point_offset;  // folded sub-batches: index of this batch's first point in the table
  uint32_t mont;          // scalars arrive in Montgomery form (arkworks' in-memory Fr) and are converted while they
                          // are decomposed: 0 no (canonical), 1 BN254 Fr, 2 BLS12-381 Fr
};
MSM_HD uint32_t task_of(uint32_t i, const Geometry& geo) { return geo.num_chunks == 1 ? 0u : i / geo.chunk_len; }

constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;
constexpr uint32_t BIN_TILE = 16384;

struct Plan {
  Geometry geo;
  uint32_t n_lines, S, n_slices, Q, RW, PG, n_tasks;
  uint32_t slices_cap;  // upper bound of the slices of one sub-batch
  // k_accumulate grids are whole waves rounded down (make_plan): slices of one wave, waves of the longest sub-batch;
  // waves == 0: S was imposed (environment, minimum for few buckets) and every sub-batch keeps it
  uint32_t wave_slices = 0, waves = 0;
  uint32_t acc_blocks_per_sm = 4;  // resident blocks of the bucket kernel per SM (occupancy API; 4 without a device)
  uint32_t W_sets;  // bucket sets per task: W, or 1 when the bases are a folded window table
  uint32_t n_sub;   // sub-batches of a pipelined call: parts of one MSM (they continue one shared bucket array) ...
  bool by_task = false;  // ... or groups of whole tasks of a many-task row (each owns its range of the bucket array)
  uint32_t sub_first[9];  // sub-batch k covers scalars [sub_first[k], sub_first[k+1]); sizes grow geometrically
  uint32_t sub_max;       // longest sub-batch
  mutable uint32_t scatter_passes = 1;  // filled in by enqueue_msm (0: two-level partition scatter)
  bool partition;   // two-level scatter through a (bucket id, entry) temporary, global cursor atomics (measured slower)
  uint32_t sort_mode;  // 0: single-level atomic sort, 2: binned sort (shared-memory atomics)
  uint32_t bin_shift;  // binned sort: low bucket-id bits sorted inside a bin
  uint64_t E_max;
  size_t scratch_bytes;
  // affine halving rounds before the XYZZ slice kernel (bucket_affine.cuh); 0: none
  uint32_t ba_rounds = 0, ba_threads = 0, ba_batch = 0, ba_bps = 0;
  uint32_t S_tail = 0, n_slices_tail = 0;  // slice geometry of the XYZZ kernel over the points the rounds leave
  uint64_t ba_cap[2] = {0, 0};             // capacity (points) of the two ping-pong point arrays
};


struct SortBuffers {
  uint32_t *counts, *bucket_start, *cursor, *tile_sums, *entries;
};

// exclusive scan of b.counts[0 .. NB) -> b.bucket_start[0 .. NB] (closing element = total) and a copy in b.cursor
void enqueue_bucket_scan(cudaStream_t st, const Geometry& g, const SortBuffers& b);
// the same over counts[g] = ceil((off_in[g+1] - off_in[g]) / 2): offsets of the next affine halving round
void enqueue_halve_scan(cudaStream_t st, const Geometry& g, const uint32_t* off_in, const SortBuffers& b);
// The three sorts.  All leave bucket_start[NB+1] (exclusive scan of the bucket sizes, bucket_start[NB] = number of
// non-zero digits) and the entries in bucket order.  sg is the geometry of the (sub-)batch, E_max its digit bound;
// temporaries come from the scratch arena; `on`: the stream to run on (default: the context's main stream).
int enqueue_sort_binned(msm_ctx* ctx, DeviceCtx& dc, const Plan& pl, const Geometry& sg, uint64_t E_max,
                        const uint32_t* scalars, const SortBuffers& b, cudaStream_t on = nullptr);
int enqueue_sort_partition(msm_ctx* ctx, DeviceCtx& dc, const Plan& pl, const Geometry& sg, uint64_t E_max,
                           const uint32_t* scalars, const SortBuffers& b, cudaStream_t on = nullptr);
int enqueue_sort_atomic(msm_ctx* ctx, DeviceCtx& dc, const Plan& pl, const Geometry& sg, uint64_t E_max,
                        const uint32_t* scalars, const SortBuffers& b, cudaStream_t on = nullptr);

}  // namespace msm
