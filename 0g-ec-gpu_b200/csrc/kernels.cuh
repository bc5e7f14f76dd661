// kernels.cuh -- the device side of the MSM engine.
//
// Pipeline for one call (T = num_chunks scalar-side tasks, W windows of c bits, B = 2^(c-1)
// buckets per (task, window), `lines` base lines sharing the scalar row):
//
//   sort              signed-digit (Booth) decomposition of every scalar and a counting sort of the
//                     (point index | sign) entries by bucket id -> bucket_start[NB+1], entries[]:
//                       small calls: k_digits<COUNT>, k_scan_*, k_digits<SCATTER> (one L2 atomic per digit)
//                       large calls: k_bin_count, k_bin_scan, k_partition, k_bin_hist, k_scan_*,
//                                    k_bin_place (every per-digit atomic in shared memory)
//   k_accumulate      one thread per fixed-size slice of the sorted entry list; XYZZ mixed adds;
//                     whole buckets are written directly, buckets cut by a slice boundary go to
//                     per-slice partial slots
//   k_fixup_cut       combines the partial slots of the buckets cut by slice boundaries (listed by k_accumulate)
//   k_bucket_reduce   running-sum reduction  sum_b b*S_b  split over threads + shared-memory tree
//   k_window_combine  sums the per-window partials, Horner over windows, XYZZ -> Jacobian
//
// This replaces POINT_multiexp / POINT_multiexp_chunk / POINT_aggregate_chunk
// (ag-build/cl/multiexp.cl:62-264), where one thread owns a whole (task, window) and scans the
// chunk serially against buckets in global memory.
#pragma once
#include "ec.cuh"

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
};
MSM_HD uint32_t task_of(uint32_t i, const Geometry& geo) { return geo.num_chunks == 1 ? 0u : i / geo.chunk_len; }

// ---------------------------------------------------------------------------------------------
// Signed-digit decomposition.  k = sum_w d_w 2^(c w),  d_w in [-(2^(c-1) - 1), 2^(c-1)].
// W*c >= scalar_bits + 1 guarantees no carry out of the top window (the reference's kernel drops
// that carry, TODO at ag-build/cl/multiexp.cl:60).  Bits are taken LSB-first from the canonical
// little-endian scalar (the reference indexes MSB-first, ag-build/cl/field.cl:380-392; the digit
// set is a free choice because the result does not depend on it).
// f(w, bucket_1based, negative) is called for every non-zero digit.
// ---------------------------------------------------------------------------------------------
template <class F> MSM_D void for_each_digit(const uint32_t k[8], uint32_t c, uint32_t W, F&& f) {
  const uint32_t half = 1u << (c - 1);
  const uint32_t mask = (1u << c) - 1u;
  uint64_t buf = 0;
  uint32_t nbits = 0, w = 0, carry = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    buf |= (uint64_t)k[j] << nbits;
    nbits += 32;
    while (nbits >= c && w < W) {
      uint32_t raw = ((uint32_t)buf & mask) + carry;
      buf >>= c;
      nbits -= c;
      carry = raw > half;
      if (raw != 0 && raw != (1u << c)) {
        if (carry) f(w, (1u << c) - raw, true);
        else f(w, raw, false);
      }
      w++;
    }
  }
  // remaining high bits (fewer than c)
  while (w < W) {
    uint32_t raw = ((uint32_t)buf & mask) + carry;
    buf >>= c;
    carry = raw > half;
    if (raw != 0 && raw != (1u << c)) {
      if (carry) f(w, (1u << c) - raw, true);
      else f(w, raw, false);
    }
    w++;
  }
}

// Same decomposition with the window size known at compile time: every digit is one funnel shift
// and one mask on registers (the generic loop above spends ~3x the instructions on 64-bit buffer
// shifts; with 4-5 decomposition passes per call that was ~2 ms of a 2^24 MSM).
template <int C, class F> MSM_D void for_each_digit_c(const uint32_t k[8], uint32_t W, F&& f) {
  constexpr uint32_t half = 1u << (C - 1);
  constexpr uint32_t mask = (1u << C) - 1u;
  constexpr int MAXW = (256 + C - 1) / C + 1;
  uint32_t carry = 0;
#pragma unroll
  for (int w = 0; w < MAXW; w++) {
    if ((uint32_t)w >= W) break;
    const int bit = w * C, word = bit >> 5, sh = bit & 31;
    const uint32_t lo = word < 8 ? k[word < 8 ? word : 0] : 0u;
    const uint32_t hi = word + 1 < 8 ? k[word + 1 < 8 ? word + 1 : 0] : 0u;
    const uint32_t raw = (__funnelshift_r(lo, hi, sh) & mask) + carry;
    carry = raw > half;
    if (raw != 0 && raw != (1u << C)) {
      if (carry) f((uint32_t)w, (1u << C) - raw, true);
      else f((uint32_t)w, raw, false);
    }
  }
}

MSM_D void load_scalar(const uint32_t* scalars, uint32_t i, uint32_t k[8]) {
  const uint4* p = reinterpret_cast<const uint4*>(scalars) + 2 * (size_t)i;
  uint4 a = __ldg(p), b = __ldg(p + 1);
  k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w;
  k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
}

// SCATTER = false: counts[g]++ ;  SCATTER = true: entries[cursor[g]++] = i | sign<<31
// g_lo / g_hi: only digits whose bucket id lies in [g_lo, g_hi) are handled; the scatter runs in
// several such passes so that the randomly written slice of `entries` stays resident in the L2.
template <bool SCATTER, int C>
__global__ void k_digits(const uint32_t* __restrict__ scalars, Geometry geo,
                         uint32_t* __restrict__ counts_or_cursor, uint32_t* __restrict__ entries,
                         uint32_t g_lo, uint32_t g_hi) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= geo.L) return;
  uint32_t k[8];
  load_scalar(scalars, i, k);
  const uint32_t task = i / geo.chunk_len;
  const uint32_t base = task * geo.W;
  auto body = [&](uint32_t w, uint32_t bucket, bool neg) {
    const uint32_t g = geo.fold ? task_of(i, geo) * geo.B + (bucket - 1) : (base + w) * geo.B + (bucket - 1);
    if (g < g_lo || g >= g_hi) return;
    // The top window only carries the few leftover scalar bits, so all points share a handful of
    // its buckets: aggregate those atomics per warp (one atomic per distinct bucket).
    uint32_t rank = 0, total = 1, leader_lane = 0;
    const bool aggregate = (w + 1 == geo.W);
    uint32_t peers = 0;
    if (aggregate) {
      peers = __match_any_sync(__activemask(), g);
      leader_lane = __ffs(peers) - 1;
      total = __popc(peers);
      rank = __popc(peers & ((1u << (threadIdx.x & 31)) - 1));
    }
    if (SCATTER) {
      uint32_t pos;
      if (aggregate) {
        uint32_t base_pos = 0;
        if ((threadIdx.x & 31) == leader_lane) base_pos = atomicAdd(&counts_or_cursor[g], total);
        pos = __shfl_sync(peers, base_pos, leader_lane) + rank;
      } else {
        pos = atomicAdd(&counts_or_cursor[g], 1u);
      }
      const uint32_t idx = geo.fold ? w * geo.table_stride + geo.point_offset + i : i;
      entries[pos] = idx | (neg ? 0x80000000u : 0u);
    } else {
      if (!aggregate) atomicAdd(&counts_or_cursor[g], 1u);
      else if ((threadIdx.x & 31) == leader_lane) atomicAdd(&counts_or_cursor[g], total);
    }
  };
  if (C == 0) for_each_digit(k, geo.c, geo.W, body);
  else for_each_digit_c<(C == 0 ? 8 : C)>(k, geo.W, body);
}

// launch with the window size as a template argument where an instantiation exists
template <bool SCATTER>
inline void launch_digits(uint32_t grid, uint32_t block, cudaStream_t st, const uint32_t* scalars, const Geometry& geo,
                          uint32_t* counts_or_cursor, uint32_t* entries, uint32_t g_lo, uint32_t g_hi) {
#define MSM_DIGITS_CASE(CC) \
  case CC: k_digits<SCATTER, CC><<<grid, block, 0, st>>>(scalars, geo, counts_or_cursor, entries, g_lo, g_hi); break;
  switch (geo.c) {
    MSM_DIGITS_CASE(6) MSM_DIGITS_CASE(7) MSM_DIGITS_CASE(8) MSM_DIGITS_CASE(9) MSM_DIGITS_CASE(10)
    MSM_DIGITS_CASE(11) MSM_DIGITS_CASE(12) MSM_DIGITS_CASE(13) MSM_DIGITS_CASE(14) MSM_DIGITS_CASE(15)
    MSM_DIGITS_CASE(16) MSM_DIGITS_CASE(17) MSM_DIGITS_CASE(18) MSM_DIGITS_CASE(19) MSM_DIGITS_CASE(20)
    MSM_DIGITS_CASE(21) MSM_DIGITS_CASE(22) MSM_DIGITS_CASE(23) MSM_DIGITS_CASE(24)
    default: k_digits<SCATTER, 0><<<grid, block, 0, st>>>(scalars, geo, counts_or_cursor, entries, g_lo, g_hi);
  }
#undef MSM_DIGITS_CASE
}

// ---------------------------------------------------------------------------------------------
// Exclusive scan of n uint32 (three small kernels; n <= a few million, HBM-trivial).
// ---------------------------------------------------------------------------------------------
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_BLOCK * SCAN_ITEMS;

template <int BS> __device__ __forceinline__ uint32_t block_exclusive_scan_t(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[BS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t s = lane < BS / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < BS / 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane < BS / 32) warp_sums[lane] = s;
  }
  __syncthreads();
  const uint32_t warp_off = wid ? warp_sums[wid - 1] : 0;
  *total = warp_sums[BS / 32 - 1];
  __syncthreads();
  return warp_off + x - v;
}
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  return block_exclusive_scan_t<SCAN_BLOCK>(v, total);
}

static __global__ void k_scan_tiles(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out,
                             uint32_t* __restrict__ tile_sums) {
  const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; j++) {
    v[j] = base + j < n ? in[base + j] : 0;
    s += v[j];
  }
  uint32_t total;
  uint32_t off = block_exclusive_scan(s, &total);
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; j++) {
    if (base + j < n) out[base + j] = off;
    off += v[j];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// single block: exclusive scan of tile sums in place, total written to *grand_total
static __global__ void k_scan_tile_sums(uint32_t* tile_sums, uint32_t n_tiles, uint32_t* grand_total) {
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n_tiles; base += SCAN_BLOCK) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < n_tiles ? tile_sums[i] : 0;
    uint32_t total;
    uint32_t off = block_exclusive_scan(v, &total);
    if (i < n_tiles) tile_sums[i] = carry + off;
    carry += total;
  }
  if (threadIdx.x == 0) *grand_total = carry;
}
// out[i] += tile_offset; also writes the closing element out[n] = grand_total and a copy (cursor)
static __global__ void k_scan_finish(uint32_t* __restrict__ out, uint32_t n, const uint32_t* __restrict__ tile_sums,
                              const uint32_t* __restrict__ grand_total, uint32_t* __restrict__ cursor) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t v = out[i] + tile_sums[i / SCAN_TILE];
    out[i] = v;
    cursor[i] = v;
  } else if (i == n) {
    out[n] = *grand_total;
  }
}

// ---------------------------------------------------------------------------------------------
// Two-level scatter (large calls).  A single-pass scatter writes 4-byte entries at random over
// hundreds of MB: every write costs a 32-byte sector of DRAM traffic.  Instead:
//   k_partition      each block decomposes a tile of scalars, groups its digits by the HIGH bits
//                    of the bucket id in shared memory and appends every group to that high-bin's
//                    region of a temporary (bucket id, entry) array in coalesced runs;
//   k_final_scatter  walks the temporary array (now ordered by high-bin) and places every entry
//                    with the usual cursor atomic -- the cursors and the destination slice touched
//                    at any moment are a few MB and stay in the L2.
// hb = g >> bin_shift; hb_region[hb] = bucket_start[hb << bin_shift] is where bin hb starts.
// ---------------------------------------------------------------------------------------------
constexpr int PART_BLOCK = 512;
template <int C>
__global__ void __launch_bounds__(PART_BLOCK)
k_partition(const uint32_t* __restrict__ scalars, Geometry geo, uint32_t tile, uint32_t bin_shift, uint32_t n_bins,
            const uint32_t* __restrict__ region_start, uint32_t region_shift, uint32_t* __restrict__ bin_cursor,
            uint32_t* __restrict__ tmp_g, uint32_t* __restrict__ tmp_v) {
  extern __shared__ uint32_t part_smem[];
  uint32_t* hist = part_smem;                 // [n_bins] counts, then running cursors
  uint32_t* off = hist + n_bins;              // [n_bins] exclusive offsets inside the block
  uint32_t* gbase = off + n_bins;             // [n_bins] global base of this block's run
  uint32_t* stage_g = gbase + n_bins;         // [tile * W]
  uint32_t* stage_v = stage_g + (size_t)tile * geo.W;
  __shared__ uint32_t total_sh;
  const uint32_t first = blockIdx.x * tile;
  for (uint32_t b = threadIdx.x; b < n_bins; b += PART_BLOCK) hist[b] = 0;
  __syncthreads();
  // pass 1: histogram of high bins
  for (uint32_t t = threadIdx.x; t < tile; t += PART_BLOCK) {
    const uint32_t i = first + t;
    if (i >= geo.L) break;
    uint32_t k[8];
    load_scalar(scalars, i, k);
    const uint32_t base = (i / geo.chunk_len) * geo.W;
    auto body = [&](uint32_t w, uint32_t bucket, bool) {
      const uint32_t g = geo.fold ? task_of(i, geo) * geo.B + (bucket - 1) : (base + w) * geo.B + (bucket - 1);
      atomicAdd(&hist[g >> bin_shift], 1u);
    };
    if (C == 0) for_each_digit(k, geo.c, geo.W, body);
    else for_each_digit_c<(C == 0 ? 8 : C)>(k, geo.W, body);
  }
  __syncthreads();
  // exclusive scan of the bins (n_bins <= 4 * PART_BLOCK), one global reservation per non-empty bin
  {
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint32_t b = threadIdx.x * 4 + q;
      v[q] = b < n_bins ? hist[b] : 0;
      sum += v[q];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan_t<PART_BLOCK>(sum, &total);
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint32_t b = threadIdx.x * 4 + q;
      if (b < n_bins) {
        off[b] = run;
        gbase[b] = v[q] ? region_start[b << region_shift] + atomicAdd(&bin_cursor[b], v[q]) : 0;
        hist[b] = 0;
      }
      run += v[q];
    }
    if (threadIdx.x == 0) total_sh = total;
  }
  __syncthreads();
  // pass 2: same decomposition, place (g, entry) in the block-local bin order
  for (uint32_t t = threadIdx.x; t < tile; t += PART_BLOCK) {
    const uint32_t i = first + t;
    if (i >= geo.L) break;
    uint32_t k[8];
    load_scalar(scalars, i, k);
    const uint32_t base = (i / geo.chunk_len) * geo.W;
    auto body = [&](uint32_t w, uint32_t bucket, bool neg) {
      const uint32_t g = geo.fold ? task_of(i, geo) * geo.B + (bucket - 1) : (base + w) * geo.B + (bucket - 1);
      const uint32_t hb = g >> bin_shift;
      const uint32_t slot = off[hb] + atomicAdd(&hist[hb], 1u);
      const uint32_t idx = geo.fold ? w * geo.table_stride + geo.point_offset + i : i;
      stage_g[slot] = g;
      stage_v[slot] = idx | (neg ? 0x80000000u : 0u);
    };
    if (C == 0) for_each_digit(k, geo.c, geo.W, body);
    else for_each_digit_c<(C == 0 ? 8 : C)>(k, geo.W, body);
  }
  __syncthreads();
  // write every bin's run to its region: consecutive slots of one bin are consecutive in memory
  const uint32_t total = total_sh;
  for (uint32_t sidx = threadIdx.x; sidx < total; sidx += PART_BLOCK) {
    const uint32_t g = stage_g[sidx];
    const uint32_t hb = g >> bin_shift;
    const uint32_t dst = gbase[hb] + (sidx - off[hb]);
    tmp_g[dst] = g;
    tmp_v[dst] = stage_v[sidx];
  }
}

static __global__ void k_final_scatter(const uint32_t* __restrict__ tmp_g, const uint32_t* __restrict__ tmp_v,
                                       const uint32_t* __restrict__ E_ptr, uint32_t* __restrict__ cursor,
                                       uint32_t* __restrict__ entries) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= __ldg(E_ptr)) return;
  const uint32_t pos = atomicAdd(&cursor[__ldg(tmp_g + i)], 1u);
  entries[pos] = __ldg(tmp_v + i);
}

template <int C> inline cudaError_t partition_set_smem(size_t bytes) {
  return cudaFuncSetAttribute(k_partition<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
inline cudaError_t launch_partition(uint32_t grid, size_t smem, cudaStream_t st, const uint32_t* scalars,
                                    const Geometry& geo, uint32_t tile, uint32_t bin_shift, uint32_t n_bins,
                                    const uint32_t* region_start, uint32_t region_shift, uint32_t* bin_cursor,
                                    uint32_t* tmp_g, uint32_t* tmp_v) {
  cudaError_t e = cudaSuccess;
#define MSM_PART_CASE(CC)                                                                                        \
  case CC:                                                                                                       \
    e = partition_set_smem<CC>(smem);                                                                            \
    if (e == cudaSuccess)                                                                                        \
      k_partition<CC><<<grid, PART_BLOCK, smem, st>>>(scalars, geo, tile, bin_shift, n_bins, region_start,       \
                                                      region_shift, bin_cursor, tmp_g, tmp_v);                   \
    break;
  switch (geo.c) {
    MSM_PART_CASE(16) MSM_PART_CASE(17) MSM_PART_CASE(18) MSM_PART_CASE(19) MSM_PART_CASE(20)
    MSM_PART_CASE(21) MSM_PART_CASE(22) MSM_PART_CASE(23) MSM_PART_CASE(24)
    default:
      e = partition_set_smem<0>(smem);
      if (e == cudaSuccess)
        k_partition<0><<<grid, PART_BLOCK, smem, st>>>(scalars, geo, tile, bin_shift, n_bins, region_start, region_shift,
                                                       bin_cursor, tmp_g, tmp_v);
  }
#undef MSM_PART_CASE
  return e;
}

// ---------------------------------------------------------------------------------------------
// Binned sort (large calls): the same two levels, with every per-digit atomic in SHARED memory.
// The L2 executes ~80 atomics per clock for the whole chip; a 2^24-point call needs 4 x 10^8 of
// them in the single-level sort above (histogram + scatter) and that is what its 4.9 ms are.  Here:
//   k_bin_count   coarse histogram (bins = high bits of the bucket id), per-block in shared memory
//   k_bin_scan    bin offsets, and the number of fixed-size tiles each bin is cut into
//   k_partition   (above) groups the digits by bin into tmp_g / tmp_v in coalesced runs
//   k_bin_hist    one block per tile of one bin: shared-memory histogram of the low bits, then one
//                 global add per non-empty bucket of the tile
//   (scan)        bucket_start / cursor as before
//   k_bin_place   one block per tile: reserves the tile's range of every bucket with one atomic per
//                 non-empty bucket, then places the entries with shared-memory cursors; the writes
//                 of a tile land in the few hundred KB of its bin's slice of `entries`
// Tiles make skewed inputs (and the short top window of a folded table, whose digits all fall into
// the first bins) a matter of more blocks, not of longer ones.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t BIN_TILE = 16384;
constexpr int BIN_BLOCK = 256;
constexpr uint32_t BIN_COUNT_SCALARS = 2048;  // scalars per block in k_bin_count

template <int C>
__global__ void __launch_bounds__(BIN_BLOCK)
k_bin_count(const uint32_t* __restrict__ scalars, Geometry geo, uint32_t bin_shift, uint32_t n_bins,
            uint32_t* __restrict__ bin_count) {
  extern __shared__ uint32_t bin_smem[];
  for (uint32_t b = threadIdx.x; b < n_bins; b += BIN_BLOCK) bin_smem[b] = 0;
  __syncthreads();
  const uint32_t first = blockIdx.x * BIN_COUNT_SCALARS;
  for (uint32_t t = threadIdx.x; t < BIN_COUNT_SCALARS; t += BIN_BLOCK) {
    const uint32_t i = first + t;
    if (i >= geo.L) break;
    uint32_t k[8];
    load_scalar(scalars, i, k);
    const uint32_t base = (i / geo.chunk_len) * geo.W;
    auto body = [&](uint32_t w, uint32_t bucket, bool) {
      const uint32_t g = geo.fold ? task_of(i, geo) * geo.B + (bucket - 1) : (base + w) * geo.B + (bucket - 1);
      atomicAdd(&bin_smem[g >> bin_shift], 1u);
    };
    if (C == 0) for_each_digit(k, geo.c, geo.W, body);
    else for_each_digit_c<(C == 0 ? 8 : C)>(k, geo.W, body);
  }
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < n_bins; b += BIN_BLOCK)
    if (bin_smem[b]) atomicAdd(&bin_count[b], bin_smem[b]);
}
inline void launch_bin_count(uint32_t grid, cudaStream_t st, const uint32_t* scalars, const Geometry& geo,
                             uint32_t bin_shift, uint32_t n_bins, uint32_t* bin_count) {
  const size_t smem = (size_t)n_bins * 4;
#define MSM_BINC_CASE(CC) \
  case CC: k_bin_count<CC><<<grid, BIN_BLOCK, smem, st>>>(scalars, geo, bin_shift, n_bins, bin_count); break;
  switch (geo.c) {
    MSM_BINC_CASE(8) MSM_BINC_CASE(9) MSM_BINC_CASE(10) MSM_BINC_CASE(11) MSM_BINC_CASE(12) MSM_BINC_CASE(13)
    MSM_BINC_CASE(14) MSM_BINC_CASE(15) MSM_BINC_CASE(16) MSM_BINC_CASE(17) MSM_BINC_CASE(18) MSM_BINC_CASE(19)
    MSM_BINC_CASE(20) MSM_BINC_CASE(21) MSM_BINC_CASE(22) MSM_BINC_CASE(23) MSM_BINC_CASE(24)
    default: k_bin_count<0><<<grid, BIN_BLOCK, smem, st>>>(scalars, geo, bin_shift, n_bins, bin_count);
  }
#undef MSM_BINC_CASE
}

// single block; n_bins <= 4 * SCAN_BLOCK.  bin_start / tile_start have n_bins + 1 elements.
static __global__ void k_bin_scan(const uint32_t* __restrict__ bin_count, uint32_t n_bins,
                                  uint32_t* __restrict__ bin_start, uint32_t* __restrict__ tile_start) {
  uint32_t c[4], tl[4], sc = 0, stl = 0;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint32_t b = threadIdx.x * 4 + q;
    c[q] = b < n_bins ? bin_count[b] : 0;
    tl[q] = (c[q] + BIN_TILE - 1) / BIN_TILE;
    sc += c[q];
    stl += tl[q];
  }
  uint32_t total_c, total_t;
  uint32_t off_c = block_exclusive_scan(sc, &total_c);
  uint32_t off_t = block_exclusive_scan(stl, &total_t);
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const uint32_t b = threadIdx.x * 4 + q;
    if (b < n_bins) {
      bin_start[b] = off_c;
      tile_start[b] = off_t;
    }
    off_c += c[q];
    off_t += tl[q];
  }
  if (threadIdx.x == 0) {
    bin_start[n_bins] = total_c;
    tile_start[n_bins] = total_t;
  }
}

// which bin does tile `tile` belong to, and which entries of tmp_* does it cover
MSM_D bool bin_tile_range(const uint32_t* __restrict__ bin_start, const uint32_t* __restrict__ tile_start,
                          uint32_t n_bins, uint32_t tile, uint32_t& bin, uint32_t& lo, uint32_t& hi) {
  if (tile >= __ldg(tile_start + n_bins)) return false;
  uint32_t a = 0, b = n_bins;  // invariant: tile_start[a] <= tile < tile_start[b]
  while (b - a > 1) {
    const uint32_t mid = (a + b) >> 1;
    if (__ldg(tile_start + mid) <= tile) a = mid;
    else b = mid;
  }
  bin = a;
  lo = __ldg(bin_start + a) + (tile - __ldg(tile_start + a)) * BIN_TILE;
  hi = min(lo + BIN_TILE, __ldg(bin_start + a + 1));
  return true;
}

static __global__ void __launch_bounds__(BIN_BLOCK)
k_bin_hist(const uint32_t* __restrict__ tmp_g, const uint32_t* __restrict__ bin_start,
           const uint32_t* __restrict__ tile_start, uint32_t n_bins, uint32_t bin_shift, uint32_t NB,
           uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t bin_smem[];
  const uint32_t bpb = 1u << bin_shift;
  uint32_t bin, lo, hi;
  if (!bin_tile_range(bin_start, tile_start, n_bins, blockIdx.x, bin, lo, hi)) return;
  for (uint32_t b = threadIdx.x; b < bpb; b += BIN_BLOCK) bin_smem[b] = 0;
  __syncthreads();
  for (uint32_t p = lo + threadIdx.x; p < hi; p += BIN_BLOCK) atomicAdd(&bin_smem[__ldg(tmp_g + p) & (bpb - 1)], 1u);
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < bpb; b += BIN_BLOCK) {
    const uint32_t c = bin_smem[b], g = (bin << bin_shift) + b;
    if (c && g < NB) atomicAdd(&counts[g], c);
  }
}

// Placement with the tile sorted in shared memory first, so that the entries of one bucket leave as
// one run of consecutive 4-byte stores (a 32-byte sector for the typical 8 entries per bucket and
// tile) instead of 8 scattered ones: 4-byte scattered stores cost the L2 as much as atomics do.
constexpr int PLACE_BLOCK = 1024;
static __global__ void __launch_bounds__(PLACE_BLOCK)
k_bin_place(const uint32_t* __restrict__ tmp_g, const uint32_t* __restrict__ tmp_v,
            const uint32_t* __restrict__ bin_start, const uint32_t* __restrict__ tile_start, uint32_t n_bins,
            uint32_t bin_shift, uint32_t NB, uint32_t* __restrict__ cursor, uint32_t* __restrict__ entries) {
  extern __shared__ uint32_t bin_smem[];
  const uint32_t bpb = 1u << bin_shift;
  uint32_t* hist = bin_smem;            // [bpb] counts, then running ranks, then (global position - slot) of the bucket
  uint32_t* off = hist + bpb;           // [bpb] first slot of the bucket inside the tile
  uint32_t* stage_v = off + bpb;        // [BIN_TILE]
  uint16_t* stage_b = reinterpret_cast<uint16_t*>(stage_v + BIN_TILE);  // [BIN_TILE] low bucket bits (bpb <= 2^13)
  __shared__ uint32_t warp_sums[PLACE_BLOCK / 32];
  uint32_t bin, lo, hi;
  if (!bin_tile_range(bin_start, tile_start, n_bins, blockIdx.x, bin, lo, hi)) return;
  for (uint32_t b = threadIdx.x; b < bpb; b += PLACE_BLOCK) hist[b] = 0;
  __syncthreads();
  for (uint32_t p = lo + threadIdx.x; p < hi; p += PLACE_BLOCK) atomicAdd(&hist[__ldg(tmp_g + p) & (bpb - 1)], 1u);
  __syncthreads();
  // exclusive scan of hist over the block: thread t owns buckets [t*ipt, (t+1)*ipt), ipt <= 8
  const uint32_t ipt = (bpb + PLACE_BLOCK - 1) / PLACE_BLOCK;
  const uint32_t b0 = threadIdx.x * ipt;
  uint32_t sum = 0;
  for (uint32_t q = 0; q < ipt; q++) sum += b0 + q < bpb ? hist[b0 + q] : 0;
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= (uint32_t)o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    uint32_t v = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= (uint32_t)o) v += y;
    }
    warp_sums[lane] = v;
  }
  __syncthreads();
  // the owner keeps (first global position - first slot) of its buckets in registers until the tile is
  // staged: two shared arrays instead of three let two 1024-thread blocks share an SM
  uint32_t run = (wid ? warp_sums[wid - 1] : 0) + x - sum;
  uint32_t delta[8];
#pragma unroll
  for (uint32_t q = 0; q < 8; q++) {
    const uint32_t b = b0 + q;
    delta[q] = 0;
    if (q < ipt && b < bpb) {
      const uint32_t c = hist[b], g = (bin << bin_shift) + b;
      off[b] = run;
      delta[q] = ((c && g < NB) ? atomicAdd(&cursor[g], c) : 0u) - run;
      run += c;
      hist[b] = 0;
    }
  }
  __syncthreads();
  for (uint32_t p = lo + threadIdx.x; p < hi; p += PLACE_BLOCK) {
    const uint32_t lb = __ldg(tmp_g + p) & (bpb - 1);
    const uint32_t slot = off[lb] + atomicAdd(&hist[lb], 1u);
    stage_v[slot] = __ldg(tmp_v + p);
    stage_b[slot] = (uint16_t)lb;
  }
  __syncthreads();
#pragma unroll
  for (uint32_t q = 0; q < 8; q++)
    if (q < ipt && b0 + q < bpb) hist[b0 + q] = delta[q];
  __syncthreads();
  for (uint32_t slot = threadIdx.x; slot < hi - lo; slot += PLACE_BLOCK) entries[hist[stage_b[slot]] + slot] = stage_v[slot];
}
inline size_t bin_place_smem(uint32_t bin_shift) { return ((size_t)2 << bin_shift) * 4 + (size_t)BIN_TILE * 6; }

// ---------------------------------------------------------------------------------------------
// 128-bit vector loads / stores of plain structs (sizeof multiple of 16, 16-byte aligned).
// ---------------------------------------------------------------------------------------------
template <class T> MSM_D void store_vec(T* dst, const T& v) {
  static_assert(sizeof(T) % 16 == 0, "vector store needs a multiple of 16 bytes");
  constexpr int V = sizeof(T) / 16;
  uint4* q = reinterpret_cast<uint4*>(dst);
  const uint32_t* o = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int j = 0; j < V; j++) q[j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
}
template <class T> MSM_D T load_vec(const T* src) {
  static_assert(sizeof(T) % 16 == 0, "vector load needs a multiple of 16 bytes");
  T r;
  constexpr int V = sizeof(T) / 16;
  const uint4* q = reinterpret_cast<const uint4*>(src);
  uint32_t* o = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int j = 0; j < V; j++) {
    uint4 t = q[j];
    o[4 * j] = t.x; o[4 * j + 1] = t.y; o[4 * j + 2] = t.z; o[4 * j + 3] = t.w;
  }
  return r;
}
// resident base point: 128-bit loads on the read-only path, then re-slice into field limbs.
// Returns false for the identity encoding (all words zero).
template <class F> MSM_D bool load_base(const PackedAffine<F>* p, Affine<F>& out) {
  constexpr int WORDS = 2 * F::PACKED_WORDS;
  static_assert(WORDS % 4 == 0, "packed point must be a multiple of 16 bytes");
  uint32_t w[WORDS];
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint32_t any = 0;
#pragma unroll
  for (int j = 0; j < WORDS / 4; j++) {
    uint4 t = __ldg(q + j);
    w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
    any |= t.x | t.y | t.z | t.w;
  }
  out.x = F::unpack(w);
  out.y = F::unpack(w + F::PACKED_WORDS);
  return any != 0;
}

// The same in two steps, so that the raw words of the NEXT point can be in flight during a mixed add.
template <class F> MSM_D void load_base_words(const PackedAffine<F>* p, uint32_t* w) {
  constexpr int WORDS = 2 * F::PACKED_WORDS;
  static_assert(WORDS % 4 == 0, "packed point must be a multiple of 16 bytes");
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int j = 0; j < WORDS / 4; j++) {
    uint4 t = __ldg(q + j);
    w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
  }
}
template <class F> MSM_D bool unpack_base(const uint32_t* w, Affine<F>& out) {
  constexpr int WORDS = 2 * F::PACKED_WORDS;
  uint32_t any = 0;
#pragma unroll
  for (int j = 0; j < WORDS; j++) any |= w[j];
  out.x = F::unpack(w);
  out.y = F::unpack(w + F::PACKED_WORDS);
  return any != 0;
}

// largest g in [0, NB) with bucket_start[g] <= pos  (pos < bucket_start[NB])
MSM_D uint32_t find_bucket(const uint32_t* __restrict__ bucket_start, uint32_t NB, uint32_t pos) {
  uint32_t lo = 0, hi = NB;  // invariant: bucket_start[lo] <= pos < bucket_start[hi]
  while (hi - lo > 1) {
    uint32_t mid = (lo + hi) >> 1;
    if (__ldg(bucket_start + mid) <= pos) lo = mid;
    else hi = mid;
  }
  return lo;
}

// ---------------------------------------------------------------------------------------------
// Bucket accumulation.  Thread t of line `blockIdx.y` owns sorted entries [t*S, (t+1)*S).
// ---------------------------------------------------------------------------------------------
// Resident blocks per SM: 4 for 8-limb fields (106 registers), 3 for 12-limb fields (168 registers);
// measured: 5 blocks (spills) gains nothing for BN254, forcing 4 on BLS12-381 costs 5 %.  The Fq2
// instantiations (16 / 24 words per element) take the full register file.
template <class F>
__global__ void __launch_bounds__(128, (F::N <= 9 ? 4 : F::N <= 12 ? 3 : F::N <= 16 ? 2 : 1))
k_accumulate(const PackedAffine<F>* __restrict__ bases, uint32_t line_stride,
             const uint32_t* __restrict__ entries, const uint32_t* __restrict__ bucket_start,
             uint32_t NB, const uint32_t* __restrict__ E_ptr, uint32_t S, uint32_t n_slices,
             Xyzz<F>* __restrict__ bucket_acc, Xyzz<F>* __restrict__ partials, uint32_t carry_in,
             uint32_t* __restrict__ cut_count, uint32_t* __restrict__ cut_list, uint32_t cut_cap) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t E = __ldg(E_ptr);  // = bucket_start[NB], number of non-zero digits
  if (t >= n_slices || (uint64_t)t * S >= E) return;
  const uint32_t line = blockIdx.y;
  bases += (size_t)line * line_stride;
  bucket_acc += (size_t)line * NB;
  partials += (size_t)line * 2 * n_slices;

  const uint32_t s = t * S;
  const uint32_t e = min(s + S, E);
  uint32_t g = find_bucket(bucket_start, NB, s);
  uint32_t gend = __ldg(bucket_start + g + 1);
  bool started_before = __ldg(bucket_start + g) < s;
  // carry_in (sub-batches after the first of a pipelined call): a bucket continues from the value the
  // earlier sub-batches left in bucket_acc; only the thread that owns the bucket's first entry reads it
  Xyzz<F> acc = (carry_in && !started_before) ? load_vec(&bucket_acc[g]) : xyzz_inf<F>();

  // The point of entry pos+1 is fetched (raw words, 128-bit loads) before the mixed addition of
  // entry pos starts, and the entry index one further ahead: the two dependent loads of the gather
  // are off the critical path of the IMAD chains.
  constexpr int WORDS = 2 * F::PACKED_WORDS;
  // entries == nullptr: the points to add are bases[s .. e) themselves, already signed (the output of the
  // affine halving rounds, bucket_affine.cuh)
  const bool direct = entries == nullptr;
  uint32_t nxt[WORDS];
  uint32_t ent = direct ? s : __ldg(entries + s);
  load_base_words<F>(bases + (ent & 0x7fffffffu), nxt);
  uint32_t ent_ahead = s + 1 < e ? (direct ? s + 1 : __ldg(entries + s + 1)) : ent;
  for (uint32_t pos = s; pos < e; pos++) {
    if (pos == gend) {
      // bucket g is complete: flush and move to the next non-empty bucket
      store_vec(started_before ? &partials[2 * t] : &bucket_acc[g], acc);
      started_before = false;
      do {
        g++;
        gend = __ldg(bucket_start + g + 1);
      } while (gend == pos);
      acc = carry_in ? load_vec(&bucket_acc[g]) : xyzz_inf<F>();
    }
    Affine<F> pt;
    const bool finite = unpack_base<F>(nxt, pt);
    const bool negate = (ent >> 31) != 0;
    ent = ent_ahead;
    if (pos + 1 < e) {
      load_base_words<F>(bases + (ent & 0x7fffffffu), nxt);
      if (pos + 2 < e) ent_ahead = direct ? pos + 2 : __ldg(entries + pos + 2);
    }
    if (finite) {
      pt = aff_cneg<F>(pt, negate);
      xyzz_madd<F>(acc, pt);
    }
  }
  if (gend == e) {
    store_vec(started_before ? &partials[2 * t] : &bucket_acc[g], acc);
  } else {
    store_vec(started_before ? &partials[2 * t] : &partials[2 * t + 1], acc);
    // bucket g starts in this slice and continues in the next one(s): exactly one thread sees that
    if (!started_before) cut_list[(size_t)line * cut_cap + atomicAdd(&cut_count[line], 1u)] = g;
  }
}

// Block-wide sum of one XYZZ value per thread (blockDim.x <= 128, power of two); result in thread 0.
template <class F> MSM_D Xyzz<F> block_sum_xyzz(Xyzz<F> v, Xyzz<F>* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (uint32_t stride = blockDim.x >> 1; stride >= 1; stride >>= 1) {
    if (threadIdx.x < stride) {
      v = xyzz_add<F>(v, sh[threadIdx.x + stride]);
      sh[threadIdx.x] = v;
    }
    __syncthreads();
  }
  return v;
}

// Buckets cut by slice boundaries: one thread per entry of the list k_accumulate appended to sums
// the bucket's partial slots (dense: every lane of a warp has the same work, unlike a
// thread-per-bucket sweep where ~30 % of the lanes take this branch).  Empty buckets need no
// visit: the bucket array is zero-filled (= infinity) before the first sub-batch.  Buckets spread
// over more than HEAVY_SPAN slices (skewed scalars; the short top window of a folded table) go to
// a second list that k_fixup_heavy reduces with one warp each.
constexpr uint32_t HEAVY_SPAN = 16;
constexpr uint32_t HEAVY_CHUNK = 256;  // partial slots one warp sums in k_fixup_heavy
// work lists of one line: [cut_count, heavy_count, chunk_count] then the arrays below
struct FixupLists {
  uint32_t* counts;        // [3 * n_lines]: cut, heavy, chunk counters of every line
  uint32_t* cut_list;      // [n_lines * cut_cap]
  uint32_t* heavy_list;    // [n_lines * heavy_cap] bucket id of every heavy bucket
  uint32_t* heavy_chunk0;  // [n_lines * heavy_cap] first chunk item of the bucket
  uint32_t* chunk_list;    // [n_lines * chunk_cap] heavy slot of every chunk item
  uint32_t cut_cap, heavy_cap, chunk_cap, n_lines;
};
template <class F>
__global__ void __launch_bounds__(128)
k_fixup_cut(const uint32_t* __restrict__ bucket_start, uint32_t NB, uint32_t S, uint32_t n_slices,
            Xyzz<F>* __restrict__ bucket_acc, const Xyzz<F>* __restrict__ partials, FixupLists fl) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t line = blockIdx.y;
  if (i >= fl.counts[line]) return;
  bucket_acc += (size_t)line * NB;
  partials += (size_t)line * 2 * n_slices;
  const uint32_t g = fl.cut_list[(size_t)line * fl.cut_cap + i];
  const uint32_t t0 = bucket_start[g] / S, t1 = (bucket_start[g + 1] - 1) / S;
  if (t1 - t0 > HEAVY_SPAN) {
    // spread over many slices: one warp per HEAVY_CHUNK partial slots (k_fixup_heavy), so that a
    // bucket holding a large share of all digits (equal or tiny scalars, a short top window) is
    // reduced by many warps and not by one
    const uint32_t nch = (t1 - t0 + 1 + HEAVY_CHUNK - 1) / HEAVY_CHUNK;
    const uint32_t slot = atomicAdd(&fl.counts[fl.n_lines + line], 1u);
    const uint32_t c0 = atomicAdd(&fl.counts[2 * fl.n_lines + line], nch);
    if (slot < fl.heavy_cap && c0 + nch <= fl.chunk_cap) {  // the caps cannot be exceeded (see make_plan)
      fl.heavy_list[(size_t)line * fl.heavy_cap + slot] = g;
      fl.heavy_chunk0[(size_t)line * fl.heavy_cap + slot] = c0;
      for (uint32_t q = 0; q < nch; q++) fl.chunk_list[(size_t)line * fl.chunk_cap + c0 + q] = slot;
    }
    return;
  }
  Xyzz<F> acc = load_vec(&partials[2 * t0 + 1]);
  for (uint32_t t = t0 + 1; t <= t1; t++) acc = xyzz_add<F>(acc, load_vec(&partials[2 * t]));
  store_vec(&bucket_acc[g], acc);
}

// sum of one value per lane over a warp (5-level tree through shared memory); result in lane 0
template <class F> MSM_D Xyzz<F> warp_sum_xyzz(Xyzz<F> acc, Xyzz<F>* sh, uint32_t lane) {
  sh[lane] = acc;
  __syncwarp();
  for (uint32_t stride = 16; stride >= 1; stride >>= 1) {
    if (lane < stride) {
      acc = xyzz_add<F>(acc, sh[lane + stride]);
      sh[lane] = acc;
    }
    __syncwarp();
  }
  return acc;
}

// One warp per chunk item: lane-strided sums of up to HEAVY_CHUNK partial slots of one heavy
// bucket, then a warp tree.  A bucket with a single chunk is finished here; otherwise the chunk
// sums go to chunk_out and k_fixup_heavy_final adds them up.
template <class F>
__global__ void __launch_bounds__(128)
k_fixup_heavy(const uint32_t* __restrict__ bucket_start, uint32_t NB, uint32_t S, uint32_t n_slices,
              Xyzz<F>* __restrict__ bucket_acc, const Xyzz<F>* __restrict__ partials, FixupLists fl,
              Xyzz<F>* __restrict__ chunk_out) {
  extern __shared__ uint4 smem_raw[];
  Xyzz<F>* sh = reinterpret_cast<Xyzz<F>*>(smem_raw) + (threadIdx.x & ~31u);
  const uint32_t line = blockIdx.y, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  bucket_acc += (size_t)line * NB;
  partials += (size_t)line * 2 * n_slices;
  chunk_out += (size_t)line * fl.chunk_cap;
  const uint32_t count = min(fl.counts[2 * fl.n_lines + line], fl.chunk_cap);
  for (uint32_t item = blockIdx.x * warps + (threadIdx.x >> 5); item < count; item += gridDim.x * warps) {
    const uint32_t slot = fl.chunk_list[(size_t)line * fl.chunk_cap + item];
    const uint32_t g = fl.heavy_list[(size_t)line * fl.heavy_cap + slot];
    const uint32_t q = item - fl.heavy_chunk0[(size_t)line * fl.heavy_cap + slot];
    const uint32_t t0 = bucket_start[g] / S, t1 = (bucket_start[g + 1] - 1) / S;
    // element 0 = slot 2*t0+1, element k = slot 2*(t0+k), k = 1 .. t1-t0
    const uint32_t k_lo = q * HEAVY_CHUNK, k_hi = min(k_lo + HEAVY_CHUNK, t1 - t0 + 1);
    Xyzz<F> acc = xyzz_inf<F>();
    for (uint32_t k = k_lo + lane; k < k_hi; k += 32)
      acc = xyzz_add<F>(acc, load_vec(k == 0 ? &partials[2 * t0 + 1] : &partials[2 * (t0 + k)]));
    acc = warp_sum_xyzz<F>(acc, sh, lane);
    if (lane == 0) store_vec(t1 - t0 + 1 <= HEAVY_CHUNK ? &bucket_acc[g] : &chunk_out[item], acc);
    __syncwarp();
  }
}
template <class F>
__global__ void __launch_bounds__(128)
k_fixup_heavy_final(const uint32_t* __restrict__ bucket_start, uint32_t NB, uint32_t S,
                    Xyzz<F>* __restrict__ bucket_acc, FixupLists fl, const Xyzz<F>* __restrict__ chunk_out) {
  extern __shared__ uint4 smem_raw[];
  Xyzz<F>* sh = reinterpret_cast<Xyzz<F>*>(smem_raw) + (threadIdx.x & ~31u);
  const uint32_t line = blockIdx.y, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  bucket_acc += (size_t)line * NB;
  chunk_out += (size_t)line * fl.chunk_cap;
  const uint32_t count = min(fl.counts[fl.n_lines + line], fl.heavy_cap);
  for (uint32_t slot = blockIdx.x * warps + (threadIdx.x >> 5); slot < count; slot += gridDim.x * warps) {
    const uint32_t g = fl.heavy_list[(size_t)line * fl.heavy_cap + slot];
    const uint32_t span = (bucket_start[g + 1] - 1) / S - bucket_start[g] / S + 1;
    if (span <= HEAVY_CHUNK) continue;  // finished by k_fixup_heavy
    const uint32_t nch = (span + HEAVY_CHUNK - 1) / HEAVY_CHUNK;
    const Xyzz<F>* src = chunk_out + fl.heavy_chunk0[(size_t)line * fl.heavy_cap + slot];
    Xyzz<F> acc = xyzz_inf<F>();
    for (uint32_t k = lane; k < nch; k += 32) acc = xyzz_add<F>(acc, load_vec(&src[k]));
    acc = warp_sum_xyzz<F>(acc, sh, lane);
    if (lane == 0) store_vec(&bucket_acc[g], acc);
    __syncwarp();
  }
}

// Plain sum of `count` consecutive points per group, 1024 per block: out[group][ceil(count/1024)].
template <class F>
__global__ void __launch_bounds__(128)
k_reduce_points(const Xyzz<F>* __restrict__ in, uint32_t count, uint32_t out_count, Xyzz<F>* __restrict__ out) {
  extern __shared__ uint4 smem_raw[];
  Xyzz<F>* sh = reinterpret_cast<Xyzz<F>*>(smem_raw);
  const uint32_t group = blockIdx.y, chunk = blockIdx.x;
  const Xyzz<F>* src = in + (size_t)group * count;
  const uint32_t lo = chunk * 1024, hi = min(lo + 1024, count);
  Xyzz<F> acc = xyzz_inf<F>();
  for (uint32_t k = lo + threadIdx.x; k < hi; k += blockDim.x) acc = xyzz_add<F>(acc, load_vec(&src[k]));
  acc = block_sum_xyzz<F>(acc, sh);
  if (threadIdx.x == 0) store_vec(&out[(size_t)group * out_count + chunk], acc);
}

// ---------------------------------------------------------------------------------------------
// Bucket reduction: for every group (line, task, window) compute sum_{b=1..B} b * S_b.
// Thread j owns Q consecutive buckets: local running sum, plus (first_weight-1) * (plain sum);
// then a segmented shared-memory tree over RW threads.  Output: one partial per RW threads.
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128)
k_bucket_reduce(const Xyzz<F>* __restrict__ bucket_acc, uint32_t n_threads, uint32_t B, uint32_t Q,
                uint32_t RW, Xyzz<F>* __restrict__ out) {
  extern __shared__ uint4 smem_raw[];
  Xyzz<F>* sh = reinterpret_cast<Xyzz<F>*>(smem_raw);
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  Xyzz<F> res = xyzz_inf<F>();
  if (tid < n_threads) {
    const size_t f = (size_t)tid * Q;
    const uint32_t b0 = (uint32_t)(f % B);  // weight of bucket f+k is b0 + k + 1
    Xyzz<F> run = xyzz_inf<F>();
    for (int k = (int)Q - 1; k >= 0; k--) {
      run = xyzz_add<F>(run, load_vec(&bucket_acc[f + k]));
      res = xyzz_add<F>(res, run);
    }
    if (b0) res = xyzz_add<F>(res, xyzz_mul_small<F>(run, b0));
  }
  sh[threadIdx.x] = res;
  __syncthreads();
  const uint32_t lane = threadIdx.x & (RW - 1);
  for (uint32_t stride = RW >> 1; stride >= 1; stride >>= 1) {
    if (lane < stride) {
      res = xyzz_add<F>(res, sh[threadIdx.x + stride]);
      sh[threadIdx.x] = res;
    }
    __syncthreads();
  }
  if (lane == 0 && tid < n_threads) store_vec(&out[tid / RW], res);
}

// One block per task (line, chunk): thread w sums the PG partials of window w; thread 0 folds the
// windows Horner-style (c doublings per step) and writes the Jacobian result in the API layout.
template <class F>
__global__ void k_window_combine(const Xyzz<F>* __restrict__ group_partials, uint32_t W, uint32_t PG,
                                 uint32_t c, ApiJacobian<F>* __restrict__ out) {
  extern __shared__ uint4 smem_raw[];
  Xyzz<F>* sh = reinterpret_cast<Xyzz<F>*>(smem_raw);
  const uint32_t task = blockIdx.x;
  for (uint32_t w = threadIdx.x; w < W; w += blockDim.x) {
    const Xyzz<F>* src = group_partials + ((size_t)task * W + w) * PG;
    Xyzz<F> s = load_vec(src);
    for (uint32_t k = 1; k < PG; k++) s = xyzz_add<F>(s, load_vec(src + k));
    sh[w] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Xyzz<F> acc = sh[W - 1];
    for (int w = (int)W - 2; w >= 0; w--) {
      for (uint32_t k = 0; k < c; k++) acc = xyzz_dbl<F>(acc);
      acc = xyzz_add<F>(acc, sh[w]);
    }
    xyzz_to_api_jacobian<F>(acc, &out[task]);
  }
}

// Many small tasks (batched per-segment commitments, the AMT shape): one THREAD per task does the
// same fold, so a warp finishes 32 tasks in the time the block-per-task form finishes one.
template <class F>
__global__ void __launch_bounds__(128)
k_window_combine_batched(const Xyzz<F>* __restrict__ group_partials, uint32_t n_tasks, uint32_t W, uint32_t PG,
                         uint32_t c, ApiJacobian<F>* __restrict__ out) {
  const uint32_t task = blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= n_tasks) return;
  Xyzz<F> acc = xyzz_inf<F>();
  for (int w = (int)W - 1; w >= 0; w--) {
    if (w != (int)W - 1)
      for (uint32_t k = 0; k < c; k++) acc = xyzz_dbl<F>(acc);
    const Xyzz<F>* src = group_partials + ((size_t)task * W + w) * PG;
    for (uint32_t k = 0; k < PG; k++) acc = xyzz_add<F>(acc, load_vec(src + k));
  }
  xyzz_to_api_jacobian<F>(acc, &out[task]);
}

// ---------------------------------------------------------------------------------------------
// Boundary and helper kernels: resident-copy conversion, sum of Jacobian points, Jacobian ->
// affine, per-primitive test kernels (counterpart of ag-build/cl/test.cl), synthetic inputs.
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void k_convert_bases(const ApiAffine<F>* __restrict__ in, uint32_t n, PackedAffine<F>* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  ApiAffine<F> a = in[i];
  PackedAffine<F> o;
  F::api_to_packed(a.x, o.x);  // identity (0,0) stays all-zero
  F::api_to_packed(a.y, o.y);
  out[i] = o;
}

// Window table for resident bases: T[w][i] = 2^(c w) * P_i in the packed affine layout, w < W.
// One thread per point: (W-1)*c doublings in XYZZ, then one shared inversion (Montgomery's trick
// over the thread's W-1 results).  With the table every window of a large MSM lands in ONE bucket
// set: no per-window bucket arrays and no Horner doublings at the end of each call.
constexpr int TABLE_MAX_W = 32;
template <class F>
__global__ void __launch_bounds__(64)
k_build_tables(const PackedAffine<F>* __restrict__ bases, uint32_t n, uint32_t c, uint32_t W,
               PackedAffine<F>* __restrict__ table) {
  using E = typename F::Elem;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PackedAffine<F> src = bases[i];
  table[i] = src;
  Affine<F> p;
  p.x = F::unpack(src.x);
  p.y = F::unpack(src.y);
  PackedAffine<F> zero;
  for (int k = 0; k < F::PACKED_WORDS; k++) zero.x[k] = zero.y[k] = 0;
  if (aff_is_identity<F>(p)) {
    for (uint32_t w = 1; w < W; w++) table[(size_t)w * n + i] = zero;
    return;
  }
  Xyzz<F> cur;
  cur.x = p.x;
  cur.y = p.y;
  cur.zz = F::one();
  cur.zzz = F::one();
  Xyzz<F> pts[TABLE_MAX_W];
  E pref[TABLE_MAX_W];
  E prod = F::one();
  for (uint32_t w = 1; w < W; w++) {
    for (uint32_t k = 0; k < c; k++) cur = xyzz_dbl<F>(cur);
    pts[w] = cur;
    pref[w] = prod;
    if (!xyzz_is_inf<F>(cur)) prod = F::mul(prod, F::mul(cur.zz, cur.zzz));
  }
  E inv = F::inv(prod);
  for (uint32_t w = W - 1; w >= 1; w--) {
    if (xyzz_is_inf<F>(pts[w])) {
      table[(size_t)w * n + i] = zero;
      continue;
    }
    const E zi = F::mul(inv, pref[w]);  // 1 / (zz*zzz)
    inv = F::mul(inv, F::mul(pts[w].zz, pts[w].zzz));
    PackedAffine<F> o;
    F::to_packed(F::mul(pts[w].x, F::mul(zi, pts[w].zzz)), o.x);
    F::to_packed(F::mul(pts[w].y, F::mul(zi, pts[w].zz)), o.y);
    table[(size_t)w * n + i] = o;
  }
}

template <class F>
__global__ void k_sum_points(const ApiJacobian<F>* __restrict__ in, uint32_t count, ApiJacobian<F>* __restrict__ out) {
  if (blockIdx.x || threadIdx.x) return;
  Xyzz<F> acc = xyzz_inf<F>();
  for (uint32_t i = 0; i < count; i++) acc = xyzz_add<F>(acc, xyzz_from_api_jacobian<F>(&in[i]));
  xyzz_to_api_jacobian<F>(acc, &out[0]);
}

template <class F>
__global__ void k_to_affine(const ApiJacobian<F>* __restrict__ in, uint32_t count, int mont_out,
                            ApiAffine<F>* __restrict__ out, uint8_t* __restrict__ is_inf) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Xyzz<F> x = xyzz_from_api_jacobian<F>(&in[i]);
  is_inf[i] = xyzz_to_api_affine<F>(x, mont_out != 0, &out[i]) ? 1 : 0;
}

template <class F> struct ApiElem {
  uint32_t w[F::API_WORDS];
};

// r2 / unit: API words of R^2 mod p and of the integer 1 (for the mont / unmont test ops)
template <class F>
__global__ void k_test_fq(int op, const ApiElem<F>* __restrict__ a, const ApiElem<F>* __restrict__ b,
                          ApiElem<F> r2, ApiElem<F>* __restrict__ o, uint32_t count) {
  using E = typename F::Elem;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  ApiElem<F> wa = a[i], wb = wa, wo;
  if (b) wb = b[i];
  E x = F::norm(F::from_api(wa.w)), y = F::norm(F::from_api(wb.w)), r;
  switch (op) {
    case 0: r = F::add(x, y); break;
    case 1: r = F::template sub<2, 1>(x, y); break;
    case 2: r = F::mul(x, y); break;
    case 3: r = F::sqr(x); break;
    case 4: r = F::add(x, x); break;
    case 5: r = F::mul(x, F::from_api(r2.w)); break;                 // to Montgomery form
    case 6: {                                                         // from Montgomery form
      ApiElem<F> unit;
      for (int k = 0; k < F::API_WORDS; k++) unit.w[k] = k == 0 ? 1u : 0u;
      r = F::mul(x, F::from_api(unit.w));
      break;
    }
    case 7: r = F::inv(x); break;
    default: r = F::template neg<2, 1>(x); break;
  }
  F::to_api(r, wo.w);
  o[i] = wo;
}

template <class F>
__global__ void k_test_ec(int op, const ApiJacobian<F>* __restrict__ a, const void* __restrict__ b,
                          ApiJacobian<F>* __restrict__ o, uint32_t count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  Xyzz<F> x = xyzz_from_api_jacobian<F>(&a[i]), r;
  if (op == 0) {
    r = xyzz_add<F>(x, xyzz_from_api_jacobian<F>(reinterpret_cast<const ApiJacobian<F>*>(b) + i));
  } else if (op == 1) {
    Affine<F> q = affine_from_api<F>(reinterpret_cast<const ApiAffine<F>*>(b) + i);
    r = x;
    if (!aff_is_identity<F>(q)) xyzz_madd<F>(r, q);
  } else {
    r = xyzz_dbl<F>(x);
  }
  xyzz_to_api_jacobian<F>(r, &o[i]);
}

// Exponents handed over in Montgomery form (arkworks' in-memory Fr) -> canonical integers, the
// device-side replacement of the host pass PrimeFieldRepr::to_bigint (ag-types/src/impls.rs:7-18,
// "10ms for 1M" in ag-cuda-ec/benches/multiexp.rs:28-36).  In place is allowed.
template <class PR>
__global__ void k_scalars_unmont(const uint32_t* __restrict__ in, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp<PR> a;
  uint32_t k[8];
  load_scalar(in, i, k);
#pragma unroll
  for (int j = 0; j < 8; j++) a.v[j] = k[j];
  const Fp<PR> c = fp_from_mont<PR>(a);
  uint4* o = reinterpret_cast<uint4*>(out) + 2 * (size_t)i;
  o[0] = make_uint4(c.v[0], c.v[1], c.v[2], c.v[3]);
  o[1] = make_uint4(c.v[4], c.v[5], c.v[6], c.v[7]);
}

// --- synthetic inputs -------------------------------------------------------------------------
MSM_HD uint64_t splitmix64(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct ScalarField {
  uint32_t r[8];  // modulus, little-endian
  uint32_t bits;
};

static __global__ void k_synth_scalars(ScalarField fr, uint64_t seed, uint64_t start, uint32_t n,
                                uint32_t* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t top_mask = (fr.bits % 64) ? ((1ull << (fr.bits % 64)) - 1) : ~0ull;
  uint64_t k[4];
  for (uint64_t attempt = 0;; attempt++) {
    const uint64_t s = seed + attempt * 0xD1B54A32D192ED03ull;
#pragma unroll
    for (int j = 0; j < 4; j++) k[j] = splitmix64(s, 4 * (start + i) + j);
    k[3] &= top_mask;
    // k < r ?
    bool lt = false, decided = false;
#pragma unroll
    for (int j = 3; j >= 0; j--) {
      const uint64_t rj = ((uint64_t)fr.r[2 * j + 1] << 32) | fr.r[2 * j];
      if (!decided && k[j] != rj) {
        lt = k[j] < rj;
        decided = true;
      }
    }
    if (lt) break;
  }
  uint4* o = reinterpret_cast<uint4*>(out) + 2 * (size_t)i;
  o[0] = make_uint4((uint32_t)k[0], (uint32_t)(k[0] >> 32), (uint32_t)k[1], (uint32_t)(k[1] >> 32));
  o[1] = make_uint4((uint32_t)k[2], (uint32_t)(k[2] >> 32), (uint32_t)k[3], (uint32_t)(k[3] >> 32));
}

// P_i = (a + (start+i) b) G.  Each thread produces RUN consecutive points: one double-and-add for
// its first point, RUN-1 mixed additions of D = b*G, then one shared inversion (Montgomery's
// trick) to normalise them.  gen / d are affine in the API layout; so is the output.
constexpr int SYNTH_RUN = 16;
template <class F>
__global__ void __launch_bounds__(64)
k_synth_points(ApiAffine<F> gen_api, ApiAffine<F> d_api, uint64_t a, uint64_t b, uint64_t start, uint32_t n,
               ApiAffine<F>* __restrict__ out) {
  using E = typename F::Elem;
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t i0 = (uint64_t)t * SYNTH_RUN;
  if (i0 >= n) return;
  const uint32_t cnt = (uint32_t)min((uint64_t)SYNTH_RUN, (uint64_t)n - i0);
  const Affine<F> gen = affine_from_api<F>(&gen_api), d = affine_from_api<F>(&d_api);
  // k = a + (start + i0) * b, up to 129 bits
  const uint64_t m = start + i0;
  const uint64_t lo = m * b, hi = __umul64hi(m, b);
  uint64_t k0 = lo + a;
  uint64_t c0 = k0 < lo ? 1 : 0;
  uint64_t k1 = hi + c0;
  uint64_t k2 = k1 < hi ? 1 : 0;
  Xyzz<F> acc = xyzz_inf<F>();
  for (int bit = 128; bit >= 0; bit--) {
    acc = xyzz_dbl<F>(acc);
    const uint64_t word = bit >= 128 ? k2 : (bit >= 64 ? k1 : k0);
    if ((word >> (bit & 63)) & 1) xyzz_madd<F>(acc, gen);
  }
  Xyzz<F> pts[SYNTH_RUN];
  E pref[SYNTH_RUN];
  E prod = F::one();
  for (uint32_t j = 0; j < cnt; j++) {
    pts[j] = acc;
    pref[j] = prod;
    prod = F::mul(prod, F::mul(acc.zz, acc.zzz));  // never infinity: k < group order
    xyzz_madd<F>(acc, d);
  }
  E inv = F::inv(prod);
  for (int j = (int)cnt - 1; j >= 0; j--) {
    const E zi = F::mul(inv, pref[j]);  // 1 / (zz*zzz)
    inv = F::mul(inv, F::mul(pts[j].zz, pts[j].zzz));
    ApiAffine<F> o;
    F::to_api(F::mul(pts[j].x, F::mul(zi, pts[j].zzz)), o.x);
    F::to_api(F::mul(pts[j].y, F::mul(zi, pts[j].zz)), o.y);
    out[i0 + j] = o;
  }
}

}  // namespace msm
