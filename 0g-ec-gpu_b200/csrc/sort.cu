// sort.cu -- signed-digit decomposition and the counting sorts of the (window, bucket, point) digits.
// Field-independent: compiled once (see sort.h).  Kernel descriptions: kernels.cuh header comment.
#include "sort.h"
#include "kernels.cuh"

namespace msm {

// Scalar i as canonical little-endian words.  geo.mont != 0: the row holds Fr elements in Montgomery form and the
// conversion PrimeFieldRepr::to_bigint does on the host (ag-types/src/impls.rs:7-18; FIELD_unmont,
// ag-build/cl/field.cl:365-377) happens here, fused into every decomposition pass (one Fr product per read, no
// extra pass over the row and no second copy of it).
MSM_D void load_scalar_geo(const uint32_t* scalars, uint32_t i, const Geometry& geo, uint32_t k[8]) {
  load_scalar(scalars, i, k);
  if (geo.mont == 1) {
    Fp<Bn254Fr> a;
#pragma unroll
    for (int j = 0; j < 8; j++) a.v[j] = k[j];
    a = fp_from_mont<Bn254Fr>(a);
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = a.v[j];
  } else if (geo.mont == 2) {
    Fp<Bls381Fr> a;
#pragma unroll
    for (int j = 0; j < 8; j++) a.v[j] = k[j];
    a = fp_from_mont<Bls381Fr>(a);
#pragma unroll
    for (int j = 0; j < 8; j++) k[j] = a.v[j];
  }
}

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
  load_scalar_geo(scalars, i, geo, k);
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
constexpr int PART_ITEMS = 2;  // scalars per thread when the window size is a template argument (tile = 1024)

// The decomposition of for_each_digit_c into a register array: dg[w] = (bucket - 1) | negative << 31, or
// 0xffffffff for a zero digit.  k_partition decomposes every scalar ONCE and keeps its digits in registers
// between the histogram pass and the placement pass (it used to read and decompose the row twice).
template <int C> MSM_D void decompose_c(const uint32_t k[8], uint32_t W, uint32_t* dg) {
  constexpr uint32_t half = 1u << (C - 1);
  constexpr uint32_t mask = (1u << C) - 1u;
  constexpr int MAXW = (256 + C - 1) / C + 1;
  uint32_t carry = 0;
#pragma unroll
  for (int w = 0; w < MAXW; w++) {
    dg[w] = 0xffffffffu;
    if ((uint32_t)w < W) {
      const int bit = w * C, word = bit >> 5, sh = bit & 31;
      const uint32_t lo = word < 8 ? k[word < 8 ? word : 0] : 0u;
      const uint32_t hi = word + 1 < 8 ? k[word + 1 < 8 ? word + 1 : 0] : 0u;
      const uint32_t raw = (__funnelshift_r(lo, hi, sh) & mask) + carry;
      carry = raw > half;
      if (raw != 0 && raw != (1u << C)) dg[w] = carry ? (((1u << C) - raw - 1u) | 0x80000000u) : (raw - 1u);
    }
  }
}

template <int C>
__global__ void __launch_bounds__(PART_BLOCK)
k_partition(const uint32_t* __restrict__ scalars, Geometry geo, uint32_t tile, uint32_t bin_shift, uint32_t n_bins,
            const uint32_t* __restrict__ region_start, uint32_t region_shift, uint32_t* __restrict__ bin_cursor,
            uint2* __restrict__ tmp) {
  extern __shared__ uint32_t part_smem[];
  uint32_t* hist = part_smem;                 // [n_bins] counts, then running cursors
  uint32_t* off = hist + n_bins;              // [n_bins] exclusive offsets inside the block
  uint32_t* gbase = off + n_bins;             // [n_bins] global base of this block's run
  uint2* stage = reinterpret_cast<uint2*>(gbase + n_bins + (n_bins & 1));  // [tile * W] (bucket id, entry), 8-byte aligned
  __shared__ uint32_t total_sh;
  const uint32_t first = blockIdx.x * tile;
  for (uint32_t b = threadIdx.x; b < n_bins; b += PART_BLOCK) hist[b] = 0;
  __syncthreads();
  constexpr int CC = C == 0 ? 8 : C;
  constexpr int MAXW = (256 + CC - 1) / CC + 1;
  uint32_t dg[C == 0 ? 1 : PART_ITEMS][C == 0 ? 1 : MAXW];
  // pass 1: histogram of high bins
  if (C == 0) {
    for (uint32_t t = threadIdx.x; t < tile; t += PART_BLOCK) {
      const uint32_t i = first + t;
      if (i >= geo.L) break;
      uint32_t k[8];
      load_scalar_geo(scalars, i, geo, k);
      const uint32_t base = (i / geo.chunk_len) * geo.W;
      for_each_digit(k, geo.c, geo.W, [&](uint32_t w, uint32_t bucket, bool) {
        const uint32_t g = geo.fold ? task_of(i, geo) * geo.B + (bucket - 1) : (base + w) * geo.B + (bucket - 1);
        atomicAdd(&hist[g >> bin_shift], 1u);
      });
    }
  } else {
#pragma unroll
    for (int it = 0; it < PART_ITEMS; it++) {
      const uint32_t i = first + it * PART_BLOCK + threadIdx.x;
#pragma unroll
      for (int w = 0; w < MAXW; w++) dg[it][w] = 0xffffffffu;
      if (it * PART_BLOCK + threadIdx.x < tile && i < geo.L) {
        uint32_t k[8];
        load_scalar_geo(scalars, i, geo, k);
        decompose_c<CC>(k, geo.W, dg[it]);
        const uint32_t gb = geo.fold ? task_of(i, geo) * geo.B : (i / geo.chunk_len) * geo.W * geo.B;
#pragma unroll
        for (int w = 0; w < MAXW; w++)
          if (dg[it][w] != 0xffffffffu) {
            const uint32_t g = gb + (geo.fold ? 0u : (uint32_t)w * geo.B) + (dg[it][w] & 0x7fffffffu);
            atomicAdd(&hist[g >> bin_shift], 1u);
          }
      }
    }
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
  // pass 2: place (g, entry) in the block-local bin order
  if (C == 0) {
    for (uint32_t t = threadIdx.x; t < tile; t += PART_BLOCK) {
      const uint32_t i = first + t;
      if (i >= geo.L) break;
      uint32_t k[8];
      load_scalar_geo(scalars, i, geo, k);
      const uint32_t base = (i / geo.chunk_len) * geo.W;
      for_each_digit(k, geo.c, geo.W, [&](uint32_t w, uint32_t bucket, bool neg) {
        const uint32_t g = geo.fold ? task_of(i, geo) * geo.B + (bucket - 1) : (base + w) * geo.B + (bucket - 1);
        const uint32_t hb = g >> bin_shift;
        const uint32_t slot = off[hb] + atomicAdd(&hist[hb], 1u);
        const uint32_t idx = geo.fold ? w * geo.table_stride + geo.point_offset + i : i;
        stage[slot] = make_uint2(g, idx | (neg ? 0x80000000u : 0u));
      });
    }
  } else {
#pragma unroll
    for (int it = 0; it < PART_ITEMS; it++) {
      const uint32_t i = first + it * PART_BLOCK + threadIdx.x;
      const uint32_t gb = geo.fold ? task_of(i, geo) * geo.B : (i / geo.chunk_len) * geo.W * geo.B;
#pragma unroll
      for (int w = 0; w < MAXW; w++)
        if (dg[it][w] != 0xffffffffu) {
          const uint32_t g = gb + (geo.fold ? 0u : (uint32_t)w * geo.B) + (dg[it][w] & 0x7fffffffu);
          const uint32_t hb = g >> bin_shift;
          const uint32_t slot = off[hb] + atomicAdd(&hist[hb], 1u);
          const uint32_t idx = geo.fold ? (uint32_t)w * geo.table_stride + geo.point_offset + i : i;
          stage[slot] = make_uint2(g, idx | (dg[it][w] & 0x80000000u));
        }
    }
  }
  __syncthreads();
  // write every bin's run to its region: consecutive slots of one bin are consecutive 8-byte pairs in memory
  const uint32_t total = total_sh;
  for (uint32_t sidx = threadIdx.x; sidx < total; sidx += PART_BLOCK) {
    const uint2 pr = stage[sidx];
    const uint32_t hb = pr.x >> bin_shift;
    tmp[gbase[hb] + (sidx - off[hb])] = pr;
  }
}

static __global__ void k_final_scatter(const uint2* __restrict__ tmp,
                                       const uint32_t* __restrict__ E_ptr, uint32_t* __restrict__ cursor,
                                       uint32_t* __restrict__ entries) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= __ldg(E_ptr)) return;
  const uint2 pr = __ldg(tmp + i);
  entries[atomicAdd(&cursor[pr.x], 1u)] = pr.y;
}

template <int C> inline cudaError_t partition_set_smem(size_t bytes) {
  return cudaFuncSetAttribute(k_partition<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
inline cudaError_t launch_partition(uint32_t grid, size_t smem, cudaStream_t st, const uint32_t* scalars,
                                    const Geometry& geo, uint32_t tile, uint32_t bin_shift, uint32_t n_bins,
                                    const uint32_t* region_start, uint32_t region_shift, uint32_t* bin_cursor,
                                    uint2* tmp) {
  cudaError_t e = cudaSuccess;
#define MSM_PART_CASE(CC)                                                                                        \
  case CC:                                                                                                       \
    e = partition_set_smem<CC>(smem);                                                                            \
    if (e == cudaSuccess)                                                                                        \
      k_partition<CC><<<grid, PART_BLOCK, smem, st>>>(scalars, geo, tile, bin_shift, n_bins, region_start,       \
                                                      region_shift, bin_cursor, tmp);                            \
    break;
  switch (geo.c) {
    MSM_PART_CASE(16) MSM_PART_CASE(17) MSM_PART_CASE(18) MSM_PART_CASE(19) MSM_PART_CASE(20)
    MSM_PART_CASE(21) MSM_PART_CASE(22) MSM_PART_CASE(23) MSM_PART_CASE(24)
    default:
      e = partition_set_smem<0>(smem);
      if (e == cudaSuccess)
        k_partition<0><<<grid, PART_BLOCK, smem, st>>>(scalars, geo, tile, bin_shift, n_bins, region_start, region_shift,
                                                       bin_cursor, tmp);
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
    load_scalar_geo(scalars, i, geo, k);
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
k_bin_hist(const uint2* __restrict__ tmp, const uint32_t* __restrict__ bin_start,
           const uint32_t* __restrict__ tile_start, uint32_t n_bins, uint32_t bin_shift, uint32_t NB,
           uint32_t* __restrict__ counts) {
  extern __shared__ uint32_t bin_smem[];
  const uint32_t bpb = 1u << bin_shift;
  uint32_t bin, lo, hi;
  if (!bin_tile_range(bin_start, tile_start, n_bins, blockIdx.x, bin, lo, hi)) return;
  for (uint32_t b = threadIdx.x; b < bpb; b += BIN_BLOCK) bin_smem[b] = 0;
  __syncthreads();
  for (uint32_t p = lo + threadIdx.x; p < hi; p += BIN_BLOCK) atomicAdd(&bin_smem[__ldg(&tmp[p].x) & (bpb - 1)], 1u);
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
k_bin_place(const uint2* __restrict__ tmp,
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
  for (uint32_t p = lo + threadIdx.x; p < hi; p += PLACE_BLOCK) atomicAdd(&hist[__ldg(&tmp[p].x) & (bpb - 1)], 1u);
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
    const uint2 pr = __ldg(tmp + p);
    const uint32_t lb = pr.x & (bpb - 1);
    const uint32_t slot = off[lb] + atomicAdd(&hist[lb], 1u);
    stage_v[slot] = pr.y;
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


// counts of the next halving round from the offsets of this one: n_out[g] = ceil(n_in[g] / 2)
static __global__ void k_halve_counts(const uint32_t* __restrict__ off_in, uint32_t NB, uint32_t* __restrict__ counts) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < NB) counts[g] = (__ldg(off_in + g + 1) - __ldg(off_in + g) + 1) >> 1;
}

// ---------------------------------------------------------------------------------------------
// The three sorts (kernels.cuh).  All leave bucket_start[NB+1] (exclusive scan of the bucket sizes,
// bucket_start[NB] = number of non-zero digits) and the entries in bucket order.  sg is the geometry
// of the (sub-)batch, E_max its digit bound; temporaries come from the scratch arena.
// ---------------------------------------------------------------------------------------------

void enqueue_bucket_scan(cudaStream_t st, const Geometry& g, const SortBuffers& b) {
  const uint32_t n_tiles = (g.NB + SCAN_TILE - 1) / SCAN_TILE;
  k_scan_tiles<<<n_tiles, SCAN_BLOCK, 0, st>>>(b.counts, g.NB, b.bucket_start, b.tile_sums);
  k_scan_tile_sums<<<1, SCAN_BLOCK, 0, st>>>(b.tile_sums, n_tiles, b.tile_sums + n_tiles);
  k_scan_finish<<<(g.NB + 1 + 255) / 256, 256, 0, st>>>(b.bucket_start, g.NB, b.tile_sums, b.tile_sums + n_tiles, b.cursor);
}

// digits staged per k_partition block (8 bytes each in shared memory): measured best of 6144 .. 12288
static uint32_t partition_tile(uint32_t W) {
  uint32_t part_entries = 12288;
  if (const char* env = getenv("MSM_B200_PART_ENTRIES")) part_entries = (uint32_t)atoi(env);
  const uint32_t tile = part_entries / W;
  return tile > 1024 ? 1024 : (tile < 64 ? 64 : tile);
}

// Binned sort: every per-digit atomic in shared memory (large calls).
int enqueue_sort_binned(msm_ctx* ctx, DeviceCtx& dc, const Plan& pl, const Geometry& sg, uint64_t E_max,
                               const uint32_t* scalars, const SortBuffers& b, cudaStream_t on) {
  cudaStream_t st = on ? on : dc.stream;
  CU_TRY(ctx, cudaMemsetAsync(b.counts, 0, (size_t)(sg.NB + 1) * 4, st));
  if (sg.L == 0) {
    enqueue_bucket_scan(st, sg, b);
    return MSM_OK;
  }
  const size_t cap = (size_t)pl.sub_max * sg.W + sg.W;
  uint2* tmp = dc.arena.take<uint2>(cap);  // (bucket id, entry) pairs grouped by bin
  uint32_t* bin_count = dc.arena.take<uint32_t>(1024);  // [bin_count | bin_cursor]: one memset
  uint32_t* bin_cursor = dc.arena.take<uint32_t>(1024);
  uint32_t* bin_start = dc.arena.take<uint32_t>(1025);
  uint32_t* tile_start = dc.arena.take<uint32_t>(1025);
  const uint32_t bin_shift = pl.bin_shift;
  const uint32_t n_bins = (uint32_t)(((uint64_t)sg.NB + (1ull << bin_shift) - 1) >> bin_shift);
  CU_TRY(ctx, cudaMemsetAsync(bin_count, 0, (size_t)((char*)bin_start - (char*)bin_count), st));
  launch_bin_count((sg.L + BIN_COUNT_SCALARS - 1) / BIN_COUNT_SCALARS, st, scalars, sg, bin_shift, n_bins, bin_count);
  k_bin_scan<<<1, SCAN_BLOCK, 0, st>>>(bin_count, n_bins, bin_start, tile_start);
  const uint32_t tile = partition_tile(sg.W);
  const size_t smem = ((size_t)3 * n_bins + (size_t)2 * tile * sg.W) * 4 + 8;
  CU_TRY(ctx, launch_partition((sg.L + tile - 1) / tile, smem, st, scalars, sg, tile, bin_shift, n_bins, bin_start, 0u,
                               bin_cursor, tmp));
  const uint32_t max_tiles = (uint32_t)(E_max / BIN_TILE) + n_bins + 1;
  k_bin_hist<<<max_tiles, BIN_BLOCK, (size_t)4 << bin_shift, st>>>(tmp, bin_start, tile_start, n_bins, bin_shift, sg.NB,
                                                                  b.counts);
  enqueue_bucket_scan(st, sg, b);
  const size_t psmem = bin_place_smem(bin_shift);
  CU_TRY(ctx, cudaFuncSetAttribute(k_bin_place, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
  k_bin_place<<<max_tiles, PLACE_BLOCK, psmem, st>>>(tmp, bin_start, tile_start, n_bins, bin_shift, sg.NB, b.cursor,
                                                    b.entries);
  pl.scatter_passes = 0;
  dc.launches += 5;
  return MSM_OK;
}

// Two-level scatter with global cursor atomics in the second level (MSM_B200_PARTITION=1: measured slower
// than both other sorts, kept as evidence).
int enqueue_sort_partition(msm_ctx* ctx, DeviceCtx& dc, const Plan& pl, const Geometry& sg, uint64_t E_max,
                                  const uint32_t* scalars, const SortBuffers& b, cudaStream_t on) {
  cudaStream_t st = on ? on : dc.stream;
  CU_TRY(ctx, cudaMemsetAsync(b.counts, 0, (size_t)(sg.NB + 1) * 4, st));
  const uint32_t db = 256, dg = (sg.L + db - 1) / db;
  if (dg) launch_digits<false>(dg, db, st, scalars, sg, b.counts, nullptr, 0u, sg.NB);
  enqueue_bucket_scan(st, sg, b);
  if (!dg) return MSM_OK;
  const size_t cap = (size_t)pl.sub_max * sg.W + sg.W;
  uint2* tmp = dc.arena.take<uint2>(cap);  // (bucket id, entry) pairs grouped by bin
  uint32_t* bin_cursor = dc.arena.take<uint32_t>(4096);
  uint32_t bins = 16;
  while (bins < 1024 && (uint64_t)bins * (4u << 20) < E_max * 4) bins <<= 1;  // ~4 MB of entries per bin
  uint32_t nb_log = 0;
  while ((1ull << nb_log) < sg.NB) nb_log++;
  uint32_t bins_log = 0;
  while ((1u << bins_log) < bins) bins_log++;
  const uint32_t bin_shift = nb_log > bins_log ? nb_log - bins_log : 0;
  const uint32_t n_bins = (uint32_t)(((uint64_t)sg.NB + (1ull << bin_shift) - 1) >> bin_shift);
  const uint32_t tile = partition_tile(sg.W);
  const size_t smem = ((size_t)3 * n_bins + (size_t)2 * tile * sg.W) * 4 + 8;
  CU_TRY(ctx, cudaMemsetAsync(bin_cursor, 0, (size_t)n_bins * 4, st));
  CU_TRY(ctx, launch_partition((sg.L + tile - 1) / tile, smem, st, scalars, sg, tile, bin_shift, n_bins, b.bucket_start,
                               bin_shift, bin_cursor, tmp));
  k_final_scatter<<<(uint32_t)((E_max + 255) / 256), 256, 0, st>>>(tmp, b.bucket_start + sg.NB, b.cursor, b.entries);
  pl.scatter_passes = 0;
  dc.launches += 2;
  return MSM_OK;
}

// Single-level sort: one L2 atomic per digit in the histogram and in the scatter (small calls).  The
// scatter runs in bucket-range passes: each pass writes a bounded slice of `entries` at random, which the
// 126 MB L2 partly absorbs; every pass re-reads the scalars (sequential).
int enqueue_sort_atomic(msm_ctx* ctx, DeviceCtx& dc, const Plan& pl, const Geometry& sg, uint64_t E_max,
                               const uint32_t* scalars, const SortBuffers& b, cudaStream_t on) {
  cudaStream_t st = on ? on : dc.stream;
  CU_TRY(ctx, cudaMemsetAsync(b.counts, 0, (size_t)(sg.NB + 1) * 4, st));
  const uint32_t db = 256, dg = (sg.L + db - 1) / db;
  if (dg) launch_digits<false>(dg, db, st, scalars, sg, b.counts, nullptr, 0u, sg.NB);
  enqueue_bucket_scan(st, sg, b);
  uint32_t passes = (uint32_t)((E_max * 4 + (200u << 20) - 1) / (200u << 20));
  if (const char* env = getenv("MSM_B200_SCATTER_PASSES")) passes = (uint32_t)atoi(env);
  passes = passes < 1 ? 1 : (passes > 8 ? 8 : passes);
  pl.scatter_passes = passes;
  const uint32_t per = (sg.NB + passes - 1) / passes;
  for (uint32_t ps = 0; ps < passes && dg; ps++) {
    const uint32_t lo = ps * per, hi = lo + per < sg.NB ? lo + per : sg.NB;
    if (lo >= hi) break;
    launch_digits<true>(dg, db, st, scalars, sg, b.cursor, b.entries, lo, hi);
    dc.launches += 1;
  }
  return MSM_OK;
}


void enqueue_halve_scan(cudaStream_t st, const Geometry& g, const uint32_t* off_in, const SortBuffers& b) {
  k_halve_counts<<<(g.NB + 255) / 256, 256, 0, st>>>(off_in, g.NB, b.counts);
  enqueue_bucket_scan(st, g, b);
}

}  // namespace msm
