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
#include "sort.h"

namespace msm {

// Geometry, Plan and the sort entry points live in sort.h (the sort is field-independent: one unit, sort.cu)

MSM_D void load_scalar(const uint32_t* scalars, uint32_t i, uint32_t k[8]) {
  const uint4* p = reinterpret_cast<const uint4*>(scalars) + 2 * (size_t)i;
  uint4 a = __ldg(p), b = __ldg(p + 1);
  k[0] = a.x; k[1] = a.y; k[2] = a.z; k[3] = a.w;
  k[4] = b.x; k[5] = b.y; k[6] = b.z; k[7] = b.w;
}

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
             uint32_t* __restrict__ cut_count, uint32_t* __restrict__ cut_list, uint32_t cut_cap,
             const uint32_t* __restrict__ plane_x, const uint32_t* __restrict__ plane_y) {
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
  // plane_x != nullptr: the points to add are items s .. e of the x / y planes the affine halving rounds left
  // (bucket_affine.cuh), already signed; `entries` and `bases` are not used
  const bool direct = plane_x != nullptr;
  auto fetch = [&](uint32_t idx, uint32_t* w) {
    if (direct) {
      const uint4* qx = reinterpret_cast<const uint4*>(plane_x + (size_t)idx * F::PACKED_WORDS);
      const uint4* qy = reinterpret_cast<const uint4*>(plane_y + (size_t)idx * F::PACKED_WORDS);
#pragma unroll
      for (int j = 0; j < F::PACKED_WORDS / 4; j++) {
        const uint4 a = __ldg(qx + j), b = __ldg(qy + j);
        w[4 * j] = a.x; w[4 * j + 1] = a.y; w[4 * j + 2] = a.z; w[4 * j + 3] = a.w;
        w[F::PACKED_WORDS + 4 * j] = b.x; w[F::PACKED_WORDS + 4 * j + 1] = b.y;
        w[F::PACKED_WORDS + 4 * j + 2] = b.z; w[F::PACKED_WORDS + 4 * j + 3] = b.w;
      }
    } else {
      load_base_words<F>(bases + idx, w);
    }
  };
  uint32_t nxt[WORDS];
  uint32_t ent = direct ? s : __ldg(entries + s);
  fetch(ent & 0x7fffffffu, nxt);
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
      fetch(ent & 0x7fffffffu, nxt);
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
constexpr uint32_t HEAVY_SPAN = 32;
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
      run = xyzz_add_inl<F>(run, load_vec(&bucket_acc[f + k]));
      res = xyzz_add_inl<F>(res, run);
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
