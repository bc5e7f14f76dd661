// bucket_affine.cuh -- bucket accumulation by pairwise AFFINE additions with shared inversions.
//
// The XYZZ mixed addition of k_accumulate costs 10 field products per sorted digit (8M + 2S, 1288
// multiplier-pipe instructions for BN254) and that kernel already keeps the pipe 90 % busy: the only
// lever left is fewer products per addition.  An affine addition
//     lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1
// costs 2M + 1S plus the inversion, and Montgomery's trick turns B inversions into one inversion plus
// 3(B-1) products: 6 products per addition once B is in the hundreds.
//
// A bucket with n entries needs n - 1 additions in any order, so the sorted entry list is reduced in
// ROUNDS: round r pairs up neighbours inside every bucket (entries 2j, 2j+1 of the bucket -> output j,
// an odd last entry is passed through), halving every bucket.  Outputs are dense: the output array of a
// round is laid out bucket by bucket (offsets = exclusive scan of ceil(n/2), one small scan per round), so
// a bucket of any size -- the short top window of a folded table, equal scalars -- is simply many
// consecutive work items and needs no special handling.  After a few rounds the batches get too short to
// amortise the inversion; the remaining points (a handful per bucket) go through the existing XYZZ slice
// kernel, which also covers carry-in from earlier sub-batches.
//
// Work mapping (v2; the first version gave every thread a contiguous range and measured 165 - 200 ps per
// addition against 150 for XYZZ: 10 scattered 32-byte sectors per addition ran into the DRAM's random-access
// rate, not its bandwidth): a WARP owns a contiguous span of outputs and lane l takes outputs l, l + 32, ...
// of it, so that every load and store of the round -- the two inputs of an output are neighbours -- is a
// contiguous run across the warp.  Points between rounds are stored as two planes (all x, then all y): the
// forward pass needs x only.  Per batch of at most M_max outputs per thread:
//   forward   denominator d_i (x2 - x1, or 2 y1 for equal points), running product prefix_i = d_0 ... d_(i-1)
//             stored to scratch ([item][thread]: coalesced);
//   invert    ONE inversion per thread block: the threads' batch products are multiplied up a binary tree in
//             shared memory, one thread inverts the root (the other blocks of the SM keep the multiplier pipe
//             busy meanwhile), and the inverses are walked back down (2 products per level);
//   backward  for every item in reverse: 1/d_i = inv * prefix_i, inv *= d_i, then the affine formulas.
// Exceptional inputs are classified in the forward pass (kind stored next to the input index):
// identity operands (0,0), P + P (tangent slope, denominator 2y), P + (-P) (result identity).
//
// Replaces, for large calls, the per-thread serial bucket scan of POINT_multiexp_chunk
// (ag-build/cl/multiexp.cl:62-134) together with most of k_accumulate's XYZZ additions.
#pragma once
#include "ec.cuh"

namespace msm {

// largest g in [0, NB) with off[g] <= pos  (pos < off[NB])
MSM_D uint32_t ba_find_bucket(const uint32_t* __restrict__ off, uint32_t NB, uint32_t pos) {
  uint32_t lo = 0, hi = NB;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= pos) lo = mid;
    else hi = mid;
  }
  return lo;
}

// Pull the line(s) holding [p, p + bytes) into the L2 without occupying a register: the rounds stream through
// tens of GB with one dependent load chain per item, and ncu showed the warps waiting on those loads
// (long_scoreboard 3.8 warps per issue) rather than on the multiplier.
MSM_D void ba_prefetch(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

constexpr int BA_BLOCK = 128;
constexpr uint32_t BA_KIND_COPY = 0, BA_KIND_ADD = 1, BA_KIND_DBL = 2, BA_KIND_SPECIAL = 3;

// Input / output of a round.  Round 0 gathers: item a is the sorted entry entries[a] = table index | sign << 31.
// Later rounds read the planes of the previous round: x of item a at x + a * N words, y at y + a * N.
template <class F> struct BaPoints {
  const PackedAffine<F>* table;  // gather source (round 0), else nullptr
  const uint32_t* entries;
  uint32_t* x;                   // planes (input of rounds >= 1, output of every round)
  uint32_t* y;
};

template <class F> MSM_D void ba_store_elem(uint32_t* dst, const typename F::Elem& e) {
  static_assert(F::N % 4 == 0, "limb count must be a multiple of 4");
  uint4* q = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int j = 0; j < F::N / 4; j++) q[j] = make_uint4(e.v[4 * j], e.v[4 * j + 1], e.v[4 * j + 2], e.v[4 * j + 3]);
}
template <class F> MSM_D typename F::Elem ba_load_elem(const uint32_t* src) {
  typename F::Elem e;
  const uint4* q = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int j = 0; j < F::N / 4; j++) {
    const uint4 t = q[j];
    e.v[4 * j] = t.x; e.v[4 * j + 1] = t.y; e.v[4 * j + 2] = t.z; e.v[4 * j + 3] = t.w;
  }
  return e;
}
template <class F> MSM_D typename F::Elem ba_load_elem_ro(const uint32_t* src) {  // read-only path
  typename F::Elem e;
  const uint4* q = reinterpret_cast<const uint4*>(src);
#pragma unroll
  for (int j = 0; j < F::N / 4; j++) {
    const uint4 t = __ldg(q + j);
    e.v[4 * j] = t.x; e.v[4 * j + 1] = t.y; e.v[4 * j + 2] = t.z; e.v[4 * j + 3] = t.w;
  }
  return e;
}
// x / y of input item a (y with the entry's sign applied when gathering)
template <class F, bool GATHER> MSM_D typename F::Elem ba_in_x(const BaPoints<F>& in, uint32_t a) {
  if (GATHER) return ba_load_elem_ro<F>(in.table[__ldg(in.entries + a) & 0x7fffffffu].x);
  return ba_load_elem_ro<F>(in.x + (size_t)a * F::N);
}
template <class F, bool GATHER> MSM_D typename F::Elem ba_in_y(const BaPoints<F>& in, uint32_t a) {
  if (GATHER) {
    const uint32_t ent = __ldg(in.entries + a);
    typename F::Elem y = ba_load_elem_ro<F>(in.table[ent & 0x7fffffffu].y);
    if (ent >> 31) y = F::template neg<2, 1>(y);  // (0,0) stays (0,0): neg maps 0 to 0
    return y;
  }
  return ba_load_elem_ro<F>(in.y + (size_t)a * F::N);
}
template <class F> MSM_D bool ba_is_zero(const typename F::Elem& e) { return F::template is_multiple_of_p<0, 2>(e); }
template <class F> MSM_D void ba_out(const BaPoints<F>& out, uint32_t k, const typename F::Elem& x, const typename F::Elem& y) {
  // Planes hold the field's register form (lazy [0, 2p) values are not folded): the identity is written as
  // all-zero words explicitly, and a finite point never has x = y = 0 mod p, so "all words zero" still means
  // the identity and nothing else for every reader (the next round, k_accumulate's plane mode).
  static_assert(F::PACKED_WORDS == F::N, "planes store the register limbs");
  ba_store_elem<F>(out.x + (size_t)k * F::N, x);
  ba_store_elem<F>(out.y + (size_t)k * F::N, y);
}

// Position of a thread in the bucket structure while it walks its outputs in increasing order.
struct BaWalk {
  uint32_t g, ostart, oend, istart, n_in;
  bool ready;
};

// Geometry of a round: O outputs over T threads; a warp owns 32 * per consecutive outputs, lane l of warp w
// takes k = w * 32 * per + i * 32 + l for i < per.
struct BaGeom {
  uint32_t T, O, per;
};
MSM_D BaGeom ba_geom(uint32_t T, uint32_t O) {
  BaGeom gm;
  gm.T = T;
  gm.O = O;
  gm.per = (uint32_t)(((uint64_t)O + T - 1) / T);
  return gm;
}
MSM_D uint64_t ba_item(const BaGeom& gm, uint32_t t, uint32_t i) { return (uint64_t)(t >> 5) * 32 * gm.per + (uint64_t)i * 32 + (t & 31); }

// Forward pass of one batch (items i0 .. i0 + m of thread t): returns the product of the batch's denominators.
template <class F, bool GATHER>
MSM_D typename F::Elem ba_forward(uint32_t t, const BaGeom& gm, uint32_t i0, uint32_t m, const BaPoints<F>& in,
                                  const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out, uint32_t NB,
                                  BaWalk& wk, uint32_t* __restrict__ scratch_prefix, uint32_t* __restrict__ scratch_idx) {
  using E = typename F::Elem;
  constexpr int N = F::N;
  E prefix = F::one();
  const uint32_t n_items = __ldg(off_in + NB);  // inputs of the round
  for (uint32_t ii = 0; ii < m; ii++) {
    if (i0 + ii >= gm.per) break;
    const uint64_t k64 = ba_item(gm, t, i0 + ii);
    if (k64 >= gm.O) break;  // items grow with ii: nothing further either
    const uint32_t k = (uint32_t)k64;
    if (!wk.ready) {
      wk.g = ba_find_bucket(off_out, NB, k);
      wk.ostart = __ldg(off_out + wk.g);
      wk.oend = __ldg(off_out + wk.g + 1);
      wk.istart = __ldg(off_in + wk.g);
      wk.n_in = __ldg(off_in + wk.g + 1) - wk.istart;
      wk.ready = true;
    }
    while (k >= wk.oend) {  // next non-empty bucket
      wk.g++;
      wk.ostart = wk.oend;
      wk.oend = __ldg(off_out + wk.g + 1);
      wk.istart = __ldg(off_in + wk.g);
      wk.n_in = __ldg(off_in + wk.g + 1) - wk.istart;
    }
    const uint32_t j = k - wk.ostart, a = wk.istart + 2 * j;
    // the next item of this lane is 32 outputs = about 64 inputs further on: pull its line into the L2 now
    // (plane rounds: -7 %; requesting the gather round's random table rows the same way doubled its time)
    if (!GATHER && a + 65 < n_items) ba_prefetch(in.x + (size_t)(a + 64) * N);
    uint32_t kind = BA_KIND_COPY;
    if (2 * j + 1 < wk.n_in) {
      const E xa = ba_in_x<F, GATHER>(in, a), xb = ba_in_x<F, GATHER>(in, a + 1);
      E den = F::template sub<2, 1>(xb, xa);
      kind = BA_KIND_ADD;
      if (ba_is_zero<F>(den) || F::is_zero_limbs(xa) || F::is_zero_limbs(xb)) {
        // rare: equal x (P + P or P - P) or a coordinate that may belong to the identity (0,0)
        const E ya = ba_in_y<F, GATHER>(in, a), yb = ba_in_y<F, GATHER>(in, a + 1);
        const bool ida = F::is_zero_limbs(xa) && F::is_zero_limbs(ya), idb = F::is_zero_limbs(xb) && F::is_zero_limbs(yb);
        if (ida || idb) {
          kind = BA_KIND_SPECIAL;
        } else if (ba_is_zero<F>(den)) {
          if (ba_is_zero<F>(F::template sub<2, 1>(yb, ya)) && !ba_is_zero<F>(ya)) {
            kind = BA_KIND_DBL;
            den = F::add(ya, ya);
          } else {
            kind = BA_KIND_SPECIAL;  // P + (-P), or 2P with y = 0: the identity
          }
        }
      }
      if (kind != BA_KIND_SPECIAL) {
        ba_store_elem<F>(scratch_prefix + ((size_t)ii * gm.T + t) * N, prefix);
        prefix = F::mul(prefix, den);
      }
    }
    scratch_idx[(size_t)ii * gm.T + t] = a | (kind << 30);
  }
  return prefix;
}

// Backward pass of the same batch; inv = 1 / (product returned by ba_forward).
template <class F, bool GATHER>
MSM_D void ba_backward(uint32_t t, const BaGeom& gm, uint32_t i0, uint32_t m, const BaPoints<F>& in, const BaPoints<F>& out,
                       typename F::Elem inv, const uint32_t* __restrict__ scratch_prefix,
                       const uint32_t* __restrict__ scratch_idx) {
  using E = typename F::Elem;
  constexpr int N = F::N;
  for (uint32_t ii = m; ii-- > 0;) {
    if (i0 + ii >= gm.per) continue;
    const uint64_t k64 = ba_item(gm, t, i0 + ii);
    if (k64 >= gm.O) continue;
    const uint32_t k = (uint32_t)k64;
    const uint32_t word = scratch_idx[(size_t)ii * gm.T + t];
    const uint32_t a = word & 0x3fffffffu, kind = word >> 30;
    // the item this lane handles next (ii - 1) sits about 64 inputs back
    if (ii && !GATHER) {
      ba_prefetch(scratch_prefix + ((size_t)(ii - 1) * gm.T + t) * N);
      if (a >= 64) {
        ba_prefetch(in.x + (size_t)(a - 64) * N);
        ba_prefetch(in.y + (size_t)(a - 64) * N);
      }
    }
    const E xa = ba_in_x<F, GATHER>(in, a), ya = ba_in_y<F, GATHER>(in, a);
    if (kind == BA_KIND_COPY) {
      ba_out<F>(out, k, xa, ya);
      continue;
    }
    const E xb = ba_in_x<F, GATHER>(in, a + 1), yb = ba_in_y<F, GATHER>(in, a + 1);
    if (kind == BA_KIND_SPECIAL) {
      const bool ida = F::is_zero_limbs(xa) && F::is_zero_limbs(ya), idb = F::is_zero_limbs(xb) && F::is_zero_limbs(yb);
      if (ida) ba_out<F>(out, k, xb, yb);
      else if (idb) ba_out<F>(out, k, xa, ya);
      else ba_out<F>(out, k, F::zero(), F::zero());
      continue;
    }
    E den, num;
    if (kind == BA_KIND_DBL) {
      den = F::add(ya, ya);
      const E xx = F::sqr(xa);
      num = F::add(F::add(xx, xx), xx);
    } else {
      den = F::template sub<2, 1>(xb, xa);
      num = F::template sub<2, 1>(yb, ya);
    }
    const E pre = ba_load_elem<F>(scratch_prefix + ((size_t)ii * gm.T + t) * N);
    const E dinv = F::mul(inv, pre);
    inv = F::mul(inv, den);
    const E lam = F::mul(num, dinv);
    const E x3 = F::template sub<2, 1>(F::template sub<2, 1>(F::sqr(lam), xa), xb);
    const E y3 = F::template sub<2, 1>(F::mul(lam, F::template sub<2, 1>(xa, x3)), ya);
    ba_out<F>(out, k, x3, y3);
  }
}

#if defined(__CUDACC__)
// One halving round.  off_in / off_out: [NB + 1] exclusive offsets of the buckets in the input and output
// arrays (off_out = scan of ceil(n_in / 2)).  scratch_prefix: [M_max][T][N words], scratch_idx: [M_max][T],
// T = gridDim.x * blockDim.x.
template <class F, bool GATHER, int BPS>
__global__ void __launch_bounds__(BA_BLOCK, BPS)
k_affine_round(BaPoints<F> in, const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out, uint32_t NB,
               uint32_t M_max, BaPoints<F> out, uint32_t* __restrict__ scratch_prefix, uint32_t* __restrict__ scratch_idx) {
  using E = typename F::Elem;
  constexpr int N = F::N;
  __shared__ uint4 tree_raw[2 * BA_BLOCK * N / 4];
  uint32_t* tree = reinterpret_cast<uint32_t*>(tree_raw);  // node n at tree + n * N; leaves BA_BLOCK .. 2 BA_BLOCK - 1
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x, tid = threadIdx.x;
  const BaGeom gm = ba_geom(gridDim.x * blockDim.x, __ldg(off_out + NB));
  BaWalk wk;
  wk.ready = false;
  // the thread that inverts sits in a different warp slot from block to block: the blocks of an SM then do not
  // queue their inversions on one scheduler
  const uint32_t inverter = 32u * (blockIdx.x & (BA_BLOCK / 32 - 1));
  for (uint32_t i0 = 0; i0 < gm.per; i0 += M_max) {  // same trip count for every thread of the grid
    const E prod = ba_forward<F, GATHER>(t, gm, i0, M_max, in, off_in, off_out, NB, wk, scratch_prefix, scratch_idx);
    ba_store_elem<F>(tree + (BA_BLOCK + tid) * N, prod);
    __syncthreads();
    for (uint32_t s = BA_BLOCK / 2; s >= 1; s >>= 1) {
      if (tid < s) {
        const uint32_t n = s + tid;
        ba_store_elem<F>(tree + n * N, F::mul(ba_load_elem<F>(tree + 2 * n * N), ba_load_elem<F>(tree + (2 * n + 1) * N)));
      }
      __syncthreads();
    }
    if (tid == inverter) ba_store_elem<F>(tree + N, F::inv(ba_load_elem<F>(tree + N)));
    __syncthreads();
    for (uint32_t s = 1; s < BA_BLOCK; s <<= 1) {
      if (tid < s) {
        const uint32_t n = s + tid;
        const E inv_n = ba_load_elem<F>(tree + n * N);
        const E l = ba_load_elem<F>(tree + 2 * n * N), r = ba_load_elem<F>(tree + (2 * n + 1) * N);
        ba_store_elem<F>(tree + 2 * n * N, F::mul(inv_n, r));
        ba_store_elem<F>(tree + (2 * n + 1) * N, F::mul(inv_n, l));
      }
      __syncthreads();
    }
    const E inv = ba_load_elem<F>(tree + (BA_BLOCK + tid) * N);
    __syncthreads();  // the tree is rewritten by the next batch
    ba_backward<F, GATHER>(t, gm, i0, M_max, in, out, inv, scratch_prefix, scratch_idx);
  }
}
#endif

}  // namespace msm
