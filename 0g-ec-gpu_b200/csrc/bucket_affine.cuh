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
// consecutive work items and needs no special handling.  After a few rounds the per-thread batches get
// too short to amortise the inversion; the remaining points (a handful per bucket) go through the
// existing XYZZ slice kernel, which also covers carry-in from earlier sub-batches.
//
// One thread owns a contiguous range of output items and runs, per batch of at most M_max items:
//   forward   for every item: denominator d_i (x2 - x1, or 2 y1 for equal points), running product
//             prefix_i = d_0 ... d_(i-1) stored to a scratch column (coalesced: [item][thread]);
//   invert    one field inversion of the batch's total product;
//   backward  for every item in reverse: 1/d_i = inv * prefix_i, inv *= d_i, then the affine formulas.
// Exceptional inputs are classified in the forward pass (kind stored next to the input index):
// identity operands (0,0), P + P (tangent slope, denominator 2y), P + (-P) (result identity).
//
// Replaces, for large calls, the per-thread serial bucket scan of POINT_multiexp_chunk
// (ag-build/cl/multiexp.cl:62-134) together with k_accumulate's XYZZ additions.
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

constexpr int BA_BLOCK = 128;
constexpr uint32_t BA_KIND_COPY = 0, BA_KIND_ADD = 1, BA_KIND_DBL = 2, BA_KIND_SPECIAL = 3;

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

// Input item `a` of a round: GATHER = true reads the sorted entry (table index | sign) and fetches the base
// point with the sign applied; otherwise the point is element `a` of the previous round's output.
template <class F, bool GATHER>
MSM_D const PackedAffine<F>* ba_src(const PackedAffine<F>* __restrict__ src, const uint32_t* __restrict__ entries,
                                    uint32_t a, bool& negate) {
  if (GATHER) {
    const uint32_t ent = __ldg(entries + a);
    negate = (ent >> 31) != 0;
    return src + (ent & 0x7fffffffu);
  }
  negate = false;
  return src + a;
}
template <class F> MSM_D typename F::Elem ba_load_x(const PackedAffine<F>* p) {
  uint32_t w[F::PACKED_WORDS];
  const uint4* q = reinterpret_cast<const uint4*>(p->x);
#pragma unroll
  for (int j = 0; j < F::PACKED_WORDS / 4; j++) {
    const uint4 t = __ldg(q + j);
    w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
  }
  return F::unpack(w);
}
template <class F> MSM_D typename F::Elem ba_load_y(const PackedAffine<F>* p, bool negate) {
  uint32_t w[F::PACKED_WORDS];
  const uint4* q = reinterpret_cast<const uint4*>(p->y);
#pragma unroll
  for (int j = 0; j < F::PACKED_WORDS / 4; j++) {
    const uint4 t = __ldg(q + j);
    w[4 * j] = t.x; w[4 * j + 1] = t.y; w[4 * j + 2] = t.z; w[4 * j + 3] = t.w;
  }
  typename F::Elem y = F::unpack(w);
  if (negate) y = F::template neg<2, 1>(y);  // (0,0) stays (0,0): neg maps 0 to 0
  return y;
}
template <class F> MSM_D bool ba_is_zero(const typename F::Elem& e) { return F::template is_multiple_of_p<0, 2>(e); }

template <class F> MSM_D void ba_store_point(PackedAffine<F>* dst, const typename F::Elem& x, const typename F::Elem& y) {
  PackedAffine<F> o;
  F::to_packed(x, o.x);
  F::to_packed(y, o.y);
  uint4* q = reinterpret_cast<uint4*>(dst);
  const uint32_t* w = o.x;  // x then y, contiguous
#pragma unroll
  for (int j = 0; j < 2 * F::PACKED_WORDS / 4; j++) q[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}

// One halving round.  off_in / off_out: [NB + 1] exclusive offsets of the buckets in the input and output
// arrays (off_out = scan of ceil(n_in / 2)).  scratch_prefix: [M_max][T][N words], scratch_idx: [M_max][T],
// T = gridDim.x * blockDim.x.  Every thread handles ceil(O / T) consecutive outputs, O = off_out[NB].
// (the body of one thread; also compiled for the host, where tests/test_host_arith.py runs it thread by
// thread against the oracle -- threads of a round are independent)
template <class F, bool GATHER>
MSM_D void ba_round_thread(uint32_t t, uint32_t T, const PackedAffine<F>* __restrict__ src,
                           const uint32_t* __restrict__ entries, const uint32_t* __restrict__ off_in,
                           const uint32_t* __restrict__ off_out, uint32_t NB, uint32_t M_max,
                           PackedAffine<F>* __restrict__ dst, uint32_t* __restrict__ scratch_prefix,
                           uint32_t* __restrict__ scratch_idx) {
  using E = typename F::Elem;
  constexpr int N = F::N;
  const uint32_t O = __ldg(off_out + NB);
  const uint32_t per = (O + T - 1) / T;
  const uint64_t k_first = (uint64_t)t * per;
  if (k_first >= O) return;
  const uint32_t k_last = (uint32_t)(k_first + per < (uint64_t)O ? k_first + per : (uint64_t)O);
  uint32_t* my_prefix = scratch_prefix + (size_t)t * N;
  uint32_t* my_idx = scratch_idx + t;
  const size_t pstride = (size_t)T * N;

  // bucket of the first output: largest g with off_out[g] <= k_first
  uint32_t g = ba_find_bucket(off_out, NB, (uint32_t)k_first);
  uint32_t ostart = __ldg(off_out + g), oend = __ldg(off_out + g + 1);
  uint32_t istart = __ldg(off_in + g), n_in = __ldg(off_in + g + 1) - istart;

  for (uint32_t k0 = (uint32_t)k_first; k0 < k_last; k0 += M_max) {
    const uint32_t m = M_max < k_last - k0 ? M_max : k_last - k0;
    // ---------------- forward: denominators and their running product
    E prefix = F::one();
    for (uint32_t i = 0; i < m; i++) {
      const uint32_t k = k0 + i;
      while (k >= oend) {  // next non-empty bucket
        g++;
        ostart = oend;
        oend = __ldg(off_out + g + 1);
        istart = __ldg(off_in + g);
        n_in = __ldg(off_in + g + 1) - istart;
      }
      const uint32_t j = k - ostart, a = istart + 2 * j;
      uint32_t kind = BA_KIND_COPY;
      if (2 * j + 1 < n_in) {
        bool na, nb;
        const PackedAffine<F>* pa = ba_src<F, GATHER>(src, entries, a, na);
        const PackedAffine<F>* pb = ba_src<F, GATHER>(src, entries, a + 1, nb);
        const E xa = ba_load_x<F>(pa), xb = ba_load_x<F>(pb);
        E den = F::template sub<2, 1>(xb, xa);
        kind = BA_KIND_ADD;
        if (ba_is_zero<F>(den) || F::is_zero_limbs(xa) || F::is_zero_limbs(xb)) {
          // rare: equal x (P + P or P - P) or a coordinate that may belong to the identity (0,0)
          const E ya = ba_load_y<F>(pa, na), yb = ba_load_y<F>(pb, nb);
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
          ba_store_elem<F>(my_prefix + (size_t)i * pstride, prefix);
          prefix = F::mul(prefix, den);
        }
      }
      my_idx[(size_t)i * T] = a | (kind << 30);
    }
    // ---------------- one inversion for the whole batch
    E inv = F::inv(prefix);
    // ---------------- backward: the additions, last item first
    for (uint32_t i = m; i-- > 0;) {
      const uint32_t word = my_idx[(size_t)i * T];
      const uint32_t a = word & 0x3fffffffu, kind = word >> 30;
      bool na, nb;
      const PackedAffine<F>* pa = ba_src<F, GATHER>(src, entries, a, na);
      const E xa = ba_load_x<F>(pa), ya = ba_load_y<F>(pa, na);
      PackedAffine<F>* out = dst + (k0 + i);
      if (kind == BA_KIND_COPY) {
        ba_store_point<F>(out, xa, ya);
        continue;
      }
      const PackedAffine<F>* pb = ba_src<F, GATHER>(src, entries, a + 1, nb);
      const E xb = ba_load_x<F>(pb), yb = ba_load_y<F>(pb, nb);
      if (kind == BA_KIND_SPECIAL) {
        const bool ida = F::is_zero_limbs(xa) && F::is_zero_limbs(ya), idb = F::is_zero_limbs(xb) && F::is_zero_limbs(yb);
        if (ida) ba_store_point<F>(out, xb, yb);
        else if (idb) ba_store_point<F>(out, xa, ya);
        else ba_store_point<F>(out, F::zero(), F::zero());
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
      const E pre = ba_load_elem<F>(my_prefix + (size_t)i * pstride);
      const E dinv = F::mul(inv, pre);
      inv = F::mul(inv, den);
      const E lam = F::mul(num, dinv);
      const E x3 = F::template sub<2, 1>(F::template sub<2, 1>(F::sqr(lam), xa), xb);
      const E y3 = F::template sub<2, 1>(F::mul(lam, F::template sub<2, 1>(xa, x3)), ya);
      ba_store_point<F>(out, x3, y3);
    }
  }
}

#if defined(__CUDACC__)
template <class F, bool GATHER>
__global__ void __launch_bounds__(BA_BLOCK, (F::N <= 8 ? 3 : 2))
k_affine_round(const PackedAffine<F>* __restrict__ src, const uint32_t* __restrict__ entries,
               const uint32_t* __restrict__ off_in, const uint32_t* __restrict__ off_out, uint32_t NB, uint32_t M_max,
               PackedAffine<F>* __restrict__ dst, uint32_t* __restrict__ scratch_prefix, uint32_t* __restrict__ scratch_idx) {
  ba_round_thread<F, GATHER>(blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x, src, entries, off_in, off_out, NB,
                             M_max, dst, scratch_prefix, scratch_idx);
}

// counts of the next round from the offsets of this one: n_out[g] = ceil(n_in[g] / 2)
static __global__ void k_halve_counts(const uint32_t* __restrict__ off_in, uint32_t NB, uint32_t* __restrict__ counts) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g < NB) counts[g] = (__ldg(off_in + g + 1) - __ldg(off_in + g) + 1) >> 1;
}
#endif

}  // namespace msm
