// ntt.cuh -- radix-2 number-theoretic transform over the scalar field Fr.
//
// Replaces KERNEL FIELD_radix_fft (ag-build/cl/fft.cl:4-66) and its host driver
// SingleFftKernel::radix_fft (ec-gpu-proxy/src/fft.rs:50-136):
//     out[k] = sum_j omega^(j k) * in[j],  n = 2^log_n elements of Fr in Montgomery form (arkworks'
// in-memory layout), natural order in and out -- what serial_fft computes
// (ec-gpu-proxy/src/fft_cpu.rs:10-52).  The reference runs radix-2^deg passes (deg <= 8) that recompute
// twiddle powers per work item (FIELD_pow_lookup + FIELD_pow); here: one table tw[j] = omega^j built
// with one product per entry, a first pass that does the bit-reversal gather and up to 10 rounds in
// shared memory, then passes of up to 5 rounds on 32 x 32 tiles (32 consecutive elements per strided
// row, so every global access is a full 1 KB run).  Each element crosses HBM once per pass.
#pragma once
#include "kernels.cuh"

namespace msm {

template <class PR> MSM_D Fp<PR> ntt_ld(const uint32_t* p, size_t i) {
  const uint4* q = reinterpret_cast<const uint4*>(p) + 2 * i;
  const uint4 a = q[0], b = q[1];
  Fp<PR> r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
template <class PR> MSM_D void ntt_st(uint32_t* p, size_t i, const Fp<PR>& v) {
  uint4* q = reinterpret_cast<uint4*>(p) + 2 * i;
  q[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
  q[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}

// out[t] = (base^(2^pre_sq))^t for t < count (Montgomery form), binary method per thread
template <class PR>
__global__ void k_ntt_pow_table(Fp<PR> base, uint32_t pre_sq, uint32_t count, uint32_t* __restrict__ out) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= count) return;
  for (uint32_t i = 0; i < pre_sq; i++) base = fp_mul<PR>(base, base);
  Fp<PR> acc = fp_one<PR>();
  for (uint32_t e = t; e; e >>= 1) {
    if (e & 1) acc = fp_mul<PR>(acc, base);
    base = fp_mul<PR>(base, base);
  }
  ntt_st<PR>(out, t, acc);
}
// tw[j] = hi[j >> 10] * tw[j & 1023] for 1024 <= j < count
template <class PR>
__global__ void k_ntt_expand_table(const uint32_t* __restrict__ hi, uint32_t count, uint32_t* __restrict__ tw) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x + 1024;
  if (j >= count) return;
  ntt_st<PR>(tw, j, fp_mul<PR>(ntt_ld<PR>(hi, j >> 10), ntt_ld<PR>(tw, j & 1023)));
}

// (lo, hi) <- (lo + w hi, lo - w hi)
template <class PR> MSM_D void ntt_butterfly(Fp<PR>& lo, Fp<PR>& hi, const uint32_t* __restrict__ tw, size_t tw_index) {
  const Fp<PR> t = tw_index ? fp_mul<PR>(hi, ntt_ld<PR>(tw, tw_index)) : hi;
  hi = fp_sub<PR>(lo, t);
  lo = fp_add<PR>(lo, t);
}

constexpr int NTT_BLOCK = 512;
// Rounds 0 .. R-1 on chunks of 2^R consecutive elements of the bit-reversed input.
template <class PR>
__global__ void __launch_bounds__(NTT_BLOCK)
k_ntt_first(const uint32_t* __restrict__ x, uint32_t* __restrict__ y, uint32_t log_n, uint32_t R,
            const uint32_t* __restrict__ tw) {
  extern __shared__ uint4 ntt_smem[];
  Fp<PR>* u = reinterpret_cast<Fp<PR>*>(ntt_smem);
  const uint32_t chunk = 1u << R, base = blockIdx.x << R;
  // the 2^(R-1) twiddles these rounds use, omega^(k n / 2^R), staged once per block: round s reads
  // tws[j << (R-1-s)] from shared memory instead of one scattered 32-byte global load per butterfly
  Fp<PR>* tws = u + chunk;
  for (uint32_t k = threadIdx.x; k < chunk / 2; k += NTT_BLOCK) tws[k] = ntt_ld<PR>(tw, (size_t)k << (log_n - R));
  for (uint32_t i = threadIdx.x; i < chunk; i += NTT_BLOCK) u[i] = ntt_ld<PR>(x, __brev(base + i) >> (32 - log_n));
  __syncthreads();
  for (uint32_t s = 0; s < R; s++) {
    const uint32_t m = 1u << s;
    for (uint32_t b = threadIdx.x; b < chunk / 2; b += NTT_BLOCK) {
      const uint32_t j = b & (m - 1), i0 = ((b - j) << 1) + j, i1 = i0 + m;
      Fp<PR> lo = u[i0], hi = u[i1];
      const Fp<PR> t = j ? fp_mul<PR>(hi, tws[j << (R - 1 - s)]) : hi;
      hi = fp_sub<PR>(lo, t);
      lo = fp_add<PR>(lo, t);
      u[i0] = lo;
      u[i1] = hi;
    }
    __syncthreads();
  }
  for (uint32_t i = threadIdx.x; i < chunk; i += NTT_BLOCK) ntt_st<PR>(y, base + i, u[i]);
}

// Rounds s0 .. s0+r-1 (r <= 5, 2^s0 >= 32) in place: a tile is 32 consecutive low indices x 2^r rows
// that are 2^s0 apart; index = hi 2^(s0+r) + q 2^s0 + lo.
template <class PR>
__global__ void __launch_bounds__(NTT_BLOCK)
k_ntt_pass(uint32_t* __restrict__ y, uint32_t log_n, uint32_t s0, uint32_t r, const uint32_t* __restrict__ tw) {
  extern __shared__ uint4 ntt_smem[];
  Fp<PR>* u = reinterpret_cast<Fp<PR>*>(ntt_smem);
  const uint32_t tiles_lo = 1u << (s0 - 5);
  const uint32_t lo0 = (blockIdx.x & (tiles_lo - 1)) << 5;
  const size_t base = ((size_t)(blockIdx.x >> (s0 - 5)) << (s0 + r)) + lo0;
  const uint32_t elems = 32u << r;
  // Twiddles of the tile, staged once: round t needs omega^(j << (log_n - s0 - t - 1)) for j = (qlow << s0) + lo0 + l,
  // qlow < 2^t, l < 32 -- (2^r - 1) * 32 values per tile.  Reading them from the global table inside every butterfly
  // left the warps waiting on those loads (ncu: long_scoreboard 5 - 6.6 per issue, multiplier pipe 60 % busy); staged
  // together with the tile the latency is paid once per block.
  Fp<PR>* tws = u + elems;
  const uint32_t n_tw = ((1u << r) - 1) * 32;
  for (uint32_t idx = threadIdx.x; idx < n_tw; idx += NTT_BLOCK) {
    const uint32_t grp = (idx >> 5) + 1, l = idx & 31;
    const uint32_t t = 31 - __clz(grp), qlow = grp - (1u << t);
    const size_t j = ((size_t)qlow << s0) + lo0 + l;
    tws[idx] = ntt_ld<PR>(tw, j << (log_n - s0 - t - 1));
  }
  for (uint32_t e = threadIdx.x; e < elems; e += NTT_BLOCK) u[e] = ntt_ld<PR>(y, base + ((size_t)(e >> 5) << s0) + (e & 31));
  __syncthreads();
  for (uint32_t t = 0; t < r; t++) {
    for (uint32_t b = threadIdx.x; b < elems / 2; b += NTT_BLOCK) {
      const uint32_t l = b & 31, qb = b >> 5;
      const uint32_t qlow = qb & ((1u << t) - 1), q0 = ((qb - qlow) << 1) + qlow, q1 = q0 + (1u << t);
      Fp<PR> lo = u[q0 * 32 + l], hi = u[q1 * 32 + l];
      const Fp<PR> tm = (qlow | lo0 | l) ? fp_mul<PR>(hi, tws[(((1u << t) - 1 + qlow) << 5) + l]) : hi;  // j = 0: omega^0
      hi = fp_sub<PR>(lo, tm);
      lo = fp_add<PR>(lo, tm);
      u[q0 * 32 + l] = lo;
      u[q1 * 32 + l] = hi;
    }
    __syncthreads();
  }
  for (uint32_t e = threadIdx.x; e < elems; e += NTT_BLOCK) ntt_st<PR>(y, base + ((size_t)(e >> 5) << s0) + (e & 31), u[e]);
}

}  // namespace msm
