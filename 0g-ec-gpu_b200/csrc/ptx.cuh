// ptx.cuh -- 32-bit carry-chain primitives.
//
// Device: explicit PTX add.cc / addc.cc / sub.cc / mad.lo.cc / madc.hi.cc chains (replaces the
// wrappers of ag-build/cl/common.cl:127-248).  ptxas fuses each mad.lo.cc + madc.hi.cc pair on
// the same operands into one IMAD.WIDE.U32[.X] on sm_100a (checked with cuobjdump -sass).
//
// Host: the same functions emulate the PTX carry flag with a thread-local variable, so the field
// and curve code of fp.cuh / ec.cuh can be compiled by g++ and unit-tested on a box without a GPU
// (tests/test_host_arith.py).  The host build is a test vehicle, not a fallback: no product entry
// point calls it.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MSM_HD __host__ __device__ __forceinline__
#define MSM_D __device__ __forceinline__
// cold-path group operations: one out-of-line copy per curve keeps the hot kernel's register
// budget and the compile time down
#define MSM_COLD __host__ __device__ __noinline__
#else
#define MSM_HD inline
#define MSM_D inline
#define MSM_COLD inline
#ifndef __restrict__
#define __restrict__ __restrict
#endif
// host build (test vehicle): the few CUDA built-ins the per-thread device functions use
struct alignas(16) uint4 {
  uint32_t x, y, z, w;
};
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
template <class T> inline T __ldg(const T* p) { return *p; }
#endif

namespace msm {

#if defined(__CUDA_ARCH__)
// ---------------------------------------------------------------- device: PTX
MSM_D uint32_t add_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
MSM_D uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
MSM_D uint32_t addc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
MSM_D uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
MSM_D uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
MSM_D uint32_t subc(uint32_t a, uint32_t b) {
  uint32_t r;
  asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// (lo,hi) = a*b                                    -> IMAD.WIDE.U32
MSM_D void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) += a*b, carry-out in CC                  -> IMAD.WIDE.U32 with carry-out
MSM_D void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi)
               : "r"(a), "r"(b));
}
// (lo,hi) += a*b + CC, carry-out in CC             -> IMAD.WIDE.U32.X
MSM_D void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
               : "+r"(lo), "+r"(hi)
               : "r"(a), "r"(b));
}
// (dlo,dhi) = a*b + (clo,chi) + CC, carry-out in CC
MSM_D void madc_wide_cc3(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t clo,
                         uint32_t chi) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4;\n\tmadc.hi.cc.u32 %1, %2, %3, %5;"
               : "=r"(dlo), "=r"(dhi)
               : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// (dlo,dhi) = a*b + CC   (cannot overflow 64 bits)
MSM_D void madc_wide_0(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, 0;\n\tmadc.hi.u32 %1, %2, %3, 0;"
               : "=r"(dlo), "=r"(dhi)
               : "r"(a), "r"(b));
}
MSM_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
// carry-save forms: the wide multiply-add takes no carry in; its carry out is counted in `cnt`
// (lo,hi) += a*b ; cnt += carry                     -> IMAD.WIDE.U32 (carry-out) + IADD3.X
MSM_D void mad_wide_cs(uint32_t& lo, uint32_t& hi, uint32_t& cnt, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
               : "+r"(lo), "+r"(hi), "+r"(cnt)
               : "r"(a), "r"(b));
}
// (dlo,dhi) = a*b + (clo,chi) ; cnt += carry
MSM_D void mad_wide_cs3(uint32_t& dlo, uint32_t& dhi, uint32_t& cnt, uint32_t a, uint32_t b, uint32_t clo,
                        uint32_t chi) {
  asm volatile("mad.lo.cc.u32 %0, %3, %4, %5;\n\tmadc.hi.cc.u32 %1, %3, %4, %6;\n\taddc.u32 %2, %2, 0;"
               : "=r"(dlo), "=r"(dhi), "+r"(cnt)
               : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// x += y ; cnt += carry
MSM_D void add_cs(uint32_t& x, uint32_t& cnt, uint32_t y) {
  asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+r"(x), "+r"(cnt) : "r"(y));
}

#else
// ---------------------------------------------------------------- host: emulated carry flag
namespace detail {
inline uint32_t& cf() {
  static thread_local uint32_t flag = 0;
  return flag;
}
}  // namespace detail
inline uint32_t add_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a + b;
  detail::cf() = (uint32_t)(s >> 32);
  return (uint32_t)s;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a + b + detail::cf();
  detail::cf() = (uint32_t)(s >> 32);
  return (uint32_t)s;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + detail::cf(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
  uint64_t d = (uint64_t)a - b;
  detail::cf() = (uint32_t)(d >> 63);
  return (uint32_t)d;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
  uint64_t d = (uint64_t)a - b - detail::cf();
  detail::cf() = (uint32_t)(d >> 63);
  return (uint32_t)d;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - detail::cf(); }
inline void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
}
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  unsigned __int128 s = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)hi << 32) | lo);
  lo = (uint32_t)s;
  hi = (uint32_t)(s >> 32);
  detail::cf() = (uint32_t)(s >> 64);
}
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  unsigned __int128 s =
      (unsigned __int128)((uint64_t)a * b) + (((uint64_t)hi << 32) | lo) + detail::cf();
  lo = (uint32_t)s;
  hi = (uint32_t)(s >> 32);
  detail::cf() = (uint32_t)(s >> 64);
}
inline void madc_wide_cc3(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b, uint32_t clo,
                          uint32_t chi) {
  unsigned __int128 s =
      (unsigned __int128)((uint64_t)a * b) + (((uint64_t)chi << 32) | clo) + detail::cf();
  dlo = (uint32_t)s;
  dhi = (uint32_t)(s >> 32);
  detail::cf() = (uint32_t)(s >> 64);
}
inline void madc_wide_0(uint32_t& dlo, uint32_t& dhi, uint32_t a, uint32_t b) {
  uint64_t s = (uint64_t)a * b + detail::cf();
  dlo = (uint32_t)s;
  dhi = (uint32_t)(s >> 32);
}
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline void mad_wide_cs(uint32_t& lo, uint32_t& hi, uint32_t& cnt, uint32_t a, uint32_t b) {
  unsigned __int128 s = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)hi << 32) | lo);
  lo = (uint32_t)s;
  hi = (uint32_t)(s >> 32);
  cnt += (uint32_t)(s >> 64);
}
inline void mad_wide_cs3(uint32_t& dlo, uint32_t& dhi, uint32_t& cnt, uint32_t a, uint32_t b, uint32_t clo,
                         uint32_t chi) {
  unsigned __int128 s = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)chi << 32) | clo);
  dlo = (uint32_t)s;
  dhi = (uint32_t)(s >> 32);
  cnt += (uint32_t)(s >> 64);
}
inline void add_cs(uint32_t& x, uint32_t& cnt, uint32_t y) {
  uint64_t s = (uint64_t)x + y;
  x = (uint32_t)s;
  cnt += (uint32_t)(s >> 32);
}
#endif

}  // namespace msm
