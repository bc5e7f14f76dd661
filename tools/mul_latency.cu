// mul_latency.cu -- latency (not throughput) of one dependent field product with a single warp per
// SM: what the serial sections pay (window combine / Horner, EC-FFT scalar multiplications at small
// n, the per-thread part of the bucket reduction).  Variants: the carry-chained CIOS product
// (fp_mul_nored), the carry-save product (fp_mul_cs: no multiply takes a carry in, so the 2N
// multiply-adds of a row are independent), and the FP64-pipe product (f52.cuh).
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../0g-ec-gpu_b200/csrc/ec.cuh"
#include "f52.cuh"
using namespace msm;

template <class P, int MODE>
__global__ void k_chain(uint32_t* out, int iters, unsigned long long* cycles) {
  Fp<P> x, y;
#pragma unroll
  for (int k = 0; k < P::N; k++) { x.v[k] = threadIdx.x + k + 1; y.v[k] = 7 * k + 3; }
  x.v[P::N - 1] &= 0x0fffffff; y.v[P::N - 1] &= 0x0fffffff;
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (MODE == 0) x = fp_mul_nored<P>(x, y);
    else { x = fp_mul_cs<P>(x, y); }
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < P::N; k++) s ^= x.v[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}
template <class Q>
__global__ void k_chain52(uint32_t* out, int iters, unsigned long long* cycles) {
  double x[Q::N], y[Q::N];
#pragma unroll
  for (int k = 0; k < Q::N; k++) { x[k] = (double)(threadIdx.x + k + 1); y[k] = (double)(7 * k + 3); }
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    uint64_t r[Q::N];
    f52_mul_core<Q>(r, x, y);
#pragma unroll
    for (int k = 0; k < Q::N; k++) x[k] = u52_to_double(r[k]);
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < Q::N; k++) s ^= (uint32_t)__double_as_longlong(x[k]);
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
  uint32_t* out; unsigned long long* cyc; unsigned long long h;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  const int IT = 4000;
  printf("{");
#define RUN(name, ...) __VA_ARGS__<<<1, 32>>>(out, IT, cyc); cudaDeviceSynchronize(); __VA_ARGS__<<<1, 32>>>(out, IT, cyc); cudaDeviceSynchronize(); \
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("\"%s\": %.1f, ", name, (double)h / IT);
  RUN("bn254_cios_cycles_per_dependent_product", k_chain<Bn254Fq, 0>)
  RUN("bn254_carry_save_cycles_per_dependent_product", k_chain<Bn254Fq, 1>)
  RUN("bn254_dfma_cycles_per_dependent_product", k_chain52<Bn254Fq52>)
  RUN("bls381_cios_cycles_per_dependent_product", k_chain<Bls381Fq, 0>)
  RUN("bls381_carry_save_cycles_per_dependent_product", k_chain<Bls381Fq, 1>)
  RUN("bls381_dfma_cycles_per_dependent_product", k_chain52<Bls381Fq52>)
  printf("\"warps\": 1, \"error\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
