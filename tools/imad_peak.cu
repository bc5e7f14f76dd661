// imad_peak.cu -- measures the B200 int32 multiply-pipe peak that the MSM roofline is quoted
// against (nominal 148 SMs x 64 lanes/clk x 1.965 GHz = 1.861e13 MAC/s).
// Kernels: independent IMAD (32-bit), independent IMAD.WIDE.U32 (32x32+64), carry-chained
// IMAD.WIDE.U32.X (the form the Montgomery multiply uses), and back-to-back fp_mul.
// Prints one JSON object.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_peak imad_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../0g-ec-gpu_b200/csrc/ec.cuh"
using namespace msm;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s\"}\n", cudaGetErrorString(e), #x); return 1; } } while (0)

constexpr int ILP = 8;

__global__ void k_imad32(uint32_t* out, uint32_t a, uint32_t b, int iters, unsigned long long* cycles) {
  uint32_t acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) acc[k] = threadIdx.x + k;
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[k]) : "r"(a), "r"(b));
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void k_imadwide(uint32_t* out, uint32_t a, uint32_t b, int iters, unsigned long long* cycles) {
  uint64_t acc[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) acc[k] = threadIdx.x + k;
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a + k), "r"(b));
  }
  unsigned long long t1 = clock64();
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s ^= acc[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// 2 independent carry chains of 4 wide MACs each per step (like one row of the Montgomery product)
__global__ void k_imadwide_x(uint32_t* out, uint32_t a, uint32_t b, int iters, unsigned long long* cycles) {
  uint32_t e[8], o[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { e[k] = threadIdx.x + k; o[k] = threadIdx.x * 3 + k; }
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    mad_wide_cc(e[0], e[1], a, b);
    madc_wide_cc(e[2], e[3], a + 1, b);
    madc_wide_cc(e[4], e[5], a + 2, b);
    madc_wide_cc(e[6], e[7], a + 3, b);
    mad_wide_cc(o[0], o[1], a + 4, b);
    madc_wide_cc(o[2], o[3], a + 5, b);
    madc_wide_cc(o[4], o[5], a + 6, b);
    madc_wide_cc(o[6], o[7], a + 7, b);
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= e[k] ^ o[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <class P>
__global__ void k_fpmul(uint32_t* out, int iters, unsigned long long* cycles) {
  Fp<P> x, y;
#pragma unroll
  for (int k = 0; k < P::N; k++) { x.v[k] = threadIdx.x + k + 1; y.v[k] = blockIdx.x + 7 * k + 3; }
  x.v[P::N - 1] &= 0x0fffffff; y.v[P::N - 1] &= 0x0fffffff;
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    x = fp_mul<P>(x, y);
    y = fp_mul<P>(y, x);
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < P::N; k++) s ^= x.v[k] ^ y.v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <class F>
__global__ void k_fmul(uint32_t* out, int iters, unsigned long long* cycles) {
  typename F::Elem x = F::one(), y = F::one();
  x.v[0] += threadIdx.x; y.v[1] += blockIdx.x;
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    x = F::mul(x, y);
    y = F::mul(y, x);
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < F::N; k++) s ^= x.v[k] ^ y.v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// back-to-back mixed additions of two alternating points into one accumulator per thread
template <class F>
__global__ void __launch_bounds__(128) k_madd(uint32_t* out, int iters, unsigned long long* cycles) {
  Affine<F> p0, p1;
  p0.x = F::one(); p0.y = F::one(); p1.x = F::one(); p1.y = F::one();
  p0.x.v[0] ^= threadIdx.x + 1; p0.y.v[1] ^= blockIdx.x + 3; p1.x.v[2] ^= threadIdx.x * 7 + 5; p1.y.v[0] ^= 11;
  Xyzz<F> acc = xyzz_inf<F>();
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    xyzz_madd<F>(acc, p0);
    xyzz_madd<F>(acc, p1);
  }
  unsigned long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < F::N; k++) s ^= acc.x.v[k] ^ acc.y.v[k] ^ acc.zz.v[k] ^ acc.zzz.v[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <class K, class... A>
static int run(const char* name, double macs_per_thread_iter, int iters, int blocks_per_sm, int threads, int sms,
               uint32_t* out, unsigned long long* cyc, K kernel, A... args) {
  const int grid = sms * blocks_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kernel<<<grid, threads>>>(out, args..., iters / 10 + 1, cyc);  // warm-up
  CHECK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  kernel<<<grid, threads>>>(out, args..., iters, cyc);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long* h = new unsigned long long[grid];
  cudaMemcpy(h, cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (int i = 0; i < grid; i++) mean += (double)h[i];
  mean /= grid;
  delete[] h;
  const double total = macs_per_thread_iter * iters * (double)threads * grid;
  // all blocks_per_sm blocks of an SM run concurrently for ~mean cycles
  const double per_clk_sm = macs_per_thread_iter * iters * (double)threads * blocks_per_sm / mean;
  printf("  \"%s\": {\"ms\": %.3f, \"mac_per_s\": %.4e, \"mac_per_clk_per_sm\": %.2f, \"block_cycles\": %.0f, \"eff_mhz\": %.0f, \"warps_per_sm\": %d},\n",
         name, ms, total / (ms * 1e-3), per_clk_sm, mean, mean / (ms * 1e3), blocks_per_sm * threads / 32);
  return 0;
}

int main() {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  uint32_t* out; unsigned long long* cyc;
  CHECK(cudaMalloc(&out, (size_t)sms * 8 * 1024 * 4));
  CHECK(cudaMalloc(&cyc, (size_t)sms * 8 * 8));
  printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", prop.name, sms, prop.clockRate);
  const int it = 20000;
  for (int bps = 2; bps <= 4; bps *= 2) {
    char nm[64];
    snprintf(nm, sizeof nm, "imad32_%dw", bps * 8);      run(nm, ILP, it, bps, 256, sms, out, cyc, k_imad32, 0x9e3779b9u, 0x7f4a7c15u);
    snprintf(nm, sizeof nm, "imad_wide_%dw", bps * 8);   run(nm, ILP, it, bps, 256, sms, out, cyc, k_imadwide, 0x9e3779b9u, 0x7f4a7c15u);
    snprintf(nm, sizeof nm, "imad_wide_x_%dw", bps * 8); run(nm, 8, it, bps, 256, sms, out, cyc, k_imadwide_x, 0x9e3779b9u, 0x7f4a7c15u);
  }
  for (int bps = 1; bps <= 4; bps *= 2) {
    char nm[64];
    snprintf(nm, sizeof nm, "fp_mul_bn254_%dw", bps * 4);
    run(nm, 2.0 * 136, 2000, bps, 128, sms, out, cyc, k_fpmul<Bn254Fq>);
    snprintf(nm, sizeof nm, "fp_mul_bls381_%dw", bps * 4);
    run(nm, 2.0 * 300, 1000, bps, 128, sms, out, cyc, k_fpmul<Bls381Fq>);
  }
  run("fp_mul_bn254_32w", 2.0 * 136, 2000, 4, 256, sms, out, cyc, k_fpmul<Bn254Fq>);
  // the lazy 29-bit field, counted at the same ALGORITHMIC 136 MAC per product
  for (int bps = 1; bps <= 8; bps *= 2) {
    char nm[64];
    snprintf(nm, sizeof nm, "fmul_u29_bn254_%dw", bps * 4);
    run(nm, 2.0 * 136, 2000, bps, 128, sms, out, cyc, k_fmul<FieldU29<Bn254U29>>);
  }
  // carry-save saturated product (fp_mul_cs): no IMAD.WIDE.U32.X at all
  for (int bps = 1; bps <= 8; bps *= 2) {
    char nm[64];
    snprintf(nm, sizeof nm, "fmul_cs_bn254_%dw", bps * 4);
    run(nm, 2.0 * 136, 2000, bps, 128, sms, out, cyc, k_fmul<FieldSat<Bn254Fq, true>>);
  }
  run("fmul_cs_bls381_8w", 2.0 * 300, 1000, 2, 128, sms, out, cyc, k_fmul<FieldSat<Bls381Fq, true>>);
  run("fmul_cs_bls381_16w", 2.0 * 300, 1000, 4, 128, sms, out, cyc, k_fmul<FieldSat<Bls381Fq, true>>);
  for (int bps = 1; bps <= 4; bps *= 2) {
    char nm[64];
    snprintf(nm, sizeof nm, "madd_cs_bn254_%dw", bps * 4);
    run(nm, 2.0 * 1360, 300, bps, 128, sms, out, cyc, k_madd<FieldSat<Bn254Fq, true>>);
  }
  run("madd_cs_bls381_8w", 2.0 * 3000, 150, 2, 128, sms, out, cyc, k_madd<FieldSat<Bls381Fq, true>>);
  // values in [0, 2p), no conditional subtraction after products, fused r*d - y*ppp
  for (int bps = 1; bps <= 4; bps *= 2) {
    char nm[64];
    snprintf(nm, sizeof nm, "madd_lazy_bn254_%dw", bps * 4);
    run(nm, 2.0 * 1360, 300, bps, 128, sms, out, cyc, k_madd<FieldSatLazy<Bn254Fq>>);
  }
  run("madd_lazy_bls381_8w", 2.0 * 3000, 150, 2, 128, sms, out, cyc, k_madd<FieldSatLazy<Bls381Fq>>);
  // mixed addition, counted at the algorithmic 10 products x 136 MAC
  for (int bps = 1; bps <= 4; bps *= 2) {
    char nm[64];
    snprintf(nm, sizeof nm, "madd_sat32_bn254_%dw", bps * 4);
    run(nm, 2.0 * 1360, 300, bps, 128, sms, out, cyc, k_madd<FieldSat<Bn254Fq>>);
    snprintf(nm, sizeof nm, "madd_u29_bn254_%dw", bps * 4);
    run(nm, 2.0 * 1360, 300, bps, 128, sms, out, cyc, k_madd<FieldU29<Bn254U29>>);
  }
  run("madd_sat32_bls381_4w", 2.0 * 3000, 150, 1, 128, sms, out, cyc, k_madd<FieldSat<Bls381Fq>>);
  run("madd_sat32_bls381_8w", 2.0 * 3000, 150, 2, 128, sms, out, cyc, k_madd<FieldSat<Bls381Fq>>);
  printf("  \"nominal_mac_per_s\": %.4e\n}\n", (double)sms * 64 * 1.965e9);
  return 0;
}
