// dfma_mul.cu -- is the FP64 pipe worth using for the field product on B200?
// Measures (1) the DFMA-based Montgomery product of csrc/f52.cuh alone, (2) the IMAD-based product
// of csrc/fp.cuh alone, (3) both at once in different warps of the same blocks (k of every 8 warps
// run the DFMA form), and writes sample (a, b, r) triples for an exactness check against Python
// big integers (tools/jobs/check_dfma.py).  Prints one JSON object.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o dfma_mul dfma_mul.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../0g-ec-gpu_b200/csrc/ec.cuh"
#include "f52.cuh"
using namespace msm;

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s\"}\n", cudaGetErrorString(e), #x); return 1; } } while (0)

template <class Q> struct D52 { double v[Q::N]; };

template <class Q> __device__ __forceinline__ D52<Q> mul52(const D52<Q>& a, const D52<Q>& b) {
  uint64_t r[Q::N];
  f52_mul_core<Q>(r, a.v, b.v);
  D52<Q> o;
#pragma unroll
  for (int i = 0; i < Q::N; i++) o.v[i] = u52_to_double(r[i]);
  return o;
}

template <class Q>
__global__ void k_check(const uint64_t* a, const uint64_t* b, uint64_t* r, int n) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  double da[Q::N], db[Q::N];
  for (int i = 0; i < Q::N; i++) { da[i] = u52_to_double(a[t * Q::N + i]); db[i] = u52_to_double(b[t * Q::N + i]); }
  uint64_t o[Q::N];
  f52_mul_core<Q>(o, da, db);
  for (int i = 0; i < Q::N; i++) r[t * Q::N + i] = o[i];
}

// kdf of every 8 warps run the DFMA chain (iters_d products x2), the others the IMAD chain (iters_i x2)
template <class P, class Q>
__global__ void __launch_bounds__(256) k_mixed(uint32_t* out, int kdf, int iters_i, int iters_d) {
  const int warp = threadIdx.x >> 5;
  uint32_t s = 0;
  if ((warp & 7) < kdf) {
    D52<Q> x, y;
#pragma unroll
    for (int k = 0; k < Q::N; k++) { x.v[k] = (double)(threadIdx.x + k + 1); y.v[k] = (double)(blockIdx.x + 7 * k + 3); }
    for (int i = 0; i < iters_d; i++) {
      x = mul52<Q>(x, y);
      y = mul52<Q>(y, x);
    }
#pragma unroll
    for (int k = 0; k < Q::N; k++) s ^= (uint32_t)__double_as_longlong(x.v[k]) ^ (uint32_t)__double_as_longlong(y.v[k]);
  } else {
    Fp<P> x, y;
#pragma unroll
    for (int k = 0; k < P::N; k++) { x.v[k] = threadIdx.x + k + 1; y.v[k] = blockIdx.x + 7 * k + 3; }
    x.v[P::N - 1] &= 0x0fffffff; y.v[P::N - 1] &= 0x0fffffff;
    for (int i = 0; i < iters_i; i++) {
      x = fp_mul_nored<P>(x, y);
      y = fp_mul_nored<P>(y, x);
    }
#pragma unroll
    for (int k = 0; k < P::N; k++) s ^= x.v[k] ^ y.v[k];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static uint64_t rng_state = 0x243F6A8885A308D3ull;
static uint64_t rng() {
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <class P, class Q> int run(const char* name, int sms, double clk_hz, FILE* samples) {
  // exactness samples: random limbs < 2^52, top limb bounded so that the value is < 2p
  const int n = 4096;
  std::vector<uint64_t> ha(n * Q::N), hb(n * Q::N), hr(n * Q::N);
  for (int t = 0; t < n; t++)
    for (int i = 0; i < Q::N; i++) {
      const uint64_t top = 2 * Q::P(Q::N - 1);  // < 2p
      ha[t * Q::N + i] = i == Q::N - 1 ? rng() % top : (t == 0 ? M52 : rng() & M52);
      hb[t * Q::N + i] = i == Q::N - 1 ? rng() % top : (t <= 1 ? M52 : rng() & M52);
    }
  uint64_t *da, *db, *dr;
  CHECK(cudaMalloc(&da, ha.size() * 8)); CHECK(cudaMalloc(&db, ha.size() * 8)); CHECK(cudaMalloc(&dr, ha.size() * 8));
  CHECK(cudaMemcpy(da, ha.data(), ha.size() * 8, cudaMemcpyHostToDevice));
  CHECK(cudaMemcpy(db, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice));
  k_check<Q><<<(n + 127) / 128, 128>>>(da, db, dr, n);
  CHECK(cudaDeviceSynchronize());
  CHECK(cudaMemcpy(hr.data(), dr, hr.size() * 8, cudaMemcpyDeviceToHost));
  // host build of the same core (integer emulation of the fma pair) must agree limb for limb
  int host_mismatch = 0;
  for (int t = 0; t < n; t++) {
    double xa[Q::N], xb[Q::N];
    uint64_t o[Q::N];
    for (int i = 0; i < Q::N; i++) { xa[i] = (double)ha[t * Q::N + i]; xb[i] = (double)hb[t * Q::N + i]; }
    f52_mul_core<Q>(o, xa, xb);
    for (int i = 0; i < Q::N; i++) host_mismatch += o[i] != hr[t * Q::N + i];
  }
  if (samples) {
    for (int t = 0; t < 256; t++) {
      fprintf(samples, "%s %d", name, Q::N);
      for (int i = 0; i < Q::N; i++) fprintf(samples, " %llx", (unsigned long long)ha[t * Q::N + i]);
      for (int i = 0; i < Q::N; i++) fprintf(samples, " %llx", (unsigned long long)hb[t * Q::N + i]);
      for (int i = 0; i < Q::N; i++) fprintf(samples, " %llx", (unsigned long long)hr[t * Q::N + i]);
      fprintf(samples, "\n");
    }
  }
  uint32_t* out;
  const int blocks = sms * 4, threads = 256;
  CHECK(cudaMalloc(&out, (size_t)blocks * threads * 4));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time_it = [&](int kdf, int ii, int id) -> float {
    k_mixed<P, Q><<<blocks, threads>>>(out, kdf, ii / 8, id / 8);  // warm-up
    cudaEventRecord(e0);
    k_mixed<P, Q><<<blocks, threads>>>(out, kdf, ii, id);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
  };
  const int IT = 2000;
  const double thr = (double)blocks * threads;
  const float ms_i = time_it(0, IT, 0), ms_d = time_it(8, 0, IT);
  const double rate_i = thr * 2.0 * IT / (ms_i * 1e-3), rate_d = thr * 2.0 * IT / (ms_d * 1e-3);
  printf("\"%s\": {\"host_device_limb_mismatches\": %d, \"imad_products_per_s\": %.4g, \"imad_clk_per_product_per_sm\": %.3f, "
         "\"dfma_products_per_s\": %.4g, \"dfma_clk_per_product_per_sm\": %.3f, \"mixed\": [",
         name, host_mismatch, rate_i, clk_hz * sms / rate_i, rate_d, clk_hz * sms / rate_d);
  for (int kdf = 1; kdf < 8; kdf++) {
    // per-warp iteration counts proportional to each form's rate when it owns its pipe alone, so
    // that both kinds of warps finish together if the pipes really run side by side
    const double per_warp_i = rate_i / 8.0, per_warp_d = rate_d / 8.0;
    const double scale = IT / (per_warp_i > per_warp_d ? per_warp_i : per_warp_d);
    const int ii = (int)(per_warp_i * scale), id = (int)(per_warp_d * scale);
    const float ms = time_it(kdf, ii, id);
    const double prods = (double)blocks * 32.0 * 2.0 * ((8 - kdf) * (double)ii + kdf * (double)id);
    printf("%s{\"dfma_warps_of_8\": %d, \"products_per_s\": %.4g, \"vs_imad_alone\": %.3f}", kdf > 1 ? ", " : "", kdf,
           prods / (ms * 1e-3), prods / (ms * 1e-3) / rate_i);
  }
  printf("]}");
  cudaFree(da); cudaFree(db); cudaFree(dr); cudaFree(out);
  return 0;
}

int main(int argc, char** argv) {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  FILE* samples = argc > 1 ? fopen(argv[1], "w") : nullptr;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %d, ", prop.name, prop.multiProcessorCount, clk_khz / 1000);
  if (run<Bn254Fq, Bn254Fq52>("bn254", prop.multiProcessorCount, clk_khz * 1e3, samples)) return 1;
  printf(", ");
  if (run<Bls381Fq, Bls381Fq52>("bls12_381", prop.multiProcessorCount, clk_khz * 1e3, samples)) return 1;
  printf("}\n");
  if (samples) fclose(samples);
  return 0;
}
