// pipe_model.cu -- which integer instructions overlap on a B200 SM sub-partition?
// Each kernel runs ILP independent streams per thread; rates are per SM per clock from event time
// at the measured max clock.  Development tool; results summarised in DESIGN.md.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("error %s at %s\n", cudaGetErrorString(e), #x); return 1; } } while (0)
constexpr int ILP = 8;

template <int MODE>
__global__ void k(uint32_t* out, uint32_t a, uint32_t b, int iters) {
  uint32_t x[ILP], y[ILP];
  uint64_t w[ILP];
#pragma unroll
  for (int k = 0; k < ILP; k++) { x[k] = threadIdx.x + k; y[k] = threadIdx.x * 3 + k; w[k] = x[k]; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) {
      if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(a + k), "r"(b));
      if (MODE == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));
      if (MODE == 2) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(a + k), "r"(b));
                       asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(a)); }
      if (MODE == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(a), "r"(b));
      if (MODE == 4) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(a + k), "r"(b));
                       asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(a), "r"(b)); }
      if (MODE == 5) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(x[k]) : "r"(y[k]));
      if (MODE == 6) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(a + k), "r"(b));
                       asm volatile("add.u32 %0, %0, %1;" : "+r"(x[k]) : "r"(a));
                       asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(b)); }
      if (MODE == 7) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[k]), "+r"(y[k]) : "r"(a), "r"(b));
      if (MODE == 8) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(x[k]), "+r"(y[k]) : "r"(a + k), "r"(b));
      if (MODE == 9) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[k]) : "r"(a + k), "r"(b));
      if (MODE == 10) { asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[k]) : "r"(a + k), "r"(b));
                        asm volatile("add.u32 %0, %0, %1;" : "+r"(y[k]) : "r"(a)); }
      if (MODE == 11) { float f = __uint_as_float(x[k]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); x[k] = __float_as_uint(f); }
      if (MODE == 12) { float f = __uint_as_float(x[k]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); x[k] = __float_as_uint(f);
                        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(a + k), "r"(b)); }
      if (MODE == 13) { double d = __longlong_as_double(w[k]); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(1.0000001), "d"(0.5)); w[k] = __double_as_longlong(d); }
      if (MODE == 14) { double d = __longlong_as_double(w[k]); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d) : "d"(1.0000001), "d"(0.5)); w[k] = __double_as_longlong(d);
                        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[k]) : "r"(a + k), "r"(b)); }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < ILP; k++) s ^= x[k] ^ y[k] ^ (uint32_t)w[k] ^ (uint32_t)(w[k] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> int run(const char* name, int n_instr, uint32_t* out, int sms) {
  const int iters = 4000, threads = 256, bps = 4;
  k<MODE><<<sms * bps, threads>>>(out, 0x9e3779b9u, 0x7f4a7c15u, iters / 10);
  CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sms * bps, threads>>>(out, 0x9e3779b9u, 0x7f4a7c15u, iters);
  cudaEventRecord(e1);
  CHECK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double groups = (double)iters * ILP * threads * bps * sms;  // per-thread "groups" executed
  double per_clk_sm = groups / (ms * 1e-3) / sms / 1.965e9;
  printf("  \"%s\": {\"ms\": %.3f, \"groups_per_clk_per_sm\": %.2f, \"instr_per_group\": %d, \"thread_instr_per_clk_per_sm\": %.2f},\n",
         name, ms, per_clk_sm, n_instr, per_clk_sm * n_instr);
  return 0;
}

int main() {
  cudaDeviceProp prop; CHECK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  uint32_t* out; CHECK(cudaMalloc(&out, (size_t)sms * 4 * 256 * 4));
  printf("{\n");
  run<0>("imad_wide", 1, out, sms);
  run<1>("iadd", 1, out, sms);
  run<2>("imad_wide+iadd", 2, out, sms);
  run<3>("lop3", 1, out, sms);
  run<4>("imad_wide+lop3", 2, out, sms);
  run<5>("shf", 1, out, sms);
  run<6>("imad_wide+2iadd", 3, out, sms);
  run<7>("add.cc+addc (64-bit add)", 2, out, sms);
  run<8>("mad.lo.cc+madc.hi (wide, carry inside)", 1, out, sms);
  run<9>("imad32", 1, out, sms);
  run<10>("imad32+iadd", 2, out, sms);
  run<11>("ffma", 1, out, sms);
  run<12>("ffma+imad_wide", 2, out, sms);
  run<13>("dfma", 1, out, sms);
  run<14>("dfma+imad32", 2, out, sms);
  printf("  \"note\": \"groups = one pass over the listed instruction(s) per thread; 64 = one 16-lane pipe per SMSP\"\n}\n");
  return 0;
}
