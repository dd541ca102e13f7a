// Issue-rate probe for the packed fp32 instructions of sm_100 (FADD2 / FMUL2 / FFMA2) against the scalar ones and FMNMX3.
// Every thread runs 8 independent dependency chains of one instruction kind; 8 CTAs x 256 threads per SM.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/f32x2_rate scripts/f32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
template <int KIND>
__global__ void __launch_bounds__(256) k(float* sink, int iters, float a, float b) {
  float r[16];
  for (int i = 0; i < 16; ++i) r[i] = (float)(threadIdx.x + i);
  u64 pa, pb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(pa) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(pb) : "f"(b));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (KIND == 0) r[c] = fmaf(r[c], a, b);
      if (KIND == 1) r[c] = __fadd_rn(r[c], a);
      if (KIND == 2) r[c] = __fmul_rn(r[c], a);
      if (KIND == 3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(r[c]) : "f"(a), "f"(b));
      if (KIND == 7) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[c]));
      if (KIND == 8) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(r[c]));
      if (KIND >= 4 && KIND <= 6) {
        u64 v;
        asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(r[2 * c]), "f"(r[2 * c + 1]));
        if (KIND == 4) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(pa), "l"(pb));
        if (KIND == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(pa));
        if (KIND == 6) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(pa));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(r[2 * c]), "=f"(r[2 * c + 1]) : "l"(v));
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += r[i];
  if (s == 12345.678f) sink[0] = s;
}
template <int KIND> void run(const char* name, float* sink, int sms) {
  cudaEvent_t e0, e1;  cudaEventCreate(&e0);  cudaEventCreate(&e1);
  const int iters = 20000, blocks = sms * 8;
  k<KIND><<<blocks, 256>>>(sink, 100, 1.0000001f, 1e-7f);
  cudaEventRecord(e0);
  k<KIND><<<blocks, 256>>>(sink, iters, 1.0000001f, 1e-7f);
  cudaEventRecord(e1);  cudaEventSynchronize(e1);
  float ms;  cudaEventElapsedTime(&ms, e0, e1);
  const double instr = (double)blocks * 256 / 32 * 8.0 * iters;       // warp instructions
  printf("%-9s %8.3f ms  %7.2f warp-instr/clk/SM at 1.9 GHz (%.3e warp-instr/s)\n", name, ms, instr / (ms * 1e-3) / sms / 1.9e9, instr / (ms * 1e-3));
}
int main() {
  int sms;  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* sink;  cudaMalloc(&sink, 4);
  run<0>("FFMA", sink, sms);  run<1>("FADD", sink, sms);  run<2>("FMUL", sink, sms);  run<3>("FMNMX3", sink, sms);
  run<7>("MUFU.EX2", sink, sms);  run<8>("MUFU.RCP", sink, sms);
  run<4>("FFMA2", sink, sms);  run<5>("FADD2", sink, sms);  run<6>("FMUL2", sink, sms);
  return 0;
}
