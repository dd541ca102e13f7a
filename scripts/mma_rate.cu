// Micro-benchmark: tcgen05.mma issue rate for SS-mode bf16 MMAs as a function of N and of the A-descriptor
// stride (dense 1024-B groups vs the 1280-B halo layout).  Standalone: nvcc -arch=sm_100a, run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../stereo_3d_reconstruction_b200/csrc/ptx.cuh"
using namespace s3d::ptx;

__global__ void __launch_bounds__(128) rate_kernel(int N, int a_sbo, int iters, int two_acc, int a_stride, int b_stride, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base, 512);
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  const uint32_t sA = smem_u32(smem), sB = sA + 96 * 1024;
  const uint32_t idesc = make_instr_desc(1, 128, N);
  long long t0 = 0, t1 = 0;
  if (warp == 0 && a_stride == 0 && b_stride == 0) {
    // compile-time operand offsets: the cheapest possible issue loop
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          uint64_t ad = make_smem_desc(sA + (k & 3) * 32 + (k >> 2) * 128 * 9, 128);
          ad = (ad & ~(0x3FFFull << 32)) | ((uint64_t)(a_sbo >> 4) << 32);
          const uint64_t bd = make_smem_desc(sB + (k & 3) * 32 + (k >> 2) * 24576, 128);
          mma_bf16(tm + ((two_acc && (k >> 2)) ? 256u : 0u), ad, bd, idesc, (it | k) ? 1u : 0u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp == 0) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // a_stride / b_stride: bytes between the operand tiles of consecutive MMAs (0 = the same 8 tiles over and over)
          const uint32_t ao = a_stride ? (uint32_t)(((it * 8 + k) * a_stride) & 0xFFF0u) : (k & 3) * 32 + (k >> 2) * 128 * 9;
          const uint32_t bo = b_stride ? (uint32_t)(((it * 8 + k) * b_stride) & 0xFFF0u) : (k & 3) * 32 + (k >> 2) * 24576;
          uint64_t ad = make_smem_desc(sA + ao, 128);
          ad = (ad & ~(0x3FFFull << 32)) | ((uint64_t)(a_sbo >> 4) << 32);
          const uint64_t bd = make_smem_desc(sB + bo, 128);
          const uint32_t d = tm + ((two_acc && (k >> 2)) ? 256u : 0u);
          mma_bf16(d, ad, bd, idesc, (it | k) ? 1u : 0u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d;  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int mode = 0; mode < 4; ++mode)
    for (int sbo : {1280})
      for (int N : {48, 64, 96, 128, 192, 256}) {
        const int two = 1;
        const int a_stride = (mode & 1) ? 1280 + 32 : 0, b_stride = (mode & 2) ? 1024 + 32 : 0;
        rate_kernel<<<148, 128, 200 * 1024>>>(N, sbo, iters, two, a_stride, b_stride, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
        long long h[148];  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double mx = 0;  for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        const double cyc = mx / (iters * 8.0);
        printf("a_distinct=%d b_distinct=%d a_sbo=%d N=%3d: %.1f cycles/MMA  -> %.0f%% of the N/2-cycle math rate\n", (mode & 1), (mode >> 1), sbo, N, cyc,
               100.0 * (N / 2.0) / cyc);
      }
  return 0;
}
