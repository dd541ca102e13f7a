// Measurement: how fast does ONE SM's TMA unit deliver tiled boxes as a function of the box ROW width?  A box of R rows x
// row_bytes is one cp.async.bulk.tensor; the kernels of this repo use rows of 32 / 64 / 128 bytes (= the swizzle span).
// Persistent CTAs (1 per SM), a ring of S in-flight boxes, data L2-resident (each CTA cycles over a small window).
// Prints cycles per row and bytes per cycle per SM.   nvcc -arch=sm_100a -o tma_rate tma_rate.cu -lcuda ; ./tma_rate
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../stereo_3d_reconstruction_b200/csrc/ptx.cuh"
using namespace s3d::ptx;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kMaxStages = 32;

__global__ void __launch_bounds__(64, 1) rate_kernel(const __grid_constant__ CUtensorMap map, int rows_per_box, int box_bytes,
                                                     int iters, int window_boxes, int kStages, int mode, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kMaxStages];
  if (threadIdx.x == 0) { for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1); fence_barrier_init(); }
  __syncthreads();
  // mode bit 0: prefetch.tensormap first; bit 1: TWO warps issue, each on its own half of the ring
  const int nw = (mode & 2) ? 2 : 1, wid = threadIdx.x >> 5;
  if (wid >= nw) return;
  if ((mode & 1) && elect_one()) prefetch_tensormap(&map);
  __syncwarp();
  kStages /= nw;
  const uint32_t sm = smem_u32(smem) + wid * kStages * ((box_bytes + 1023) / 1024 * 1024), f0 = smem_u32(&full[wid * kStages]);
  iters /= nw;
  const int slot = (box_bytes + 1023) / 1024 * 1024;
  const int base = blockIdx.x * window_boxes + wid * (window_boxes / 2);
  window_boxes /= nw;
  long long t0 = 0;
  // prologue: fill the ring
  for (int i = 0; i < kStages; ++i) {
    if (elect_one()) {
      mbar_arrive_expect_tx_u32(f0 + 8 * i, box_bytes);
      tma_load_3d_u32(sm + i * slot, &map, f0 + 8 * i, 0, 0, base + (i % window_boxes));
    }
    __syncwarp();
  }
  uint32_t phase = 0;  int s = 0;
  if (mode & 4) {
    // batch mode: drain the prologue, then per round ONE expect_tx for kStages boxes, kStages back-to-back TMAs, one wait
    for (int i = 0; i < kStages; ++i) mbar_wait_u32(f0 + 8 * i, 0);
    const int slot_b = (box_bytes + 1023) / 1024 * 1024;
    t0 = clock64();
    uint32_t ph = 1;
    for (int i = 0; i < iters; i += kStages) {
      if (elect_one()) {
        mbar_arrive_expect_tx_u32(f0, kStages * box_bytes);
        for (int k = 0; k < kStages; ++k) tma_load_3d_u32(sm + k * slot_b, &map, f0, 0, 0, base + ((i + k) % window_boxes));
      }
      __syncwarp();
      mbar_wait_u32(f0, ph);
      ph ^= 1;
    }
    const long long t1b = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1b - t0;
    return;
  }
  t0 = clock64();
  for (int i = kStages; i < iters + kStages; ++i) {
    if (!(mode & 8)) mbar_wait_u32(f0 + 8 * s, phase);
    if (i < iters) {
      if (elect_one()) {
        if (mode & 16) mbar_arrive_expect_tx_u32(f0 + 8 * s, 0);
        else {
        mbar_arrive_expect_tx_u32(f0 + 8 * s, box_bytes);
        tma_load_3d_u32(sm + s * slot, &map, f0 + 8 * s, 0, 0, base + (i % window_boxes));
        }
      }
      __syncwarp();
    }
    if (++s == kStages) { s = 0; phase ^= 1; }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  void* fp = nullptr;  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  int sms = 0;  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int window = 32, iters = 2000;
  long long* dcyc;  cudaMalloc(&dcyc, sms * sizeof(long long));
  printf("%8s %6s %10s %6s %12s %12s\n", "rowB", "rows", "boxB", "stages", "cyc/row", "B/cyc/SM");
  for (int rb : {64}) {
    for (int rows : {24, 96}) {
      const int C = rb / 2;                                   // bf16 elements per row
      const size_t nbox = (size_t)sms * window;
      __nv_bfloat16* d;  cudaMalloc(&d, nbox * rows * rb);  cudaMemset(d, 0, nbox * rows * rb);
      CUtensorMap map;
      cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)nbox};
      cuuint64_t strides[2] = {(cuuint64_t)rb, (cuuint64_t)rb * rows};
      cuuint32_t box[3] = {(cuuint32_t)C, (cuuint32_t)rows, 1}, estr[3] = {1, 1, 1};
      const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
      CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int box_bytes = rows * rb, slot = (box_bytes + 1023) / 1024 * 1024;
      for (int mode : {0, 8, 16})
      for (int stages : {8}) {
      if (stages * slot > 200 * 1024) continue;
      const int smem = stages * slot + 1024;
      cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      for (int rep = 0; rep < 2; ++rep) rate_kernel<<<sms, 64, smem>>>(map, rows, box_bytes, iters, window, stages, mode, dcyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> h(sms);  cudaMemcpy(h.data(), dcyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      double avg = 0;  for (long long c : h) avg += (double)c;  avg /= sms;
      printf("%8d %6d %10d %4d m%d %9.2f %9.1f  cyc/box %7.1f\n", rb, rows, box_bytes, stages, mode, avg / ((double)iters * rows), (double)iters * box_bytes / avg, avg / iters);
      }
      cudaFree(d);
    }
  }
  return 0;
}
