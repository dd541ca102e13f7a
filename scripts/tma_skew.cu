// Feasibility check: a TMA tensor map whose X and D dimensions have the SAME stride (a skewed view x' = x - d of a
// zero-padded feature row), negative coordinates -> zero fill.  Standalone: nvcc -arch=sm_100a, run on a B200.
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

__global__ void load_kernel(const __grid_constant__ CUtensorMap map, int x0, int d, int y0, float* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar, 32 * 2 * 10 * 4);
    tma_load_5d(smem, &map, &bar, 0, x0, d, y0, 0);
  }
  mbar_wait(&bar, 0);
  const __nv_bfloat16* s = reinterpret_cast<const __nv_bfloat16*>(smem);
  for (int i = threadIdx.x; i < 40; i += blockDim.x) out[i] = __bfloat162float(s[i * 32]);   // channel 0 of each (y, x) row
}

int main() {
  const int C = 32, W = 16, PAD = 8, H = 6, D = 8, P = W + 2 * PAD;       // row pitch P pixels, zero pads on both sides
  std::vector<__nv_bfloat16> h((size_t)H * P * C, __float2bfloat16(0.f));
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int c = 0; c < C; ++c)
    h[((size_t)y * P + PAD + x) * C + c] = __float2bfloat16((float)(100 * y + x));
  __nv_bfloat16* dptr;  cudaMalloc(&dptr, h.size() * 2);  cudaMemcpy(dptr, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr;  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  // dims (C, X', D, Y, N): value(c, x', d, y) = feat[y, x' + d]  (base at the first real pixel); X' bound W, D bound D
  CUtensorMap map;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)D, (cuuint64_t)H, 1};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 2, (cuuint64_t)P * C * 2, (cuuint64_t)H * P * C * 2};
  cuuint32_t box[5] = {(cuuint32_t)C, 10, 1, 4, 1}, estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dptr + (size_t)PAD * C, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode (equal X / D strides): %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 1;
  float* dout;  cudaMalloc(&dout, 40 * 4);
  for (int d : {0, 3}) {
    // left-masked reference half: want feat[y, x] for x >= d else 0, x = x0-1 .. x0+8  ->  x' = x - d
    const int x0 = 0, y0 = 1;
    load_kernel<<<1, 64, 4096>>>(map, x0 - 1 - d, d, y0, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    float o[40];  cudaMemcpy(o, dout, sizeof(o), cudaMemcpyDeviceToHost);
    printf("d=%d: row y=%d:", d, y0);
    for (int i = 0; i < 10; ++i) printf(" %g", o[i]);
    printf("   (expect 0 for x<%d, then 100*y+x)\n", d);
  }
  return 0;
}
