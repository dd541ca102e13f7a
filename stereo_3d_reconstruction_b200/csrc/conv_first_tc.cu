// First layer of both 2-D encoders on the tensor core (bf16 storage, 32 output channels): 3x3, stride 2, pad 1 on the raw image
// (+ the disparity channel for the RGB-D encoder).  conv_first.cu computes it with fp32 FMAs -- K = 27 / 36 is no shape for an
// implicit GEMM fed by TMA -- and is bound by FMA issue (864-1152 FMAs per output pixel, 0.074-0.094 ms per 128 images against
// 0.036 ms of HBM time).  Here the threads build the im2col operand themselves:
//   1. the input patch of a 4 x 32 output tile (9 x 65 pixels) is staged in shared memory as bf16 [9][66][4]: coalesced reads of
//      the NCHW fp32 planes / the HWC bytes (+ the disparity plane), converted and rounded exactly as conv_first.cu rounds them;
//   2. thread = output pixel = operand row: 9 taps x 4 channel slots (k = tap * 4 + ci, slot 3 zero without disparity) = 72 bytes
//      of a 128-byte K-major row, written in the 128B-swizzled layout of the UMMA descriptor;
//   3. one thread issues 3 tcgen05 MMAs (M = 128 pixels, N = 32 channels, K = 48) against the weight tile, which every CTA packs
//      once into the same layout; the accumulator (32 TMEM columns) comes back to the thread that owns the pixel;
//   4. bias + activation in fp32, 64 contiguous bytes per pixel stored.
// Same arithmetic as the path it replaces: inputs and weights rounded to bf16, products exact, fp32 accumulation (in another
// order), output rounded to bf16.  No warp specialisation: a tile is a short chain (stage -> build -> MMA -> drain) and EIGHT
// CTAs of 128 threads share an SM (32 TMEM columns, 25 KB of shared memory each) to overlap their chains.
#include <string.h>
#include "common.cuh"
#include "ptx.cuh"

namespace s3d {
namespace {

constexpr int kTR = 4, kTC = 32;                 // output tile: 4 rows x 32 columns = 128 pixels
constexpr int kPR = 2 * kTR + 1, kPC = 2 * kTC + 1;   // input patch 9 x 65
constexpr int kPP = 66;                          // patch row pitch (pixels)
constexpr int kCo = 32;

struct FtArgs {
  const void* img;  const float* disp;  const __nv_bfloat16* w;  const float* bias;  __nv_bfloat16* out;
  float disp_scale, img_scale, slope;
  int B, H, W, oH, oW, cin_pad, tiles_x, tiles_y, total;
  uint32_t idesc;
};

template <int CIN, bool kU8>
__global__ void __launch_bounds__(128, 8)
conv_first_tc_kernel(const __grid_constant__ FtArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));   // operand rows (pixels)
  uint8_t* sB = sA + 128 * 128;                            // weight rows (output channels); both 128B-swizzled
  uint2* patch = reinterpret_cast<uint2*>(sB + kCo * 128); // bf16 x 4 per input pixel
  __shared__ __align__(16) float sbias[kCo];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;

  // ---- once per CTA: weight tile B[co][tap * 4 + ci], bias, barrier, TMEM ----
  for (int i = t; i < kCo * 128 / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int i = t; i < kCo * 9 * CIN; i += 128) {
    const int ci = i % CIN, tap = (i / CIN) % 9, co = i / (9 * CIN);
    const int k = tap * 4 + ci;
    *reinterpret_cast<__nv_bfloat16*>(sB + co * 128 + ((((k >> 3) ^ (co & 7))) << 4) + (k & 7) * 2) =
        a.w[((int64_t)tap * kCo + co) * a.cin_pad + ci];
  }
  if (t < kCo) sbias[t] = a.bias ? a.bias[t] : 0.f;
  if (t == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc(&tmem_slot, 32);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint64_t hi = (static_cast<uint64_t>((8 * 128) >> 4) << 32) | (1ull << 46) | (2ull << 61);   // dense 128-byte rows, 128B swizzle
  const uint64_t adesc = hi | (((ptx::smem_u32(sA) & 0x3FFFF) >> 4) | (1u << 16));
  const uint64_t bdesc = hi | (((ptx::smem_u32(sB) & 0x3FFFF) >> 4) | (1u << 16));
  const uint32_t bar_u = ptx::smem_u32(&bar);
  const int r = t >> 5, c = t & 31;                       // this thread's pixel inside the tile
  uint8_t* arow = sA + t * 128;
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < a.total; tile += gridDim.x) {
    const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, n = tile / (a.tiles_x * a.tiles_y);
    const int oy0 = ty * kTR, ox0 = tx * kTC;
    const int iy0 = 2 * oy0 - 1, ix0 = 2 * ox0 - 1;
    // ---- 1. stage the input patch (zero outside the image): all loads of the thread's 5 pixels in flight before any is used.
    // (Issuing the NEXT tile's loads before waiting for this tile's MMAs was tried: no gain -- with eight CTAs per SM the
    // kernel already moves 4 TB/s on fp32 images, 0.056-0.068 ms per 128 images against 0.074-0.087 for the SIMT kernel.)
    {
      constexpr int kIt = (kPR * kPC + 127) / 128;
      float v[kIt][4];
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int i = t + it * 128;
        const int pr = i / kPC, pc = i - pr * kPC;
        const int iy = iy0 + pr, ix = ix0 + pc;
        v[it][0] = v[it][1] = v[it][2] = v[it][3] = 0.f;
        if (i < kPR * kPC && iy >= 0 && iy < a.H && ix >= 0 && ix < a.W) {
          if (kU8) {
            const uint8_t* ip = reinterpret_cast<const uint8_t*>(a.img) + (((int64_t)n * a.H + iy) * a.W + ix) * 3;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) v[it][ci] = (float)__ldg(ip + ci);
          } else {
            const float* ip = reinterpret_cast<const float*>(a.img) + ((int64_t)n * 3 * a.H + iy) * a.W + ix;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) v[it][ci] = __ldg(ip + ci * (int64_t)a.H * a.W);
          }
          if (CIN > 3) v[it][3] = __ldg(a.disp + ((int64_t)n * a.H + iy) * a.W + ix);
        }
      }
#pragma unroll
      for (int it = 0; it < kIt; ++it) {
        const int i = t + it * 128;
        if (i < kPR * kPC) {
          const int pr = i / kPC, pc = i - pr * kPC;
          const float s3 = kU8 ? a.img_scale : 1.f;
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[it][0] * s3, v[it][1] * s3);
          const __nv_bfloat162 p1 = __floats2bfloat162_rn(v[it][2] * s3, v[it][3] * a.disp_scale);
          patch[pr * kPP + pc] = make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
        }
      }
    }
    __syncthreads();
    // ---- 2. this pixel's operand row: taps (ky, kx) -> 8 bytes each, 72 bytes + zeros up to 96 (K = 48) ----
    {
      uint2 q[12];
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) q[ky * 3 + kx] = patch[(2 * r + ky) * kPP + 2 * c + kx];
      q[9] = q[10] = q[11] = make_uint2(0, 0);
#pragma unroll
      for (int j = 0; j < 6; ++j)
        *reinterpret_cast<uint4*>(arow + ((j ^ (t & 7)) << 4)) = make_uint4(q[2 * j].x, q[2 * j].y, q[2 * j + 1].x, q[2 * j + 1].y);
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    // ---- 3. M = 128, N = 32, K = 3 x 16 ----
    if (warp == 0) {
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < 3; ++k) ptx::mma_bf16(tmem, adesc + 2 * k, bdesc + 2 * k, a.idesc, k ? 1u : 0u);
        ptx::tc_commit_u32(bar_u);
      }
      __syncwarp();
    }
    ptx::mbar_wait_u32(bar_u, phase);
    phase ^= 1;
    ptx::tc_fence_after();
    // ---- 4. drain: thread t owns TMEM lane t ----
    uint32_t v0[16], v1[16];
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    ptx::tmem_ld16(taddr, v0);
    ptx::tmem_ld16(taddr + 16, v1);
    ptx::tmem_ld_wait();
    const int oy = oy0 + r, ox = ox0 + c;
    if (oy < a.oH && ox < a.oW) {
      uint4 o4[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t wv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ch = 8 * j + 2 * e;
          float f0 = __uint_as_float(ch < 16 ? v0[ch] : v1[ch - 16]) + sbias[ch];
          float f1 = __uint_as_float(ch + 1 < 16 ? v0[ch + 1] : v1[ch + 1 - 16]) + sbias[ch + 1];
          f0 = fmax_nan(f0, 0.f) + a.slope * fmin_nan(f0, 0.f);
          f1 = fmax_nan(f1, 0.f) + a.slope * fmin_nan(f1, 0.f);
          const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
          wv[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        o4[j] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      }
      uint4* o = reinterpret_cast<uint4*>(a.out + (((int64_t)n * a.oH + oy) * a.oW + ox) * kCo);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = o4[j];
    }
    ptx::tc_fence_before();
    __syncthreads();                                        // the accumulator and both operand buffers are free again
  }
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 32); }
}

}  // namespace

// bf16, cout_pad = 32, ReLU / none / LeakyReLU: the tensor-core first layer; anything else stays on conv_first.cu's SIMT kernel.
bool conv_first_tc_eligible(int cout_pad, int dtype, int act) {
  return dtype == S3D_DTYPE_BF16 && cout_pad == kCo && (act == S3D_ACT_NONE || act == S3D_ACT_RELU || act == S3D_ACT_LEAKY) &&
         !knobs().no_conv_first_tc;
}

int conv_first_tc_launch(const void* img, int img_u8, const float* disp, float disp_scale, float img_scale, const void* w,
                         int cin_pad, const float* bias, void* out, int B, int H, int W, int oH, int oW, int cin, int act,
                         float act_param, cudaStream_t st) {
  FtArgs a;
  memset(&a, 0, sizeof(a));
  a.img = img;  a.disp = disp;  a.w = static_cast<const __nv_bfloat16*>(w);  a.bias = bias;  a.out = static_cast<__nv_bfloat16*>(out);
  a.disp_scale = disp_scale;  a.img_scale = img_scale;
  a.slope = act == S3D_ACT_NONE ? 1.f : (act == S3D_ACT_LEAKY ? act_param : 0.f);
  a.B = B;  a.H = H;  a.W = W;  a.oH = oH;  a.oW = oW;  a.cin_pad = cin_pad;
  a.tiles_x = ceil_div(oW, kTC);  a.tiles_y = ceil_div(oH, kTR);
  const int64_t total = (int64_t)B * a.tiles_x * a.tiles_y;
  S3D_CHECK_ARG(total > 0 && total < (1ll << 31), "conv_first: tile count out of range");
  a.total = (int)total;
  a.idesc = ptx::make_instr_desc(1, 128, kCo);
  int grid = num_sms() * 8;
  if (grid > a.total) grid = a.total;
  const int smem = 128 * 128 + kCo * 128 + kPR * kPP * 8 + 1024;     // 26.3 KB: eight CTAs per SM
  if (cin == 3) { if (img_u8) conv_first_tc_kernel<3, true><<<grid, 128, smem, st>>>(a); else conv_first_tc_kernel<3, false><<<grid, 128, smem, st>>>(a); }
  else          { if (img_u8) conv_first_tc_kernel<4, true><<<grid, 128, smem, st>>>(a); else conv_first_tc_kernel<4, false><<<grid, 128, smem, st>>>(a); }
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace s3d
