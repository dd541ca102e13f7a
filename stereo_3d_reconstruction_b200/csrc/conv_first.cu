// First layer of both 2-D encoders: 3x3, stride 2, pad 1 on the raw image (3 channels, +1 disparity channel for the
// RGB-D encoder).  K = 27 or 36 is no shape for the tensor core: through the implicit-GEMM engine the layer cost
// 0.34 ms (on a zero-padded 16-channel copy of the image that a staging kernel had to write first).  Here one thread
// computes two adjacent output pixels x all output channels with fp32 FMAs straight from the NCHW fp32 (or HWC uint8) image:
// the layer reads the image once and writes its channels-last output once.
// Arithmetic matches the tensor-core path it replaces: inputs and weights rounded to the storage type (bf16), fp32
// accumulation, bias + activation in fp32, output rounded to the storage type.
#include "common.cuh"

namespace s3d {
namespace {

// Packed fp32 FMA (sm_100 FFMA2): two lanes per instruction -- this kernel is bound by FMA issue slots.
__device__ __forceinline__ void ffma2(float& d0, float& d1, float x, float w0, float w1) {
  unsigned long long rd, ra, rb, rc;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ra) : "f"(x));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(w0), "f"(w1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(d0), "f"(d1));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(rd));
}

// kSplit (S3D_DTYPE_BF16X2, 'bf16x3' precision): weights are [hi(cin_pad) | lo(cin_pad)] bf16 rows (summed back to fp32 here),
// the input is NOT rounded, and the output is written as [hi(CO) | lo(CO)] bf16 per pixel.
template <int CO, int CIN, typename TW, typename TOut, bool kU8, bool kSplit = false>
__global__ void __launch_bounds__(128)
conv_first_kernel(const void* __restrict__ img_, const float* __restrict__ disp, float disp_scale, float img_scale,
                  const TW* __restrict__ w, int cin_pad, const float* __restrict__ bias, TOut* __restrict__ out,
                  int B, int H, int W, int oH, int oW, int act, float act_param) {
  // weights as fp32 [tap][ci][co]
  __shared__ __align__(16) float ws[9 * CIN * CO];
  __shared__ __align__(16) float bs[CO];
  for (int i = threadIdx.x; i < 9 * CIN * CO; i += blockDim.x) {
    const int co = i % CO, ci = (i / CO) % CIN, t = i / (CO * CIN);
    if (kSplit) ws[i] = to_f32(w[((int64_t)t * CO + co) * 2 * cin_pad + ci]) + to_f32(w[((int64_t)t * CO + co) * 2 * cin_pad + cin_pad + ci]);
    else        ws[i] = to_f32(w[((int64_t)t * CO + co) * cin_pad + ci]);
  }
  for (int i = threadIdx.x; i < CO; i += blockDim.x) bs[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  // a thread computes TWO horizontally adjacent output pixels (x = 2j, 2j+1): they share every weight read and the
  // middle input column
  const int oW2 = (oW + 1) >> 1;
  const int64_t total = (int64_t)B * oH * oW2;
  const float slope = act == S3D_ACT_NONE ? 1.f : (act == S3D_ACT_LEAKY ? act_param : 0.f);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(idx % oW2), oy = (int)((idx / oW2) % oH), n = (int)(idx / ((int64_t)oW2 * oH));
    const int ox = 2 * j;
    float acc[2][CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) { acc[0][c] = bs[c]; acc[1][c] = bs[c]; }
#pragma unroll 1                                       // (kept rolled: the fully unrolled body made this file the slowest to compile)
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = 2 * oy + ky - 1;
      if (iy < 0 || iy >= H) continue;
      // input columns 2*ox-1 .. 2*ox+3: pixel 0 uses columns 0..2 of this window, pixel 1 columns 2..4
      float v[5][CIN];
#pragma unroll
      for (int cx = 0; cx < 5; ++cx) {
        const int ix = 2 * ox + cx - 1;
        const bool in = ix >= 0 && ix < W;
        if (kU8) {
          // byte -> float without I2F (a quarter-rate conversion, 135 of them per thread): 2^23 + b is the float with mantissa
          // bits b, and fma(2^23 + b, s, -2^23 s) = b * s rounded once -- bit-identical to (float)b * s
          const uint8_t* ip = reinterpret_cast<const uint8_t*>(img_) + (((int64_t)n * H + iy) * W + ix) * 3;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci)
            v[cx][ci] = in ? fmaf(__uint_as_float(0x4B000000u | (unsigned)ip[ci]), img_scale, -8388608.f * img_scale) : 0.f;
        } else {
          const float* ip = reinterpret_cast<const float*>(img_) + ((int64_t)n * 3 * H + iy) * W + ix;
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) v[cx][ci] = in ? __ldg(ip + ci * (int64_t)H * W) : 0.f;
        }
        if (CIN > 3) v[cx][3] = in ? __ldg(disp + ((int64_t)n * H + iy) * W + ix) * disp_scale : 0.f;
        if (!kSplit) {
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) v[cx][ci] = to_f32(from_f32<TW>(v[cx][ci]));   // rounding of the staged copy it replaces
        }
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float* wt = ws + (ky * 3 + kx) * CIN * CO;
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci) {
          const float x0 = v[kx][ci], x1 = v[kx + 2][ci];
          const float4* w4 = reinterpret_cast<const float4*>(wt + ci * CO);
#pragma unroll
          for (int c4 = 0; c4 < CO / 4; ++c4) {
            const float4 q = w4[c4];
            ffma2(acc[0][4 * c4], acc[0][4 * c4 + 1], x0, q.x, q.y);
            ffma2(acc[0][4 * c4 + 2], acc[0][4 * c4 + 3], x0, q.z, q.w);
            ffma2(acc[1][4 * c4], acc[1][4 * c4 + 1], x1, q.x, q.y);
            ffma2(acc[1][4 * c4 + 2], acc[1][4 * c4 + 3], x1, q.z, q.w);
          }
        }
      }
    }
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (ox + px >= oW) break;
      TOut* o = out + (((int64_t)n * oH + oy) * oW + ox + px) * (kSplit ? 2 * CO : CO);
      if (act <= S3D_ACT_LEAKY) {
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[px][c] = fmax_nan(acc[px][c], 0.f) + slope * fmin_nan(acc[px][c], 0.f);
      } else {
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[px][c] = apply_act(acc[px][c], act, act_param);
      }
      if (kSplit) {
#pragma unroll
        for (int c8 = 0; c8 < CO / 8; ++c8) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(acc[px][8 * c8 + 2 * i], acc[px][8 * c8 + 2 * i + 1]);
            const float2 hf = __bfloat1622float2(h);
            const __nv_bfloat162 l = __floats2bfloat162_rn(acc[px][8 * c8 + 2 * i] - hf.x, acc[px][8 * c8 + 2 * i + 1] - hf.y);
            hi[i] = *reinterpret_cast<const uint32_t*>(&h);  lo[i] = *reinterpret_cast<const uint32_t*>(&l);
          }
          reinterpret_cast<uint4*>(o)[c8] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          reinterpret_cast<uint4*>(o + CO)[c8] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      } else if (sizeof(TOut) == 2) {
#pragma unroll
        for (int c8 = 0; c8 < CO / 8; ++c8) {
          uint4 pk;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
          for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(acc[px][8 * c8 + 2 * i], acc[px][8 * c8 + 2 * i + 1]);
          reinterpret_cast<uint4*>(o)[c8] = pk;
        }
      } else {
#pragma unroll
        for (int c4 = 0; c4 < CO / 4; ++c4)
          reinterpret_cast<float4*>(o)[c4] = make_float4(acc[px][4 * c4], acc[px][4 * c4 + 1], acc[px][4 * c4 + 2], acc[px][4 * c4 + 3]);
      }
    }
  }
}

template <int CO, typename TW, typename TOut, bool kSplit = false>
int launch_first(const void* img, int img_u8, const float* disp, float disp_scale, float img_scale, const void* w, int cin_pad,
                 const float* bias, void* out, int B, int H, int W, int oH, int oW, int cin, int act, float act_param,
                 cudaStream_t st) {
  const int64_t total = (int64_t)B * oH * ((oW + 1) / 2);
  int64_t blocks = ceil_div64(total, 128);
  const int64_t cap = (int64_t)num_sms() * 16;
  if (blocks > cap) blocks = cap;
#define S3D_FIRST_LAUNCH(CIN, U8)                                                                                          \
  conv_first_kernel<CO, CIN, TW, TOut, U8, kSplit><<<(int)blocks, 128, 0, st>>>(img, disp, disp_scale, img_scale, static_cast<const TW*>(w), \
      cin_pad, bias, static_cast<TOut*>(out), B, H, W, oH, oW, act, act_param)
  if (cin == 3) { if (img_u8) S3D_FIRST_LAUNCH(3, true); else S3D_FIRST_LAUNCH(3, false); }
  else          { if (img_u8) S3D_FIRST_LAUNCH(4, true); else S3D_FIRST_LAUNCH(4, false); }
#undef S3D_FIRST_LAUNCH
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

}  // namespace
}  // namespace s3d

namespace s3d {
bool conv_first_tc_eligible(int cout_pad, int dtype, int act);
int conv_first_tc_launch(const void* img, int img_u8, const float* disp, float disp_scale, float img_scale, const void* w,
                         int cin_pad, const float* bias, void* out, int B, int H, int W, int oH, int oW, int cin, int act,
                         float act_param, cudaStream_t st);
}

extern "C" int s3d_conv_first(const void* img, int img_u8, const float* disp, float disp_scale, const void* w, const float* bias,
                              void* out, int B, int H, int W, int cin, int cin_pad, int cout_pad, int dtype, int act,
                              float act_param, void* stream) {
  using namespace s3d;
  if (!img || !w || !out) { set_error("conv_first: null argument"); return S3D_ERR_INVALID; }
  S3D_CHECK_ARG(B > 0 && H > 0 && W > 0, "conv_first: bad shape");
  S3D_CHECK_ARG(cin == 3 || cin == 4, "conv_first: cin must be 3 (image) or 4 (image + disparity)");
  S3D_CHECK_ARG(cin <= cin_pad, "conv_first: cin_pad");
  S3D_CHECK_ARG(cin == 3 || disp != nullptr, "conv_first: cin = 4 needs a disparity map");
  S3D_CHECK_ARG(cout_pad == 16 || cout_pad == 32 || cout_pad == 64, "conv_first: cout_pad must be 16, 32 or 64");
  S3D_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0, "conv_first: out must be 16-byte aligned");
  const int oH = (H - 1) / 2 + 1, oW = (W - 1) / 2 + 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float img_scale = 1.0f / 255.0f;
  // bf16, 32 output channels: im2col rows built by the threads + tcgen05 MMAs (conv_first_tc.cu)
  if (conv_first_tc_eligible(cout_pad, dtype, act))
    return conv_first_tc_launch(img, img_u8, disp, disp_scale, img_scale, w, cin_pad, bias, out, B, H, W, oH, oW, cin, act, act_param, st);
#define S3D_FIRST(CO)                                                                                                    \
  (dtype == S3D_DTYPE_BF16X2                                                                                             \
       ? launch_first<CO, __nv_bfloat16, __nv_bfloat16, true>(img, img_u8, disp, disp_scale, img_scale, w, cin_pad, bias, out, B, \
                                                              H, W, oH, oW, cin, act, act_param, st)                     \
   : dtype == S3D_DTYPE_BF16                                                                                               \
       ? launch_first<CO, __nv_bfloat16, __nv_bfloat16>(img, img_u8, disp, disp_scale, img_scale, w, cin_pad, bias, out, B, H, W, \
                                                        oH, oW, cin, act, act_param, st)                                 \
       : launch_first<CO, float, float>(img, img_u8, disp, disp_scale, img_scale, w, cin_pad, bias, out, B, H, W, oH, oW, cin,    \
                                        act, act_param, st))
  if (dtype != S3D_DTYPE_BF16 && dtype != S3D_DTYPE_F32 && dtype != S3D_DTYPE_BF16X2) { set_error("conv_first: bad dtype"); return S3D_ERR_INVALID; }
  if (cout_pad == 16) return S3D_FIRST(16);
  if (cout_pad == 32) return S3D_FIRST(32);
  return S3D_FIRST(64);
#undef S3D_FIRST
}
