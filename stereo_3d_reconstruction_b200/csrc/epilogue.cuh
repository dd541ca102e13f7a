// Shared epilogue of the tcgen05 conv kernels: 16 fp32 accumulator columns of one output row ->
// (+bias) (+residual) -> activation -> bf16 / fp32 store, vectorised when the row slice is 16 B
// aligned and fully inside cout_store.  Also the host-side TMA descriptor helper.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace s3d {

struct EpiParams {
  const float* bias;
  const void* residual;
  void* out;
  int cout_store;
  int out_bf16;
  int act;
  float act_param;
  int64_t osC;        // channel stride of the output (1 = channels-last)
  const float* proj_w;  int proj_channel, proj_act;     // optional fused 1x1 projection (see s3d.h)
};

// v: raw accumulator bits of columns [cg, cg+16); off: element offset of the row's channel 0.
// Code size matters here (the epilogue is inlined into a kernel whose hot loop must stay in the
// instruction cache): none / ReLU / LeakyReLU share one branch-free formula, sigmoid / tanh (two
// layers in the whole network) go through a rolled loop.
__device__ __forceinline__ void epilogue_store16(const EpiParams& e, int64_t off, int cg, const uint32_t (&v)[16]) {
  if (cg >= e.cout_store) return;
  float f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
  if (e.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + cg);       // bias is padded to Cout, 64-B aligned groups
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b = __ldg(b4 + i);
      f[4 * i] += b.x;  f[4 * i + 1] += b.y;  f[4 * i + 2] += b.z;  f[4 * i + 3] += b.w;
    }
  }
  const bool full = cg + 16 <= e.cout_store;
  const bool planar = e.osC != 1;
  const __nv_bfloat16* rs16 = reinterpret_cast<const __nv_bfloat16*>(e.residual);
  const float* rs32 = reinterpret_cast<const float*>(e.residual);
  __nv_bfloat16* o16 = reinterpret_cast<__nv_bfloat16*>(e.out);
  float* o32 = reinterpret_cast<float*>(e.out);
  const int64_t base = off + (planar ? (int64_t)cg * e.osC : (int64_t)cg);
  const bool vec = !planar && full && (((e.out_bf16 ? (uintptr_t)(o16 + base) : (uintptr_t)(o32 + base)) & 15) == 0) &&
                   (!e.residual || ((e.out_bf16 ? (uintptr_t)(rs16 + base) : (uintptr_t)(rs32 + base)) & 15) == 0);
  // ---- residual
  if (e.residual) {
    if (vec && e.out_bf16) {
      const uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rs16 + base));
      const uint4 r1 = __ldg(reinterpret_cast<const uint4*>(rs16 + base) + 1);
      const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
      const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 g0 = __bfloat1622float2(h0[i]), g1 = __bfloat1622float2(h1[i]);
        f[2 * i] += g0.x;  f[2 * i + 1] += g0.y;  f[8 + 2 * i] += g1.x;  f[8 + 2 * i + 1] += g1.y;
      }
    } else if (vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(rs32 + base) + i);
        f[4 * i] += g.x;  f[4 * i + 1] += g.y;  f[4 * i + 2] += g.z;  f[4 * i + 3] += g.w;
      }
    } else {
      const int64_t cs = planar ? e.osC : 1;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (cg + i < e.cout_store) f[i] += e.out_bf16 ? __bfloat162float(rs16[base + i * cs]) : rs32[base + i * cs];
    }
  }
  // ---- activation
  if (e.act == S3D_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = fmax_nan(f[i], 0.f);
  } else if (e.act == S3D_ACT_LEAKY) {
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = fmax_nan(f[i], 0.f) + e.act_param * fmin_nan(f[i], 0.f);
  } else if (e.act != S3D_ACT_NONE) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (cg + i < e.cout_store) f[i] = apply_act_slow(f[i], e.act, e.act_param);
  }
  // ---- optional fused 1x1 projection into one channel of this (single) group
  if (e.proj_w) {
    float pr = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) pr = fmaf(f[i], __ldg(e.proj_w + i), pr);
    pr = apply_act_slow(pr, e.proj_act, 1.f);
#pragma unroll
    for (int i = 0; i < 16; ++i) if (i == e.proj_channel) f[i] = pr;
  }
  // ---- store
  if (vec && e.out_bf16) {
    uint4 w0, w1;
    __nv_bfloat162* p0 = reinterpret_cast<__nv_bfloat162*>(&w0);
    __nv_bfloat162* p1 = reinterpret_cast<__nv_bfloat162*>(&w1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      p0[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      p1[i] = __floats2bfloat162_rn(f[8 + 2 * i], f[8 + 2 * i + 1]);
    }
    reinterpret_cast<uint4*>(o16 + base)[0] = w0;
    reinterpret_cast<uint4*>(o16 + base)[1] = w1;
  } else if (vec) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      reinterpret_cast<float4*>(o32 + base)[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
  } else {
    // channels-last tail / unaligned slice, or planar output (consecutive rows = consecutive x, so each
    // channel is one coalesced warp store)
    const int64_t cs = planar ? e.osC : 1;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (cg + i < e.cout_store) {
        if (e.out_bf16) o16[base + i * cs] = __float2bfloat16_rn(f[i]);
        else            o32[base + i * cs] = f[i];
      }
    }
  }
}

// ---- host: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Channels-last activation [N,D,H,W,C] as a 5-D map (C,W,H,D,N) with the given box / element strides.
inline int encode_act_map(CUtensorMap* m, const void* base, int esz, bool f32, int C, int W, int H, int D, int N,
                          const cuuint32_t box[5], const cuuint32_t estr[5], CUtensorMapSwizzle sw) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return S3D_ERR_CUDA; }
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4];
  strides[0] = (cuuint64_t)C * esz;
  strides[1] = strides[0] * W;
  strides[2] = strides[1] * H;
  strides[3] = strides[2] * D;
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations) failed: %d", (int)r); return S3D_ERR_CUDA; }
  return S3D_OK;
}

// Packed weights [rows][Cout][Cin] as a 3-D map (Cin, Cout, rows), box (kc, bn, box_taps).
inline int encode_weight_map(CUtensorMap* m, const void* base, int esz, bool f32, int Cin, int Cout, int rows, int kc,
                             int bn, CUtensorMapSwizzle sw, int box_taps = 1) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return S3D_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * esz, (cuuint64_t)Cin * esz * Cout};
  cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)bn, (cuuint32_t)box_taps};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights) failed: %d", (int)r); return S3D_ERR_CUDA; }
  return S3D_OK;
}

}  // namespace s3d
