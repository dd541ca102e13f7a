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
};

// v: raw accumulator bits of columns [cg, cg+16); off: element offset of the row's channel 0.
__device__ __forceinline__ void epilogue_store16(const EpiParams& e, int64_t off, int cg, const uint32_t (&v)[16]) {
  if (cg >= e.cout_store) return;
  float f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    f[i] = __uint_as_float(v[i]);
    if (e.bias) f[i] += __ldg(e.bias + cg + i);
  }
  const bool full = cg + 16 <= e.cout_store;
  if (e.out_bf16) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + off + cg;
    const __nv_bfloat16* rs = e.residual ? reinterpret_cast<const __nv_bfloat16*>(e.residual) + off + cg : nullptr;
    const bool vec = full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0) &&
                     (!rs || (reinterpret_cast<uintptr_t>(rs) & 15) == 0);
    if (vec) {
      if (rs) {
        uint4 r0 = __ldg(reinterpret_cast<const uint4*>(rs));
        uint4 r1 = __ldg(reinterpret_cast<const uint4*>(rs) + 1);
        const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
        const __nv_bfloat162* h1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float2 g0 = __bfloat1622float2(h0[i]), g1 = __bfloat1622float2(h1[i]);
          f[2 * i] += g0.x;  f[2 * i + 1] += g0.y;
          f[8 + 2 * i] += g1.x;  f[8 + 2 * i + 1] += g1.y;
        }
      }
      uint4 w0, w1;
      __nv_bfloat162* p0 = reinterpret_cast<__nv_bfloat162*>(&w0);
      __nv_bfloat162* p1 = reinterpret_cast<__nv_bfloat162*>(&w1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        p0[i] = __floats2bfloat162_rn(apply_act(f[2 * i], e.act, e.act_param), apply_act(f[2 * i + 1], e.act, e.act_param));
        p1[i] = __floats2bfloat162_rn(apply_act(f[8 + 2 * i], e.act, e.act_param),
                                      apply_act(f[8 + 2 * i + 1], e.act, e.act_param));
      }
      reinterpret_cast<uint4*>(o)[0] = w0;
      reinterpret_cast<uint4*>(o)[1] = w1;
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (cg + i < e.cout_store) {
          float s = f[i];
          if (rs) s += __bfloat162float(rs[i]);
          o[i] = __float2bfloat16_rn(apply_act(s, e.act, e.act_param));
        }
      }
    }
  } else {
    float* o = reinterpret_cast<float*>(e.out) + off + cg;
    const float* rs = e.residual ? reinterpret_cast<const float*>(e.residual) + off + cg : nullptr;
    const bool vec = full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0) &&
                     (!rs || (reinterpret_cast<uintptr_t>(rs) & 15) == 0);
    if (vec) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 s = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
        if (rs) {
          float4 g = __ldg(reinterpret_cast<const float4*>(rs) + i);
          s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
        }
        s.x = apply_act(s.x, e.act, e.act_param);  s.y = apply_act(s.y, e.act, e.act_param);
        s.z = apply_act(s.z, e.act, e.act_param);  s.w = apply_act(s.w, e.act, e.act_param);
        reinterpret_cast<float4*>(o)[i] = s;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (cg + i < e.cout_store) {
          float s = f[i];
          if (rs) s += rs[i];
          o[i] = apply_act(s, e.act, e.act_param);
        }
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
