// Shared epilogue of the tcgen05 conv kernels: 16 fp32 accumulator columns of one output row ->
// (+bias) (+residual) -> activation -> bf16 / fp32 store, vectorised when the row slice is 16 B
// aligned and fully inside cout_store.  Also the host-side TMA descriptor helper.
#pragma once
#include <cuda.h>
#include <string.h>
#include <unordered_map>
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
  int64_t os_lo;      // split (BF16X2) output: element offset from a channel's hi part to its lo part
};

// v = hi + lo with hi = bf16(v), lo = bf16(v - hi): the operand split of the 'bf16x3' precision (s3d.h, S3D_DTYPE_BF16X2).
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// Two values -> packed (hi0, hi1) and (lo0, lo1).
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

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

// The same for a split (BF16X2) channels-last output: hi parts at [off + c], lo parts at [off + os_lo + c]; a residual is
// read the same way (hi + lo).  No planar layout, no fused projection.
__device__ __forceinline__ void epilogue_store16_split(const EpiParams& e, int64_t off, int cg, const uint32_t (&v)[16]) {
  if (cg >= e.cout_store) return;
  float f[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
  if (e.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(e.bias + cg);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 b = __ldg(b4 + i);
      f[4 * i] += b.x;  f[4 * i + 1] += b.y;  f[4 * i + 2] += b.z;  f[4 * i + 3] += b.w;
    }
  }
  const bool full = cg + 16 <= e.cout_store;
  const __nv_bfloat16* rs = reinterpret_cast<const __nv_bfloat16*>(e.residual);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out);
  const int64_t bh = off + cg, bl = bh + e.os_lo;
  const bool vec = full && (((uintptr_t)(o + bh) | (uintptr_t)(o + bl)) & 15) == 0 &&
                   (!e.residual || (((uintptr_t)(rs + bh) | (uintptr_t)(rs + bl)) & 15) == 0);
  if (e.residual) {
    if (vec) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint4 rh = __ldg(reinterpret_cast<const uint4*>(rs + bh) + h), rl = __ldg(reinterpret_cast<const uint4*>(rs + bl) + h);
        const __nv_bfloat162* ph = reinterpret_cast<const __nv_bfloat162*>(&rh);
        const __nv_bfloat162* pl = reinterpret_cast<const __nv_bfloat162*>(&rl);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 gh = __bfloat1622float2(ph[i]), gl = __bfloat1622float2(pl[i]);
          f[8 * h + 2 * i] += gh.x + gl.x;  f[8 * h + 2 * i + 1] += gh.y + gl.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (cg + i < e.cout_store) f[i] += __bfloat162float(rs[bh + i]) + __bfloat162float(rs[bl + i]);
    }
  }
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
  if (vec) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) split_bf16x2(f[2 * i], f[2 * i + 1], hi[i], lo[i]);
    reinterpret_cast<uint4*>(o + bh)[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    reinterpret_cast<uint4*>(o + bh)[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    reinterpret_cast<uint4*>(o + bl)[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    reinterpret_cast<uint4*>(o + bl)[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (cg + i < e.cout_store) {
        __nv_bfloat16 h, l;
        split_bf16(f[i], h, l);
        o[bh + i] = h;  o[bl + i] = l;
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

// cuTensorMapEncodeTiled costs ~1-2 us of host time and a layer's operands (pointer, shape, box) repeat every forward:
// encoded maps are cached per thread, keyed on every argument of the encode call (+ the device).
struct MapKey {
  const void* base;
  int32_t dev, dtype, rank, sw, l2;
  uint64_t dims[5], strides[4];
  uint32_t box[5], estr[5];
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ull; }
    return (size_t)h;
  }
};
static_assert(sizeof(MapKey) % 8 == 0, "MapKey is hashed in 8-byte words");

inline int encode_tiled_cached(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                               const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapSwizzle sw,
                               CUtensorMapL2promotion l2, const char* what) {
  static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey k;
  memset(&k, 0, sizeof(k));
  k.base = base;  k.dtype = (int)dt;  k.rank = rank;  k.sw = (int)sw;  k.l2 = (int)l2;
  cudaGetDevice(&k.dev);
  for (int i = 0; i < rank; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; k.estr[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) k.strides[i] = strides[i];
  auto it = cache.find(k);
  if (it != cache.end()) { *m = it->second; return S3D_OK; }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return S3D_ERR_CUDA; }
  CUresult r = enc(m, dt, rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, l2,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r); return S3D_ERR_CUDA; }
  if (cache.size() >= 4096) cache.clear();
  cache.emplace(k, *m);
  return S3D_OK;
}

// Channels-last activation [N,D,H,W,C] as a 5-D map (C,W,H,D,N) with the given box / element strides.
inline int encode_act_map(CUtensorMap* m, const void* base, int esz, bool f32, int C, int W, int H, int D, int N,
                          const cuuint32_t box[5], const cuuint32_t estr[5], CUtensorMapSwizzle sw) {
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
  cuuint64_t strides[4];
  strides[0] = (cuuint64_t)C * esz;
  strides[1] = strides[0] * W;
  strides[2] = strides[1] * H;
  strides[3] = strides[2] * D;
  return encode_tiled_cached(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides,
                             box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "activations");
}

// Packed weights [rows][Cout][Cin] as a 3-D map (Cin, Cout, rows), box (kc, bn, box_taps).
inline int encode_weight_map(CUtensorMap* m, const void* base, int esz, bool f32, int Cin, int Cout, int rows, int kc,
                             int bn, CUtensorMapSwizzle sw, int box_taps = 1) {
  cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)Cin * esz, (cuuint64_t)Cin * esz * Cout};
  cuuint32_t box[3] = {(cuuint32_t)kc, (cuuint32_t)bn, (cuuint32_t)box_taps};
  cuuint32_t estr[3] = {1, 1, 1};
  return encode_tiled_cached(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides,
                             box, estr, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "weights");
}

}  // namespace s3d
