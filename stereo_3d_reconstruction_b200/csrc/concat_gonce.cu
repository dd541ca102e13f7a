// Rows V + A (first aggregation layer), SHEARED form: the streaming pass.
//
// In the coordinate u = x - d (left-referenced volume; u = x + d for the right-referenced one) the target half of the
// concat cost volume does not depend on the disparity plane, so the 3x3x3 layer over it collapses to 2-D maps
// (stereo_3d_reconstruction_b200/layers.py, PackedConv.gonce_convs -- computed by the generic tensor-core engine):
//     out[a, d, y, x] = relu( bias[a] + Psum[y, x] + G[y, u]                        every plane
//                             + [d = 0]   (Pm[y, x] + Hm[y, u])                     no plane below  (dz = -1 slices, negated)
//                             + [d = D-1] (Pp[y, x] + Hp[y, u])                     no plane above  (dz = +1)
//                             + [x = xe]  (Ge[y, j] + [d = 0] Gem[y, j] + [d = D-1] Gep[y, j]) )      taps past the image edge
// with xe = w-1, j = D-1-d (left reference) or xe = 0, j = d (right reference).  This kernel only adds, clamps, rounds and
// WRITES the bf16 volume: 2.1 GB at batch 64, read by nothing but the next layer.  It is bound by that write; the maps are
// read once from HBM and then from L1 / L2 (a map row slides by one pixel per plane under a CTA's 32 pixels).
#include "common.cuh"

namespace s3d {
namespace {

constexpr int kMaps = 6 * 64;      // channels of a map pixel: [Psum | Pm | Pp | G | Hm | Hp]
constexpr int kEdge = 4 * 64;      //                 edge map: [Ge | Gem | Gep | unused]

__device__ __forceinline__ void add8(float (&v)[8], const float* __restrict__ p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] += a.x;  v[1] += a.y;  v[2] += a.z;  v[3] += a.w;  v[4] += b.x;  v[5] += b.y;  v[6] += b.z;  v[7] += b.w;
}

// grid (ceil(w / 32), h, 2B); 256 threads = 32 pixels of one image row x 8 channel octets; each thread marches over the D planes.
// kSplit ('bf16x3'): the output is the bf16 pair [hi(64) | lo(64)] per pixel, hi = bf16(v), lo = bf16(v - hi).
template <bool kSplit>
__global__ void __launch_bounds__(256)
gonce_assemble_kernel(const float* __restrict__ maps_l, const float* __restrict__ maps_r, const float* __restrict__ edge_l,
                      const float* __restrict__ edge_r, const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                      int B, int D, int h, int w, int mw) {
  const int x = blockIdx.x * 32 + (threadIdx.x >> 3), co = (threadIdx.x & 7) * 8;
  const int y = blockIdx.y, n = blockIdx.z;
  if (x >= w) return;
  const bool left_ref = n < B;
  const int b = left_ref ? n : n - B;
  const float* pm = (left_ref ? maps_l : maps_r) + (((int64_t)b * h + y) * mw + (x + 2)) * kMaps + co;      // reference maps at x
  const float* gm = (left_ref ? maps_r : maps_l) + ((int64_t)b * h + y) * mw * kMaps + 3 * 64 + co;          // target maps, column u + 2
  const bool edge = x == (left_ref ? w - 1 : 0);
  const float* em = (left_ref ? edge_r : edge_l) + ((int64_t)b * h + y) * D * kEdge + co;
  float base[8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + co)), b1 = __ldg(reinterpret_cast<const float4*>(bias + co) + 1);
    base[0] = b0.x;  base[1] = b0.y;  base[2] = b0.z;  base[3] = b0.w;  base[4] = b1.x;  base[5] = b1.y;  base[6] = b1.z;  base[7] = b1.w;
    add8(base, pm);
  }
  constexpr int kCo = kSplit ? 128 : 64;                     // bf16 elements of an output pixel
  __nv_bfloat16* op = out + ((((int64_t)n * D) * h + y) * w + x) * kCo + co;
  const int64_t plane = (int64_t)h * w * kCo;
#pragma unroll 4
  for (int d = 0; d < D; ++d) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = base[i];
    const int col = (left_ref ? x - d : x + d) + 2;          // the target maps vanish outside [0, mw): every tap reads the zero margin
    const bool in = col >= 0 && col < mw;
    if (in) add8(v, gm + (int64_t)col * kMaps);
    if (d == 0)     { add8(v, pm + 64);   if (in) add8(v, gm + (int64_t)col * kMaps + 64); }
    if (d == D - 1) { add8(v, pm + 128);  if (in) add8(v, gm + (int64_t)col * kMaps + 128); }
    if (edge) {
      const float* e = em + (int64_t)(left_ref ? D - 1 - d : d) * kEdge;
      add8(v, e);
      if (d == 0) add8(v, e + 64);
      if (d == D - 1) add8(v, e + 128);
    }
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float a0 = fmax_nan(v[2 * i], 0.f), a1 = fmax_nan(v[2 * i + 1], 0.f);
      __nv_bfloat162 t = __floats2bfloat162_rn(a0, a1);
      oh[i] = *reinterpret_cast<uint32_t*>(&t);
      if constexpr (kSplit) {
        const float2 hf = __bfloat1622float2(t);
        __nv_bfloat162 l = __floats2bfloat162_rn(a0 - hf.x, a1 - hf.y);
        ol[i] = *reinterpret_cast<uint32_t*>(&l);
      }
    }
    const uint4 o = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    __stcs(reinterpret_cast<uint4*>(op + d * plane), o);     // streaming store: the volume is read next by another kernel, not by this one
    if constexpr (kSplit) __stcs(reinterpret_cast<uint4*>(op + d * plane + 64), make_uint4(ol[0], ol[1], ol[2], ol[3]));
  }
}

}  // namespace
}  // namespace s3d

extern "C" int s3d_concat_gonce_assemble(const float* maps_l, const float* maps_r, const float* edge_l, const float* edge_r,
                                         const float* bias, void* out, int B, int D, int h, int w, int map_w, int out_dtype,
                                         void* stream) {
  using namespace s3d;
  if (!maps_l || !maps_r || !edge_l || !edge_r || !bias || !out) { set_error("concat_gonce_assemble: null argument"); return S3D_ERR_INVALID; }
  S3D_CHECK_ARG(B > 0 && D >= 2 && h > 0 && w > 0 && map_w == w + 4, "concat_gonce_assemble: B=%d D=%d h=%d w=%d map_w=%d (needs D >= 2, map_w = w + 4)", B, D, h, w, map_w);
  S3D_CHECK_ARG(((reinterpret_cast<uintptr_t>(maps_l) | reinterpret_cast<uintptr_t>(maps_r) | reinterpret_cast<uintptr_t>(edge_l) |
                  reinterpret_cast<uintptr_t>(edge_r) | reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
                "concat_gonce_assemble: pointers must be 16-byte aligned");
  S3D_CHECK_ARG(h <= 65535 && 2 * B <= 65535, "concat_gonce_assemble: grid too large");
  S3D_CHECK_ARG(out_dtype == S3D_DTYPE_BF16 || out_dtype == S3D_DTYPE_BF16X2, "concat_gonce_assemble: out_dtype %d", out_dtype);
  dim3 grid(ceil_div(w, 32), h, 2 * B);
  if (out_dtype == S3D_DTYPE_BF16X2)
    gonce_assemble_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(maps_l, maps_r, edge_l, edge_r, bias,
                                                                                     static_cast<__nv_bfloat16*>(out), B, D, h, w, map_w);
  else
    gonce_assemble_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(maps_l, maps_r, edge_l, edge_r, bias,
                                                                                      static_cast<__nv_bfloat16*>(out), B, D, h, w, map_w);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}
