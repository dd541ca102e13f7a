// Rows V and S: disparity cost volume and soft-argmin regression.  All HBM-bound kernels.
//
//   cost_volume_concat   one CTA per (volume, image row): the reference row and the target row
//                        are staged once in shared memory, then D planes of 2C channels are
//                        streamed out with 16 B coalesced stores.  Bytes: read 2*C*w*e, write
//                        D*2C*w*e per row -- write-bound.
//   soft_argmin          lanes along x (coalesced planes), online softmax over d in registers.
//   corr_soft_argmin     fused correlation + soft-argmax: the [D,h,w] cost never reaches HBM.
//                        Rows staged (transposed, zero padded by D) in shared memory; each thread
//                        owns a 4(x) x 8(d) register tile, the disparity axis is finished with
//                        warp-shuffle (max, sum, weighted-sum) reductions.
#include "common.cuh"

namespace s3d {
namespace {

// ------------------------------------------------------------------------------------------
// concat volume
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
concat_volume_kernel(const uint4* __restrict__ feat, uint4* __restrict__ vol, int B, int h, int w,
                     int vpp /* 16-B vectors per pixel of one feature map */, int D) {
  extern __shared__ uint4 srow[];          // [2][w*vpp]: ref row, tgt row
  const int n = blockIdx.x / h, y = blockIdx.x % h;
  const bool left_ref = n < B;
  const int tgt_n = left_ref ? n + B : n - B;
  const int rowv = w * vpp;
  const uint4* ref_g = feat + ((int64_t)n * h + y) * rowv;
  const uint4* tgt_g = feat + ((int64_t)tgt_n * h + y) * rowv;
  for (int i = threadIdx.x; i < rowv; i += blockDim.x) {
    srow[i] = __ldg(ref_g + i);
    srow[rowv + i] = __ldg(tgt_g + i);
  }
  __syncthreads();
  const int outv = 2 * rowv;               // vectors per (d, row) output line
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int d = 0; d < D; ++d) {
    uint4* o = vol + (((int64_t)n * D + d) * h + y) * outv;
    for (int i = threadIdx.x; i < outv; i += blockDim.x) {
      const int x = i / (2 * vpp), j = i % (2 * vpp);
      uint4 v;
      if (j < vpp) {
        v = srow[x * vpp + j];
      } else {
        const int xs = left_ref ? x - d : x + d;
        v = (xs >= 0 && xs < w) ? srow[rowv + xs * vpp + (j - vpp)] : zero;
      }
      __stcs(o + i, v);                    // streaming store: the volume is consumed much later
    }
  }
}

// Split (BF16X2) features [.., hi(C) | lo(C)] -> split volume [.., hi(ref C, tgt C) | lo(ref C, tgt C)]: the hi halves and the
// lo halves each form the plain concat volume, so a 2C-channel split layer reads it like any other split tensor.
__global__ void __launch_bounds__(256)
concat_volume_split_kernel(const uint4* __restrict__ feat, uint4* __restrict__ vol, int B, int h, int w,
                           int vpp /* 16-B vectors per pixel of one HALF of a feature map */, int D) {
  extern __shared__ uint4 srow[];          // [2][w * 2 vpp]: ref row, tgt row (physical pixels)
  const int n = blockIdx.x / h, y = blockIdx.x % h;
  const bool left_ref = n < B;
  const int tgt_n = left_ref ? n + B : n - B;
  const int rowv = w * 2 * vpp;
  const uint4* ref_g = feat + ((int64_t)n * h + y) * rowv;
  const uint4* tgt_g = feat + ((int64_t)tgt_n * h + y) * rowv;
  for (int i = threadIdx.x; i < rowv; i += blockDim.x) {
    srow[i] = __ldg(ref_g + i);
    srow[rowv + i] = __ldg(tgt_g + i);
  }
  __syncthreads();
  const int outv = 2 * rowv;               // vectors per (d, row) output line: 4 vpp per pixel
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int d = 0; d < D; ++d) {
    uint4* o = vol + (((int64_t)n * D + d) * h + y) * outv;
    for (int i = threadIdx.x; i < outv; i += blockDim.x) {
      const int x = i / (4 * vpp), j = i % (4 * vpp);
      const int part = j / vpp, jj = j % vpp;          // part: 0 ref hi, 1 tgt hi, 2 ref lo, 3 tgt lo
      const int half = part >> 1;
      uint4 v;
      if ((part & 1) == 0) {
        v = srow[x * 2 * vpp + half * vpp + jj];
      } else {
        const int xs = left_ref ? x - d : x + d;
        v = (xs >= 0 && xs < w) ? srow[rowv + xs * 2 * vpp + half * vpp + jj] : zero;
      }
      __stcs(o + i, v);
    }
  }
}

// ------------------------------------------------------------------------------------------
// standalone soft-argmin over a [N,D,h,w] fp32 cost
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
soft_argmin_kernel(const float* __restrict__ cost, float* __restrict__ disp, int64_t npix_total, int64_t plane,
                   int D, float sign) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= npix_total) return;
  const int64_t n = i / plane, pix = i % plane;
  const float* c = cost + n * D * plane + pix;
  float m = -INFINITY, s = 0.f, t = 0.f;
  int d = 0;
  for (; d + 8 <= D; d += 8) {                 // 8 independent plane loads in flight per thread
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = sign * __ldcs(c + (int64_t)(d + k) * plane);
    float mm = v[0];
#pragma unroll
    for (int k = 1; k < 8; ++k) mm = fmaxf(mm, v[k]);
    if (mm > m) { const float sc = expf(m - mm); s *= sc; t *= sc; m = mm; }
#pragma unroll
    for (int k = 0; k < 8; ++k) { const float e = expf(v[k] - m); s += e; t += e * (float)(d + k); }
  }
  for (; d < D; ++d) {
    const float v = sign * __ldcs(c + (int64_t)d * plane);
    if (v > m) { const float sc = expf(m - v); s *= sc; t *= sc; m = v; }
    const float e = expf(v - m);  s += e;  t += e * (float)d;
  }
  disp[i] = t / s;
}

// Four pixels per thread (one 16-byte load per plane, 8 planes = 128 bytes in flight per thread): the one-pixel kernel
// above reaches 0.45-0.67 of the HBM peak on this 20-60 us stream; wider loads keep more bytes in flight per warp and
// quarter the instruction count.  Needs plane % 4 == 0 (a vector never straddles two volumes) and 16-byte aligned cost.
__global__ void __launch_bounds__(256)
soft_argmin_v4_kernel(const float4* __restrict__ cost, float4* __restrict__ disp, int64_t nvec_total, int64_t plane4,
                      int D, float sign) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= nvec_total) return;
  const int64_t n = i / plane4, pix = i % plane4;
  const float4* c = cost + n * D * plane4 + pix;
  const float sl = sign * 1.4426950408889634f;           // softmax in the base-2 domain: one MUFU.EX2 per value
  float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, s[4] = {0.f, 0.f, 0.f, 0.f}, t[4] = {0.f, 0.f, 0.f, 0.f};
  int d = 0;
  for (; d + 8 <= D; d += 8) {
    float4 q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) q[k] = __ldcs(c + (int64_t)(d + k) * plane4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = sl * (j == 0 ? q[k].x : j == 1 ? q[k].y : j == 2 ? q[k].z : q[k].w);
      float mm = v[0];
#pragma unroll
      for (int k = 1; k < 8; ++k) mm = fmaxf(mm, v[k]);
      if (mm > m[j]) { const float sc = exp2f(m[j] - mm); s[j] *= sc; t[j] *= sc; m[j] = mm; }
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float e = exp2f(v[k] - m[j]); s[j] += e; t[j] += e * (float)(d + k); }
    }
  }
  for (; d < D; ++d) {
    const float4 q = __ldcs(c + (int64_t)d * plane4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = sl * (j == 0 ? q.x : j == 1 ? q.y : j == 2 ? q.z : q.w);
      if (v > m[j]) { const float sc = exp2f(m[j] - v); s[j] *= sc; t[j] *= sc; m[j] = v; }
      const float e = exp2f(v - m[j]);  s[j] += e;  t[j] += e * (float)d;
    }
  }
  disp[i] = make_float4(t[0] / s[0], t[1] / s[1], t[2] / s[2], t[3] / s[3]);
}

// ------------------------------------------------------------------------------------------
// tap-plane gather + soft-argmin: the Cout = 1 classifier conv of the aggregation stack.
// A 3x3x3 conv with ONE output channel wastes a 128-row tensor-core tile per tap, so it is computed
// as a pointwise GEMM P[pix][t] = W_t . x[pix] (27 taps = 27 output "channels", one pass over the
// activations, tensor cores) followed by this gather: cost[z,y,x] = sum_t P[z+kz-1, y+ky-1, x+kx-1][t],
// fused with the soft-argmin over z so the cost volume itself never reaches HBM.
// One thread per (n, y, x); marching over the input plane z' it keeps the three partial sums of the
// output planes z'-1, z', z'+1 in registers and retires plane z'-1 into an online softmax.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
tap_gather_soft_argmin_kernel(const float* __restrict__ P, float* __restrict__ disp, float* __restrict__ cost_out,
                              int N, int D, int h, int w, int ts, float sign) {
  // taps are LINE-PLANAR: P[n][z][y][t][x]: for a fixed tap the lanes of a warp (consecutive x) read
  // consecutive floats (one coalesced wavefront per load, neighbours reuse L1), while the producer
  // writes each image line as one contiguous ts*w*4-byte block (DRAM-friendly).
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y, n = blockIdx.z;
  if (x >= w) return;
  const int64_t plane = (int64_t)h * w;
  const float* Pn = P + (int64_t)n * D * ts * plane;
  const int64_t tstride = w, lstride = (int64_t)ts * w;      // tap stride, line stride
  float m = -INFINITY, s = 0.f, t = 0.f;
  float c0 = 0.f, c1 = 0.f, c2 = 0.f;            // partial costs of output planes z'-1, z', z'+1
  auto retire = [&](int z, float c) {
    if (cost_out) cost_out[((int64_t)n * D + z) * plane + (int64_t)y * w + x] = c;
    const float v = sign * c;
    if (v > m) { const float sc = expf(m - v); s *= sc; t *= sc; m = v; }
    const float e = expf(v - m);  s += e;  t += e * (float)z;
  };
  for (int zi = 0; zi < D; ++zi) {
    const float* Pz = Pn + (int64_t)zi * ts * plane;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;          // contributions of input plane zi with kz = 2, 1, 0
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= w) continue;
        const float* p = Pz + (int64_t)yy * lstride + (int64_t)(ky * 3 + kx) * tstride + xx;
        a0 += __ldg(p + 18 * tstride);   // kz = 2: input plane zi feeds output plane zi - 1
        a1 += __ldg(p + 9 * tstride);    // kz = 1: output plane zi
        a2 += __ldg(p);                // kz = 0: output plane zi + 1
      }
    }
    c0 += a0;  c1 += a1;  c2 += a2;
    if (zi >= 1) retire(zi - 1, c0);
    c0 = c1;  c1 = c2;  c2 = 0.f;
  }
  retire(D - 1, c0);
  disp[(int64_t)n * plane + (int64_t)y * w + x] = t / s;
}

// ------------------------------------------------------------------------------------------
// fused correlation + soft-argmax
// ------------------------------------------------------------------------------------------
constexpr int kXG = 4;   // x per thread
constexpr int kDG = 8;   // d per thread

template <typename T, bool kLeftRef>
__device__ __forceinline__ void corr_row(const T* __restrict__ ref_g, const T* __restrict__ tgt_g,
                                         float* __restrict__ disp_row, float* __restrict__ cost_row,
                                         int64_t cost_plane, int w, int C, int D, float inv_c, float* smem) {
  const int WP = ((w + 3) & ~3);
  const int rs = WP + 4;                       // ref row stride (floats); +4 spreads banks
  const int ts = WP + D + 12 + 4;              // tgt row stride: zero pad D + window slack
  float* refT = smem;                          // [C][rs]
  float* tgtT = smem + C * rs;                 // [C][ts]
  // zero fill (padding must read as 0), then transposed fill
  for (int i = threadIdx.x; i < C * ts; i += blockDim.x) tgtT[i] = 0.f;
  for (int i = threadIdx.x; i < C * rs; i += blockDim.x) refT[i] = 0.f;
  __syncthreads();
  const int shift = kLeftRef ? D : 0;
  for (int i = threadIdx.x; i < w * C; i += blockDim.x) {
    const int x = i / C, c = i % C;
    refT[c * rs + x] = to_f32(ref_g[i]);
    tgtT[c * ts + x + shift] = to_f32(tgt_g[i]);
  }
  __syncthreads();

  const int ndg = D / kDG;                     // lanes sharing one x-group (power of two <= 16)
  const int nxg = WP / kXG;
  const int groups_per_pass = blockDim.x / ndg;
  const int dgi = threadIdx.x % ndg;
  const int d0 = dgi * kDG;
  for (int xg0 = 0; xg0 < nxg; xg0 += groups_per_pass) {
    const int xg = xg0 + threadIdx.x / ndg;
    const bool active = xg < nxg;
    const int x0 = (active ? xg : 0) * kXG;
    const int base = kLeftRef ? (x0 - d0 - 8 + D) : (x0 + d0);
    float acc[kXG][kDG];
#pragma unroll
    for (int i = 0; i < kXG; ++i)
#pragma unroll
      for (int j = 0; j < kDG; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < C; ++c) {
      const float4 r4 = *reinterpret_cast<const float4*>(refT + c * rs + x0);
      const float4* tp = reinterpret_cast<const float4*>(tgtT + c * ts + base);
      const float4 t0 = tp[0], t1 = tp[1], t2 = tp[2];
      const float r[4] = {r4.x, r4.y, r4.z, r4.w};
      const float t[12] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w};
#pragma unroll
      for (int i = 0; i < kXG; ++i)
#pragma unroll
        for (int j = 0; j < kDG; ++j) acc[i][j] = fmaf(r[i], t[kLeftRef ? (8 + i - j) : (i + j)], acc[i][j]);
    }
    float res[kXG];
#pragma unroll
    for (int i = 0; i < kXG; ++i) {
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < kDG; ++j) { acc[i][j] *= inv_c; m = fmaxf(m, acc[i][j]); }
      for (int o = ndg >> 1; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s = 0.f, tt = 0.f;
#pragma unroll
      for (int j = 0; j < kDG; ++j) { const float e = __expf(acc[i][j] - m); s += e; tt += e * (float)(d0 + j); }
      for (int o = ndg >> 1; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        tt += __shfl_xor_sync(0xffffffffu, tt, o);
      }
      res[i] = tt / s;
    }
    if (active) {
      if (dgi == 0) {
#pragma unroll
        for (int i = 0; i < kXG; ++i) if (x0 + i < w) disp_row[x0 + i] = res[i];
      }
      if (cost_row) {
#pragma unroll
        for (int i = 0; i < kXG; ++i)
#pragma unroll
          for (int j = 0; j < kDG; ++j)
            if (x0 + i < w) cost_row[(int64_t)(d0 + j) * cost_plane + x0 + i] = acc[i][j];
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
corr_soft_argmin_kernel(const T* __restrict__ feat, float* __restrict__ disp, float* __restrict__ cost_out,
                        int B, int h, int w, int C, int D, float inv_c) {
  extern __shared__ float smem_f[];
  const int n = blockIdx.x / h, y = blockIdx.x % h;
  const bool left_ref = n < B;
  const int tgt_n = left_ref ? n + B : n - B;
  const T* ref_g = feat + ((int64_t)n * h + y) * w * C;
  const T* tgt_g = feat + ((int64_t)tgt_n * h + y) * w * C;
  float* disp_row = disp + ((int64_t)n * h + y) * w;
  const int64_t plane = (int64_t)h * w;
  float* cost_row = cost_out ? cost_out + (int64_t)n * D * plane + (int64_t)y * w : nullptr;
  if (left_ref) corr_row<T, true>(ref_g, tgt_g, disp_row, cost_row, plane, w, C, D, inv_c, smem_f);
  else          corr_row<T, false>(ref_g, tgt_g, disp_row, cost_row, plane, w, C, D, inv_c, smem_f);
}

// ------------------------------------------------------------------------------------------
// bilinear upsample (align_corners = False), PyTorch index / weight convention
// ------------------------------------------------------------------------------------------
__global__ void upsample_disp_kernel(const float* __restrict__ q, float* __restrict__ out, int N, int h, int w,
                                     int H, int W, float scale) {
  const int64_t total = (int64_t)N * H * W;
  const float ry = (float)h / (float)H, rx = (float)w / (float)W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % W), Y = (int)((i / W) % H);
    const int64_t n = i / ((int64_t)W * H);
    float sy = ((float)Y + 0.5f) * ry - 0.5f;  if (sy < 0.f) sy = 0.f;
    float sx = ((float)X + 0.5f) * rx - 0.5f;  if (sx < 0.f) sx = 0.f;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float* p = q + n * h * w;
    const float v = hy * (hx * p[y0 * w + x0] + lx * p[y0 * w + x1]) + ly * (hx * p[y1 * w + x0] + lx * p[y1 * w + x1]);
    out[i] = v * scale;
  }
}

// The same, four consecutive outputs per thread (one 16-byte store), 32-bit index arithmetic: the scalar kernel spends its time
// in 64-bit divisions (0.051 ms for 33.5 MB at B = 64, 8x its HBM time).  Same formula per output, bit-identical results.
__global__ void __launch_bounds__(256)
upsample_disp_v4_kernel(const float* __restrict__ q, float4* __restrict__ out, uint32_t total4, int h, int w, int H, int W4,
                        float ry, float rx, float scale) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const uint32_t X4 = i % (uint32_t)W4, row = i / (uint32_t)W4;
  const uint32_t Y = row % (uint32_t)H, n = row / (uint32_t)H;
  float sy = ((float)Y + 0.5f) * ry - 0.5f;  if (sy < 0.f) sy = 0.f;
  const int y0 = (int)sy, y1 = y0 + (y0 < h - 1 ? 1 : 0);
  const float ly = sy - (float)y0, hy = 1.f - ly;
  const float* p0 = q + ((int64_t)n * h + y0) * w;
  const float* p1 = q + ((int64_t)n * h + y1) * w;
  float r[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int X = 4 * (int)X4 + k;
    float sx = ((float)X + 0.5f) * rx - 0.5f;  if (sx < 0.f) sx = 0.f;
    const int x0 = (int)sx, x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float lx = sx - (float)x0, hx = 1.f - lx;
    const float v = hy * (hx * __ldg(p0 + x0) + lx * __ldg(p0 + x1)) + ly * (hx * __ldg(p1 + x0) + lx * __ldg(p1 + x1));
    r[k] = v * scale;
  }
  out[i] = make_float4(r[0], r[1], r[2], r[3]);
}

}  // namespace
}  // namespace s3d

using namespace s3d;

namespace s3d {
bool corr_tc_eligible(int w, int C, int D, int dtype);
int corr_tc_launch(const void* feat, float* disp, int B, int h, int w, int C, int D, float inv_c, cudaStream_t stream);
}

extern "C" int s3d_cost_volume_concat(const void* feat, void* vol, int B, int h, int w, int C, int D, int dtype,
                                      void* stream) {
  if (!feat || !vol) { set_error("cost_volume_concat: null argument"); return S3D_ERR_INVALID; }
  S3D_CHECK_ARG(dtype == S3D_DTYPE_F32 || dtype == S3D_DTYPE_BF16 || dtype == S3D_DTYPE_BF16X2, "cost_volume_concat: bad dtype");
  const int esz = dtype == S3D_DTYPE_F32 ? 4 : 2;
  S3D_CHECK_ARG(B > 0 && h > 0 && w > 0 && D > 0 && C > 0 && (C * esz) % 16 == 0,
                "cost_volume_concat: C*elem must be a multiple of 16 B");
  const int vpp = C * esz / 16;
  const bool split = dtype == S3D_DTYPE_BF16X2;
  const size_t smem = (size_t)(split ? 4 : 2) * w * vpp * sizeof(uint4);
  S3D_CHECK_ARG(smem <= 200 * 1024, "cost_volume_concat: row too large for shared memory");
  if (split) {
    S3D_CUDA(cudaFuncSetAttribute(concat_volume_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    concat_volume_split_kernel<<<2 * B * h, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(feat), static_cast<uint4*>(vol), B, h, w, vpp, D);
    S3D_LAUNCH_CHECK();
    return S3D_OK;
  }
  S3D_CUDA(cudaFuncSetAttribute(concat_volume_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  concat_volume_kernel<<<2 * B * h, 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(feat), static_cast<uint4*>(vol), B, h, w, vpp, D);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

extern "C" int s3d_soft_argmin(const float* cost, float* disp, int N, int D, int h, int w, float sign, void* stream) {
  if (!cost || !disp) { set_error("soft_argmin: null argument"); return S3D_ERR_INVALID; }
  S3D_CHECK_ARG(N > 0 && D > 0 && h > 0 && w > 0, "soft_argmin: bad shape");
  const int64_t plane = (int64_t)h * w, total = plane * N;
  if (plane % 4 == 0 && ((reinterpret_cast<uintptr_t>(cost) | reinterpret_cast<uintptr_t>(disp)) & 15) == 0) {
    soft_argmin_v4_kernel<<<(unsigned)ceil_div64(total / 4, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(cost), reinterpret_cast<float4*>(disp), total / 4, plane / 4, D, sign);
    S3D_LAUNCH_CHECK();
    return S3D_OK;
  }
  soft_argmin_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      cost, disp, total, plane, D, sign);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

extern "C" int s3d_corr_soft_argmin(const void* feat, float* disp, float* cost_out, int B, int h, int w, int C, int c_real,
                                    int D, int dtype, void* stream) {
  if (!feat || !disp) { set_error("corr_soft_argmin: null argument"); return S3D_ERR_INVALID; }
  S3D_CHECK_ARG(dtype == S3D_DTYPE_F32 || dtype == S3D_DTYPE_BF16, "corr_soft_argmin: bad dtype");
  S3D_CHECK_ARG(B > 0 && h > 0 && w > 0 && C > 0 && D > 0, "corr_soft_argmin: bad shape");
  S3D_CHECK_ARG(c_real >= 0 && c_real <= C, "corr_soft_argmin: c_real must be in [0, C] (0 = C)");
  // the mean runs over the REAL feature channels; the padded ones are zero and only widen the rows
  const float inv_c = 1.f / (float)(c_real > 0 ? c_real : C);
  // bf16 features, rows of up to 64 pixels, no debug cost output: Gram-matrix kernel on the tensor cores (corr_tc.cu)
  if (!cost_out && corr_tc_eligible(w, C, D, dtype))
    return corr_tc_launch(feat, disp, B, h, w, C, D, inv_c, static_cast<cudaStream_t>(stream));
  S3D_CHECK_ARG(D >= 8 && D <= 128 && (D & (D - 1)) == 0, "corr_soft_argmin (SIMT path): D must be a power of two in [8,128]");
  const int WP = (w + 3) & ~3;
  const size_t smem = (size_t)C * ((WP + 4) + (WP + D + 16)) * sizeof(float);
  S3D_CHECK_ARG(smem <= 200 * 1024, "corr_soft_argmin: row too large for shared memory");
  const int ndg = D / kDG;
  int threads = (WP / kXG) * ndg;
  threads = ((threads + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (threads < 32) threads = 32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == S3D_DTYPE_BF16) {
    S3D_CUDA(cudaFuncSetAttribute(corr_soft_argmin_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    corr_soft_argmin_kernel<__nv_bfloat16><<<2 * B * h, threads, smem, st>>>(
        static_cast<const __nv_bfloat16*>(feat), disp, cost_out, B, h, w, C, D, inv_c);
  } else {
    S3D_CUDA(cudaFuncSetAttribute(corr_soft_argmin_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    corr_soft_argmin_kernel<float><<<2 * B * h, threads, smem, st>>>(
        static_cast<const float*>(feat), disp, cost_out, B, h, w, C, D, inv_c);
  }
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

extern "C" int s3d_upsample_disp(const float* disp_q, float* disp, int N, int h, int w, int H, int W, float scale,
                                 void* stream) {
  if (!disp_q || !disp) { set_error("upsample_disp: null argument"); return S3D_ERR_INVALID; }
  S3D_CHECK_ARG(N > 0 && h > 0 && w > 0 && H > 0 && W > 0, "upsample_disp: bad shape");
  const int64_t total = (int64_t)N * H * W;
  if (W % 4 == 0 && total / 4 < (1ll << 31) && (reinterpret_cast<uintptr_t>(disp) & 15) == 0) {
    const uint32_t total4 = (uint32_t)(total / 4);
    upsample_disp_v4_kernel<<<(total4 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        disp_q, reinterpret_cast<float4*>(disp), total4, h, w, H, W / 4, (float)h / (float)H, (float)w / (float)W, scale);
    S3D_LAUNCH_CHECK();
    return S3D_OK;
  }
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
  upsample_disp_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(disp_q, disp, N, h, w, H, W, scale);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}

extern "C" int s3d_tap_gather_soft_argmin(const float* taps, float* disp, float* cost_out, int N, int D, int h, int w,
                                          int tap_stride, float sign, void* stream) {
  if (!taps || !disp) { set_error("tap_gather_soft_argmin: null argument"); return S3D_ERR_INVALID; }
  S3D_CHECK_ARG(N > 0 && D > 0 && h > 0 && w > 0 && tap_stride >= 27 && h < 65536 && N < 65536,
                "tap_gather_soft_argmin: bad shape");
  const int threads = w <= 32 ? 32 : (w <= 64 ? 64 : 128);
  dim3 grid(ceil_div(w, threads), h, N);
  tap_gather_soft_argmin_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(taps, disp, cost_out, N, D, h, w,
                                                                                      tap_stride, sign);
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}
