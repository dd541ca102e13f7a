// Plain SIMT convolution with the exact S3dConvParams contract of the tcgen05 engine.
// fp32 FMA accumulation in a fixed (tap, channel) order.  This is the 'fp32' precision mode of
// the product (exact, slow) and the on-GPU cross-check for conv_igemm.cu; it is still a CUDA
// kernel -- there is no CPU path anywhere in this library.
#include "common.cuh"

namespace s3d {
namespace {

template <typename TIn>
__global__ void conv_direct_kernel(const S3dConvParams p, const TIn* __restrict__ in, const TIn* __restrict__ w,
                                   const float* __restrict__ bias, const void* __restrict__ residual,
                                   void* __restrict__ out, int64_t total) {
  const int cs = p.cout_store;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = idx;
    const int co = (int)(r % cs);  r /= cs;
    const int x = (int)(r % p.oW);  r /= p.oW;
    const int y = (int)(r % p.oH);  r /= p.oH;
    const int z = (int)(r % p.oD);  r /= p.oD;
    const int n = (int)(r % p.N);   r /= p.N;
    const int cls = (int)r;
    // conv output (bias + activation) of channel `c` at this position; the fused projection re-evaluates it for
    // every input channel of the 1x1 (this engine is the exact validation path, not the fast one)
    auto conv_at = [&](int c) -> float {
      float a = 0.f;
      for (int t = 0; t < p.ntaps; ++t) {
        const int ti = cls * p.ntaps + t;
        const int xi = x * p.sx + p.dx[ti], yi = y * p.sy + p.dy[ti], zi = z * p.sz + p.dz[ti];
        if (xi < 0 || xi >= p.iW || yi < 0 || yi >= p.iH || zi < 0 || zi >= p.iD) continue;
        const TIn* ip = in + ((((int64_t)n * p.iD + zi) * p.iH + yi) * p.iW + xi) * p.Cin;
        const TIn* wp = w + ((int64_t)ti * p.Cout + c) * p.Cin;
        for (int ci = 0; ci < p.Cin; ++ci) a = fmaf(to_f32(ip[ci]), to_f32(wp[ci]), a);
      }
      if (bias) a += bias[c];
      return a;
    };
    float acc;
    bool projected = false;
    if (p.proj_w && co == p.proj_channel) {
      float pr = 0.f;
      for (int c = 0; c < p.Cout; ++c) {
        const float wc = p.proj_w[c];
        if (wc != 0.f) pr = fmaf(apply_act(conv_at(c), p.act, p.act_param), wc, pr);
      }
      acc = apply_act(pr, p.proj_act, 1.f);
      projected = true;
    } else {
      acc = conv_at(co);
    }
    const int ooz = (cls >> 2) & 1, ooy = (cls >> 1) & 1, oox = cls & 1;
    const int64_t off = (int64_t)n * p.osN + (int64_t)(z * p.omz + ooz) * p.osD +
                        (int64_t)(y * p.omy + ooy) * p.osH + (int64_t)(x * p.omx + oox) * p.osW + (int64_t)co * p.osC;
    if (p.out_dtype == S3D_DTYPE_BF16) {
      if (residual && !projected) acc += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(residual)[off]);
      reinterpret_cast<__nv_bfloat16*>(out)[off] = __float2bfloat16_rn(projected ? acc : apply_act(acc, p.act, p.act_param));
    } else {
      if (residual && !projected) acc += reinterpret_cast<const float*>(residual)[off];
      reinterpret_cast<float*>(out)[off] = projected ? acc : apply_act(acc, p.act, p.act_param);
    }
  }
}

}  // namespace
}  // namespace s3d

extern "C" int s3d_conv_direct(const S3dConvParams* p, const void* in, const void* w, const float* bias,
                               const void* residual, void* out, void* stream) {
  using namespace s3d;
  if (!p || !in || !w || !out) { set_error("s3d_conv_direct: null argument"); return S3D_ERR_INVALID; }
  S3D_CHECK_ARG(p->ntaps >= 1 && (p->n_classes == 1 || p->n_classes == 8) &&
                p->ntaps * p->n_classes <= S3D_MAX_TAPS, "direct: bad tap table");
  S3D_CHECK_ARG(p->cout_store >= 1 && p->cout_store <= p->Cout, "direct: cout_store");
  const int64_t total = (int64_t)p->n_classes * p->N * p->oD * p->oH * p->oW * p->cout_store;
  if (total == 0) return S3D_OK;
  const int threads = 256;
  int64_t blocks = ceil_div64(total, threads);
  const int64_t cap = (int64_t)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p->in_dtype == S3D_DTYPE_BF16)
    conv_direct_kernel<__nv_bfloat16><<<(int)blocks, threads, 0, st>>>(
        *p, static_cast<const __nv_bfloat16*>(in), static_cast<const __nv_bfloat16*>(w), bias, residual, out, total);
  else if (p->in_dtype == S3D_DTYPE_F32)
    conv_direct_kernel<float><<<(int)blocks, threads, 0, st>>>(
        *p, static_cast<const float*>(in), static_cast<const float*>(w), bias, residual, out, total);
  else { set_error("direct: bad in_dtype"); return S3D_ERR_INVALID; }
  S3D_LAUNCH_CHECK();
  return S3D_OK;
}
