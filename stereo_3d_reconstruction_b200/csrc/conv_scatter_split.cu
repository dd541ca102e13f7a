// Plane-scatter kernels for split (BF16X2, 'bf16x3' precision) operands: every tap issues (x_hi, w_hi), (x_lo, w_hi),
// (x_hi, w_lo) into the same TMEM accumulator -- fp32-grade results from kind::f16 MMAs, no split kernel and no partial sums
// through HBM (conv_scatter.cuh, ScArgs::split).  A separate translation unit so that it compiles in parallel and the bf16 /
// tf32 kernels keep their code size.
#include "conv_scatter.cuh"

namespace s3d {
namespace scatter {

KernFn split_kernel(bool lean, bool pair, int prow, int cp, bool res, bool relu) {
  if (lean && pair) {
    // the 64-wide aggregation layers: 256-byte [hi | lo] rows = two K chunks, one weight stage per (tap, hi | lo weights)
    if (prow == 256 && cp == 64 && !res && relu)  return conv_scatter_kernel<false, true, 256, 64, 0, 0, true>;
    if (prow == 256 && cp == 64 && res && !relu)  return conv_scatter_kernel<false, true, 256, 64, 1, 2, true>;
  }
  return pair ? conv_scatter_kernel<false, true, 0, 0, -1, -1, true> : conv_scatter_kernel<false, false, 0, 0, -1, -1, true>;
}

}  // namespace scatter
}  // namespace s3d
