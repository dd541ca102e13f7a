// The lean per-layer-shape instantiations of the plane-scatter kernel (conv_scatter.cuh): one kernel per layer shape of the
// network, each containing only its own producer / issuer / epilogue code.
#include "conv_scatter.cuh"

namespace s3d {
namespace scatter {

KernFn spec_kernel(int rb, int cp, bool res, bool relu, bool two) {
  if (two) return (rb == 32 && cp == 16 && !res && !relu) ? conv_scatter_kernel<false, true, 32, 16, 0, 2, false, true> : nullptr;
  if (rb == 128 && cp == 64 && !res && relu)  return conv_scatter_kernel<false, true, 128, 64, 0, 0>;   // aggregation
  if (rb == 128 && cp == 64 && res && !relu)  return conv_scatter_kernel<false, true, 128, 64, 1, 2>;   // residual layers
  if (rb == 32 && cp == 16 && !res && !relu)  return conv_scatter_kernel<false, true, 32, 16, 0, 2>;    // fusion scorer
  if (rb == 64 && cp == 32 && !res && relu)   return conv_scatter_kernel<false, true, 64, 32, 0, 0>;    // enc1
  if (rb == 128 && cp == 32 && !res && !relu) return conv_scatter_kernel<false, true, 128, 32, 0, 2>;   // enc5
  if (rb == 64 && cp == 64 && !res && relu)   return conv_scatter_kernel<false, true, 64, 64, 0, 0>;    // blocked deconv
  return nullptr;
}

}  // namespace scatter
}  // namespace s3d
