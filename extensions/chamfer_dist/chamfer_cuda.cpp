// Thin torch C++ extension over the C ABI (include/s3d.h): the drop-in for the reference's `extensions/chamfer_dist`
// (/root/reference/README.md:62-65 -- `cd extensions/chamfer_dist && python setup.py install --user`; its source lives on the
// upstream Stereo2Point branch and is not on disk, so the module surface below -- `chamfer.forward(xyz1, xyz2)` returning
// (dist1, dist2, idx1, idx2) -- follows the same author's public GRNet extension from memory [RECALL]).
//
// No kernel lives here: the shim validates the tensors, allocates the outputs (and the workspace) with torch, and calls s3d_chamfer_forward_ws
// (libs3d_b200.so, csrc/chamfer.cu) on the current CUDA stream.  There is no CPU path.
#include <torch/extension.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <vector>
#include "s3d.h"

static std::vector<torch::Tensor> chamfer_forward(torch::Tensor xyz1, torch::Tensor xyz2) {
  TORCH_CHECK(xyz1.is_cuda() && xyz2.is_cuda(), "chamfer_dist: CUDA tensors required (there is no CPU fallback)");
  TORCH_CHECK(xyz1.scalar_type() == torch::kFloat32 && xyz2.scalar_type() == torch::kFloat32, "chamfer_dist: float32 points");
  TORCH_CHECK(xyz1.dim() == 3 && xyz2.dim() == 3 && xyz1.size(2) == 3 && xyz2.size(2) == 3 && xyz1.size(0) == xyz2.size(0),
              "chamfer_dist: expected [B,N,3] and [B,M,3]");
  TORCH_CHECK(xyz1.device() == xyz2.device(), "chamfer_dist: both sets on the same device");
  xyz1 = xyz1.contiguous();
  xyz2 = xyz2.contiguous();
  const int64_t B = xyz1.size(0), N = xyz1.size(1), M = xyz2.size(1);
  c10::cuda::CUDAGuard guard(xyz1.device());
  auto f = xyz1.options();
  auto i = xyz1.options().dtype(torch::kInt32);
  torch::Tensor dist1 = torch::empty({B, N}, f), dist2 = torch::empty({B, M}, f);
  torch::Tensor idx1 = torch::empty({B, N}, i), idx2 = torch::empty({B, M}, i);
  // large problems: one pass for both directions, with a torch-owned workspace (0 bytes when the problem is small)
  const int64_t ws_bytes = (B > 0 && N > 0 && M > 0) ? s3d_chamfer_workspace_bytes((int)B, (int)N, (int)M) : 0;
  torch::Tensor ws = torch::empty({ws_bytes > 0 ? ws_bytes : 0}, xyz1.options().dtype(torch::kUInt8));
  const int rc = s3d_chamfer_forward_ws(xyz1.data_ptr<float>(), xyz2.data_ptr<float>(), dist1.data_ptr<float>(), idx1.data_ptr<int32_t>(),
                                        dist2.data_ptr<float>(), idx2.data_ptr<int32_t>(), (int)B, (int)N, (int)M,
                                        ws_bytes > 0 ? ws.data_ptr() : nullptr, ws_bytes,
                                        c10::cuda::getCurrentCUDAStream().stream());
  TORCH_CHECK(rc == S3D_OK, "s3d_chamfer_forward_ws failed (rc=", rc, "): ", s3d_last_error());
  return {dist1, dist2, idx1, idx2};
}

static std::vector<torch::Tensor> chamfer_backward(torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
                                                   torch::Tensor) {
  TORCH_CHECK(false, "chamfer_dist backward is out of scope of the inference-only build (SURVEY.md 2, row 12)");
  return {};
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("forward", &chamfer_forward, "Chamfer nearest neighbours, forward (CUDA, sm_100a): (dist1, dist2, idx1, idx2)");
  m.def("backward", &chamfer_backward, "not built (inference-only)");
}
