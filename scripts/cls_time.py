import os, sys
sys.path.insert(0, '/root/repo')
import torch
from stereo_3d_reconstruction_b200 import ops
x = torch.randn(128, 32, 64, 64, 64, device='cuda').to(torch.bfloat16)
w = torch.randn(32, 64, device='cuda').to(torch.bfloat16)
out = torch.empty(128, 64, 64, device='cuda')
for _ in range(3): ops.cls_soft_argmin(x, w, -1.0, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.cls_soft_argmin(x, w, -1.0, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print('cls_fused dbg=%s: %.3f ms  %.2f TB/s' % (os.environ.get('S3D_CLS_DBG', '0'), ms, x.numel() * 2 / ms / 1e9))
