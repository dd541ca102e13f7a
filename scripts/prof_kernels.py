"""A few launches of ONE kernel at the bench shape, for `ncu --set full -k regex:...`.
usage: prof_kernels.py <agg_bf16 | concat | concat_ro | cls_chain | agg_bf16x3 | agg_res_bf16x3 | cls_fused | corr_tc | soft_argmin | chamfer | conv_first>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from stereo_3d_reconstruction_b200 import lib, ops
from stereo_3d_reconstruction_b200.layers import PackedConv
from stereo_3d_reconstruction_b200.utils import synthetic

what = sys.argv[1]
torch.manual_seed(0)
N, D, h, w, C = 128, 32, 64, 64, 64                       # 2B = 128 volumes: the bench shape (B = 64)
if what.startswith('agg'):
    split = 'x3' in what
    res = 'res' in what
    code = lib.DTYPE_BF16X2 if split else lib.DTYPE_BF16
    pc = PackedConv.from_conv(nn.Conv3d(C, C, 3, 1, 1, bias=True), None, lib.ACT_NONE if res else lib.ACT_RELU, code, 'cuda')
    x = (torch.randn(N, D, h, w, C * (2 if split else 1), device='cuda') * (0.01 if split else 1)).to(torch.bfloat16)
    r = torch.randn_like(x) if res else None
    out = torch.empty_like(x)
    for _ in range(3):
        pc(x, out=out, residual=r)
elif what in ('concat', 'concat_ro'):
    Cf = 32
    pc = PackedConv.from_conv(nn.Conv3d(2 * Cf, C, 3, 1, 1, bias=True), None, lib.ACT_RELU, lib.DTYPE_BF16, 'cuda')
    featp = torch.zeros(N, 1, h, w + 2 * D, Cf, dtype=torch.bfloat16, device='cuda')
    featp[:, :, :, D:D + w] = torch.randn(N, 1, h, w, Cf, device='cuda').to(torch.bfloat16)
    out = torch.empty(N, D, h, w, C, dtype=torch.bfloat16, device='cuda')
    for _ in range(3):
        ops.conv_concat_volume(pc, featp, N // 2, D, D, out=out, ref_once=what == 'concat_ro')
elif what == 'cls_chain':
    pc = PackedConv.from_conv(nn.Conv3d(C, C, 3, 1, 1, bias=True), None, lib.ACT_RELU, lib.DTYPE_BF16, 'cuda')
    x = torch.randn(N, D, h, w, C, device='cuda').to(torch.bfloat16)
    wt = torch.zeros(32, C, dtype=torch.bfloat16, device='cuda')
    wt[:27] = (torch.randn(27, C, device='cuda') * 0.2).to(torch.bfloat16)
    ws = torch.empty(ops.conv_cls_workspace_bytes(N, D, h, w), dtype=torch.uint8, device='cuda')
    for _ in range(3):
        ops.conv_cls_soft_argmin(pc, x, wt, -1.0, workspace=ws)
elif what == 'cls_fused':
    x = torch.randn(N, D, h, w, C, device='cuda').to(torch.bfloat16)
    wt = torch.zeros(32, C, dtype=torch.bfloat16, device='cuda')
    wt[:27] = (torch.randn(27, C, device='cuda') * 0.2).to(torch.bfloat16)
    for _ in range(3):
        ops.cls_soft_argmin(x, wt, -1.0)
elif what == 'corr_tc':
    f = torch.randn(N, 1, h, w, 32, device='cuda').to(torch.bfloat16)
    for _ in range(3):
        ops.corr_soft_argmin(f, N // 2, 64)
elif what == 'soft_argmin':
    c = torch.randn(N, 128, h, w, device='cuda')
    for _ in range(3):
        ops.soft_argmin(c, -1.0)
elif what == 'chamfer':
    a, b = synthetic.point_clouds(32, 2048, 16384, seed=2, device='cuda')
    for _ in range(3):
        ops.chamfer_forward(a, b)
torch.cuda.synchronize()
print('done', what)
