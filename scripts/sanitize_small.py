"""One small launch of each mbarrier / cluster-protocol kernel, for compute-sanitizer (racecheck / synccheck / memcheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_small.py scatter_pair

usage: sanitize_small.py <case>      cases: scatter_pair scatter_single scatter_rm scatter_split concat concat_ro sheared halo2d cls_fused corr_tc chamfer_sym igemm all"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from stereo_3d_reconstruction_b200 import lib, ops
from stereo_3d_reconstruction_b200.layers import PackedConv, to_storage

case = sys.argv[1] if len(sys.argv) > 1 else 'all'
torch.manual_seed(0)


def conv3(cin, cout, code, act=lib.ACT_RELU):
    return PackedConv.from_conv(nn.Conv3d(cin, cout, 3, 1, 1, bias=True), None, act, code, 'cuda')


def run(name):
    if name in ('scatter_pair', 'scatter_single', 'scatter_rm'):
        # scatter_pair: residual read by the epilogue threads (knob scatter_no_rm); scatter_rm: residual as an identity tap
        lib.set_knob('scatter_no_pair', int(name == 'scatter_single'))
        lib.set_knob('scatter_no_rm', int(name == 'scatter_pair'))
        pc = conv3(64, 64, lib.DTYPE_BF16, lib.ACT_NONE)
        x = torch.randn(2, 4, 33, 9, 64, device='cuda').to(torch.bfloat16)
        r = torch.randn(2, 4, 33, 9, 64, device='cuda').to(torch.bfloat16)
        pc(x, residual=r)
        lib.set_knob('scatter_no_pair', 0)
        lib.set_knob('scatter_no_rm', 0)
    elif name == 'scatter_split':
        pc = conv3(64, 64, lib.DTYPE_BF16X2)
        x = to_storage(torch.randn(2, 4, 33, 9, 64), lib.DTYPE_BF16X2).cuda()
        pc(x)
        pc = conv3(16, 16, lib.DTYPE_BF16X2, lib.ACT_LEAKY)
        x = to_storage(torch.randn(1, 3, 16, 16, 16), lib.DTYPE_BF16X2).cuda()
        pc(x)
    elif name in ('concat', 'concat_ro'):
        B, C, D, h, w = 1, 32, 4, 16, 16
        pc = conv3(2 * C, 64, lib.DTYPE_BF16)
        featp = torch.zeros(2 * B, 1, h, w + 2 * D, C, dtype=torch.bfloat16, device='cuda')
        featp[:, :, :, D:D + w] = torch.randn(2 * B, 1, h, w, C, device='cuda').to(torch.bfloat16)
        ops.conv_concat_volume(pc, featp, B, D, D, ref_once=name == 'concat_ro')
    elif name == 'sheared':
        # map_conv_kernel<5> / <3> (TMA patch ring, resident weight chunk, two TMEM buffers) + gonce_assemble_kernel
        B, C, D, h, w = 1, 32, 6, 19, 21
        pc = conv3(2 * C, 64, lib.DTYPE_BF16)
        featp = torch.zeros(2 * B, 1, h, w + 2 * D, C, dtype=torch.bfloat16, device='cuda')
        featp[:, :, :, D:D + w] = torch.randn(2 * B, 1, h, w, C, device='cuda').to(torch.bfloat16)
        ops.conv_concat_volume_sheared(pc, featp, B, D, D)
    elif name == 'halo2d':
        # conv2d_halo_kernel (two MMA issuers, TMA patch ring, transposed bf16 epilogue): 64 -> 64 with residual, 64 -> 32 plain
        x = torch.randn(3, 1, 19, 21, 64, device='cuda').to(torch.bfloat16)
        pc = PackedConv.from_conv(nn.Conv2d(64, 64, 3, 1, 1), None, lib.ACT_NONE, lib.DTYPE_BF16, 'cuda')
        pc(x, residual=torch.randn(3, 1, 19, 21, 64, device='cuda').to(torch.bfloat16))
        pc = PackedConv.from_conv(nn.Conv2d(64, 32, 3, 1, 1), None, lib.ACT_RELU, lib.DTYPE_BF16, 'cuda')
        pc(x)
    elif name == 'chamfer_sym':
        # chamfer_sym_kernel<8> / <4> + chamfer_sym_finish_kernel: shared-memory tiles, shuffles, 64-bit atomicMin keys
        from stereo_3d_reconstruction_b200.utils import synthetic
        a, b = synthetic.point_clouds(2, 300, 2500, seed=3, duplicates=True, device='cuda')
        lib.set_knob('chamfer_sym', 1)
        for r in (8, 4):
            lib.set_knob('chamfer_sym_r', r)
            ops.chamfer_forward(a, b)
        lib.set_knob('chamfer_sym', 0); lib.set_knob('chamfer_sym_r', 0)
    elif name == 'cls_fused':
        x = torch.randn(1, 4, 33, 9, 64, device='cuda').to(torch.bfloat16)
        wt = torch.zeros(32, 64, dtype=torch.bfloat16, device='cuda')
        wt[:27] = torch.randn(27, 64, device='cuda').to(torch.bfloat16) * 0.2
        ops.cls_soft_argmin(x, wt, -1.0)
    elif name == 'corr_tc':
        f = torch.randn(2, 1, 5, 40, 32, device='cuda').to(torch.bfloat16)
        ops.corr_soft_argmin(f, 1, 16)
    elif name == 'igemm':
        pc = PackedConv.from_conv(nn.Conv2d(32, 64, 3, 2, 1), None, lib.ACT_RELU, lib.DTYPE_BF16, 'cuda')
        pc(torch.randn(2, 1, 19, 23, 32, device='cuda').to(torch.bfloat16))
    else:
        raise SystemExit('unknown case %s' % name)
    torch.cuda.synchronize()
    print('ran', name)


for n in (['scatter_pair', 'scatter_single', 'scatter_rm', 'scatter_split', 'concat', 'concat_ro', 'sheared', 'halo2d', 'cls_fused', 'corr_tc', 'chamfer_sym', 'igemm'] if case == 'all' else [case]):
    run(n)
