"""Run one conv config a few times (for ncu / timing).  usage: prof_conv.py NAME [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from stereo_3d_reconstruction_b200 import lib
from stereo_3d_reconstruction_b200.layers import PackedConv
name = sys.argv[1] if len(sys.argv) > 1 else 'agg'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(0)
if name == 'agg':          # the dominant layer: 64->64 3x3x3 at D=32, 64x64, 128 volumes (B=64 pairs)
    conv = nn.Conv3d(64, 64, 3, 1, 1, bias=False); shape = (128, 32, 64, 64, 64)
elif name == 'agg32':
    conv = nn.Conv3d(64, 64, 3, 1, 1, bias=False); shape = (32, 32, 64, 64, 64)
elif name == 'mrg':
    conv = nn.Conv3d(16, 16, 3, 1, 1, bias=False); shape = (128, 32, 32, 32, 16)
elif name == 'mrg64':
    conv = nn.Conv3d(64, 16, 3, 1, 1, bias=False); shape = (128, 32, 32, 32, 64)
elif name == 'mrg_big':
    conv = nn.Conv3d(16, 16, 3, 1, 1, bias=False); shape = (32, 32, 64, 64, 16)
elif name == 'agg128':
    conv = nn.Conv3d(64, 128, 3, 1, 1, bias=False); shape = (32, 32, 64, 64, 64)
elif name == 'pw_planar' or name == 'pw_cl':
    pass
elif name == 'enc':
    conv = nn.Conv2d(64, 64, 3, 1, 1, bias=False); shape = (128, 1, 64, 64, 64)
if name.startswith('pw_'):
    shape = (128, 32, 64, 64, 64)
    pc = PackedConv.from_pointwise(torch.randn(27, 64) * 0.1, None, None, lib.ACT_NONE, lib.DTYPE_BF16, 'cuda')
else:
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, lib.DTYPE_BF16, 'cuda')
x = torch.randn(*shape, device='cuda').to(torch.bfloat16)
out = torch.empty(*shape[:4], pc.cout_pad, device='cuda', dtype=torch.bfloat16)
kw = {}
if name == 'pw_planar':
    N_, D_, h_, w_, S_ = 128, 32, 64, 64, 32
    out = torch.empty(N_, D_, h_, S_, w_, device='cuda', dtype=torch.float32)
    kw = dict(out_view=(0, (D_ * h_ * S_ * w_, h_ * S_ * w_, S_ * w_, 1, w_)), cout_store=S_)
elif name == 'pw_cl':
    out = torch.empty(128, 32, 64, 64, 32, device='cuda', dtype=torch.float32)
for _ in range(2):
    pc(x, out=out, **kw)
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(iters):
    pc(x, out=out, **kw)
t1.record(); torch.cuda.synchronize()
ms = t0.elapsed_time(t1) / iters
print('%s: %.3f ms  %.1f TFLOP/s' % (name, ms, pc.flops(*shape[:4]) / ms / 1e9))
