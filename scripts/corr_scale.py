import sys, os
sys.path.insert(0, '/root/repo')
import torch
from stereo_3d_reconstruction_b200 import ops
def t(fn, it=20):
    for _ in range(5): fn()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/it
for B in (8, 64, 256, 1024):
    f = torch.randn(2*B,1,64,64,32,device='cuda').to(torch.bfloat16)
    d = torch.empty(2*B,64,64,device='cuda')
    ms = t(lambda: ops.corr_soft_argmin(f, B, 32, out=d))
    print('B=%d: %.4f ms (warm, back-to-back)  per tile/SM %.0f ns' % (B, ms, ms*1e6/(2*B*32/148)))
