"""corr_tc_kernel time against the batch size (back-to-back launches over operand copies > L2): what part of a launch is fixed?"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from stereo_3d_reconstruction_b200 import ops
C, D = 32, 64
for B in (1, 4, 16, 64, 128, 256):
    feat = (torch.randn(2 * B, 1, 64, 64, C, device='cuda') * 0.5).to(torch.bfloat16)
    disp = torch.empty(2 * B, 64, 64, device='cuda')
    nrot = min(64, max(2, -(-320 * 2 ** 20 // (feat.numel() * 2))))
    feats = [feat] + [feat.clone() for _ in range(nrot - 1)]
    ms = bench._rotating_ms(lambda f: (lambda: ops.corr_soft_argmin(f, B, D, out=disp)), feats)
    print('B %4d: %.4f ms per launch   (%d pair tiles, %.1f per CTA)' % (B, ms, B * 32, B * 32 / 148), flush=True)
