"""CPU check of the SHEARED form of the concat cost volume + first aggregation layer (layers.py, PackedConv.gonce_convs;
include/s3d.h, s3d_concat_gonce_assemble): the packed map convolutions are run through a plain-torch emulation of the conv engine
(explicit taps, zero outside the input), assembled exactly as gonce_assemble_kernel does, and compared with conv3d over the
oracle's concat volume -- both views, border planes, the edge column.  No GPU, no library call."""
import pytest
import torch
import torch.nn.functional as F

from oracle import models as O
from stereo_3d_reconstruction_b200 import lib
from stereo_3d_reconstruction_b200.layers import PackedConv


def engine(pc, x):
    """Emulation of s3d_conv_igemm for a PackedConv with explicit taps: x [N,1,H,W,C] -> fp32 [N,oH,oW,Cout]."""
    N, _, H, W, C = x.shape
    _, oH, oW = pc.out_grid(1, H, W)
    wt = pc.weight.float()                                           # [ntaps, cout_pad, cin_pad], bf16-rounded
    xp = F.pad(x.float()[:, 0], (0, 0, 64, 64, 4, 4))                # zero margins: W by 64, H by 4
    out = torch.zeros(N, oH, oW, wt.shape[1])
    for t, (dz, dy, dx) in enumerate(pc.taps[0]):
        out += xp[:, 4 + dy:4 + dy + oH, 64 + dx:64 + dx + oW] @ wt[t].T
    return out


@pytest.mark.parametrize('B,C,h,w,D', [(2, 16, 6, 20, 8), (1, 32, 5, 9, 12), (1, 16, 4, 16, 2)])
def test_sheared_form_equals_conv3d_over_the_concat_volume(B, C, h, w, D):
    torch.manual_seed(1)
    A = 64
    conv = torch.nn.Conv3d(2 * C, A, 3, padding=1)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, lib.DTYPE_BF16, 'cpu')
    L = torch.randn(B, C, h, w).to(torch.bfloat16).float()
    R = torch.randn(B, C, h, w).to(torch.bfloat16).float()
    pad = max(D, 4)
    featp = torch.zeros(2 * B, 1, h, w + 2 * pad, C)
    featp[:B, 0, :, pad:pad + w] = L.permute(0, 2, 3, 1)
    featp[B:, 0, :, pad:pad + w] = R.permute(0, 2, 3, 1)
    g = pc.gonce_convs(C, pad, w, D)
    ml, mr = engine(g['left'], featp[:B]), engine(g['right'], featp[B:])
    el, er = engine(g['edge_left'], featp[:B]), engine(g['edge_right'], featp[B:])
    assert ml.shape == (B, h, w + 4, 384) and el.shape == (B, h, D, 256)
    out = torch.zeros(2 * B, D, h, w, 64)
    bias = pc.bias.float()
    for n in range(2 * B):
        left = n < B
        b = n % B
        pm, gm, em = (ml, mr, er) if left else (mr, ml, el)
        for d in range(D):
            for x in range(w):
                v = bias + pm[b, :, x + 2, 0:64]
                col = (x - d if left else x + d) + 2
                ok = 0 <= col < w + 4
                if ok:
                    v = v + gm[b, :, col, 192:256]
                if d == 0:
                    v = v + pm[b, :, x + 2, 64:128] + (gm[b, :, col, 256:320] if ok else 0)
                if d == D - 1:
                    v = v + pm[b, :, x + 2, 128:192] + (gm[b, :, col, 320:384] if ok else 0)
                if x == (w - 1 if left else 0):
                    j = D - 1 - d if left else d
                    v = v + em[b, :, j, 0:64]
                    if d == 0:
                        v = v + em[b, :, j, 64:128]
                    if d == D - 1:
                        v = v + em[b, :, j, 128:192]
                out[n, d, :, x] = v.clamp(min=0)
    wq = conv.weight.detach().to(torch.bfloat16).float()
    ref = []
    for direction, r, t in ((-1, L, R), (1, R, L)):
        ref.append(F.relu(F.conv3d(O.build_concat_volume(r, t, D, direction), wq, conv.bias.detach(), padding=1)))
    ref = torch.cat(ref).permute(0, 2, 3, 4, 1)
    # the maps use weights summed over kz / (kz, kx) BEFORE the bf16 rounding: ~2^-9 of the summed weight per tap
    assert (out - ref).abs().max().item() <= 6e-3 * ref.abs().max().item()
