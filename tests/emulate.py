"""CPU emulation of the S3dConvParams contract (include/s3d.h) -- TEST infrastructure only.

Executes a PackedConv's tap table + packed weights with plain torch ops so the host-side packing
(BN folding, sub-pixel classes of transposed convs, Linear-as-conv, channel padding) can be checked
against torch.nn.functional on a box without a GPU."""
import torch


def emulate(pc, x, out_channels=None):
    """pc: PackedConv packed on CPU (weight fp32/bf16 [rows,Cout_pad,Cin_pad]); x: [N,D,H,W,Cin_pad] fp32.
    Returns the dense channels-last output [N, oD*omz, oH*omy, oW*omx, Cout_pad] BEFORE activation."""
    N, iD, iH, iW, C = x.shape
    oD, oH, oW = pc.out_grid(iD, iH, iW)
    mz, my, mx = pc.out_mult
    w = pc.weight.float()
    out = torch.zeros(N, oD * mz, oH * my, oW * mx, pc.cout_pad)
    sz, sy, sx = pc.stride
    for cls, taps in enumerate(pc.taps):
        acc = torch.zeros(N, oD, oH, oW, pc.cout_pad)
        for t, (dz, dy, dx) in enumerate(taps):
            wt = w[cls * pc.ntaps + t]                         # [Cout_pad, Cin_pad]
            for z in range(oD):
                zi = z * sz + dz
                if zi < 0 or zi >= iD:
                    continue
                for y in range(oH):
                    yi = y * sy + dy
                    if yi < 0 or yi >= iH:
                        continue
                    xs = [xo for xo in range(oW) if 0 <= xo * sx + dx < iW]
                    if not xs:
                        continue
                    xi = [xo * sx + dx for xo in xs]
                    acc[:, z, y, xs] += x[:, zi, yi, xi] @ wt.t()
        acc = acc + pc.bias.float()
        if pc.n_classes == 8:
            cz, cy, cx = (cls >> 2) & 1, (cls >> 1) & 1, cls & 1
            out[:, cz::2, cy::2, cx::2] = acc
        else:
            out = acc
    return out


def to_cl(x):
    """NCHW / NCDHW -> channels-last 5-D [N,D,H,W,C]."""
    if x.dim() == 4:
        x = x.unsqueeze(2)
    return x.permute(0, 2, 3, 4, 1).contiguous()


def pad_c(x, c):
    if x.shape[-1] == c:
        return x
    out = torch.zeros(*x.shape[:-1], c, dtype=x.dtype)
    out[..., :x.shape[-1]] = x
    return out
