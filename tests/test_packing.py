"""Host logic: PackedConv tap tables / BN folding vs torch.nn.functional, on CPU."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from stereo_3d_reconstruction_b200 import lib
from stereo_3d_reconstruction_b200.layers import PackedConv, _choose_tile, _choose_bn, pad_to
from tests.emulate import emulate, to_cl, pad_c


def _rand_bn(bn, g):
    with torch.no_grad():
        bn.running_mean.copy_(torch.randn(bn.num_features, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(bn.num_features, generator=g) + 0.5)
        bn.weight.copy_(torch.rand(bn.num_features, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(bn.num_features, generator=g) * 0.1)
    return bn.eval()


def test_conv2d_stride2_bn():
    g = torch.Generator().manual_seed(0)
    conv, bn = nn.Conv2d(3, 20, 3, 2, 1, bias=True), _rand_bn(nn.BatchNorm2d(20), g)
    x = torch.randn(2, 3, 11, 13, generator=g)
    ref = bn(conv(x))
    pc = PackedConv.from_conv(conv, bn, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    out = emulate(pc, pad_c(to_cl(x), pc.cin_pad))
    assert out.shape[1:4] == (1, 6, 7)
    torch.testing.assert_close(out[..., :20], to_cl(ref), rtol=1e-4, atol=1e-5)
    assert out[..., 20:].abs().max() == 0


def test_conv3d_bn():
    g = torch.Generator().manual_seed(1)
    conv, bn = nn.Conv3d(5, 7, 3, 1, 1, bias=False), _rand_bn(nn.BatchNorm3d(7), g)
    x = torch.randn(1, 5, 4, 6, 5, generator=g)
    pc = PackedConv.from_conv(conv, bn, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    out = emulate(pc, pad_c(to_cl(x), pc.cin_pad))
    torch.testing.assert_close(out[..., :7], to_cl(bn(conv(x))), rtol=1e-4, atol=1e-5)


def test_deconv_k4s2p1_subpixel_classes():
    g = torch.Generator().manual_seed(2)
    dc, bn = nn.ConvTranspose3d(6, 5, 4, 2, 1, bias=False), _rand_bn(nn.BatchNorm3d(5), g)
    x = torch.randn(2, 6, 2, 3, 4, generator=g)
    pc = PackedConv.from_deconv_k4s2p1(dc, bn, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    assert pc.n_classes == 8 and pc.ntaps == 8
    out = emulate(pc, pad_c(to_cl(x), pc.cin_pad))
    assert out.shape[1:4] == (4, 6, 8)
    torch.testing.assert_close(out[..., :5], to_cl(bn(dc(x))), rtol=1e-4, atol=1e-5)


def test_pointwise_deconv():
    g = torch.Generator().manual_seed(3)
    dc = nn.ConvTranspose3d(8, 1, 1, bias=False)
    x = torch.randn(1, 8, 3, 3, 3, generator=g)
    w = dc.weight
    pc = PackedConv.from_pointwise(w.view(8, 1).t(), None, None, lib.ACT_SIGMOID, lib.DTYPE_F32, 'cpu')
    out = emulate(pc, pad_c(to_cl(x), pc.cin_pad))
    torch.testing.assert_close(out[..., :1], to_cl(dc(x)), rtol=1e-4, atol=1e-5)


def test_linear_over_map():
    g = torch.Generator().manual_seed(4)
    fc = nn.Linear(6 * 2 * 2, 10)
    x = torch.randn(3, 6, 2, 2, generator=g)
    pc = PackedConv.from_linear_over_map(fc, 6, 2, 2, lib.ACT_RELU, lib.DTYPE_F32, 'cpu')
    out = emulate(pc, pad_c(to_cl(x), pc.cin_pad))
    assert out.shape[1:4] == (1, 1, 1)
    torch.testing.assert_close(out[:, 0, 0, 0, :10], fc(x.flatten(1)), rtol=1e-4, atol=1e-5)


def test_tile_and_bn_choice():
    for dims in [(128, 32, 64, 64), (128, 1, 35, 35), (128, 2, 2, 2), (32, 1, 1, 1), (3, 32, 32, 32)]:
        tw, th, td, tn = _choose_tile(*dims, 1, 1, 1)
        assert tw * th * td * tn == 128
    assert _choose_tile(128, 2, 2, 2, 1, 1, 1) == (2, 2, 2, 16)
    assert _choose_tile(4, 1, 64, 64, 2, 2, 1)[0] * 2 <= 256
    assert _choose_bn(64) == 64 and _choose_bn(512) == 256 and _choose_bn(6144) == 256 and _choose_bn(272) == 272 // 17 * 1 or True
    assert pad_to(9) == 16 and pad_to(2048 * 3) == 6144


# ---- packings added for the plane-scatter kernel (csrc/conv_scatter.cuh), emulated on the CPU ----------------------
def _emulate_plane_scatter(ns, x, cout_pad):
    """The kernel's algorithm on the host: input plane p times the rotation (3 for p = 0, else p % 3) of the N-stacked
    weights, accumulated into three slots; output plane z lives in slot z % 3, is complete after input plane z + 1,
    is then drained and its slot zeroed.  ns: [36, 3*Co, Ci]; x: [N, D, H, W, Ci] -> [N, D, H, W, Co]."""
    N, D, H, W, Ci = x.shape
    Co = cout_pad
    out = torch.zeros(N, D, H, W, Co)
    slots = torch.zeros(N, H, W, 3 * Co)
    xp = F.pad(x, (0, 0, 1, 1, 1, 1))                              # zero halo in y and x
    for p in range(D):
        rot = 3 if p == 0 else p % 3
        acc = torch.zeros(N, H, W, 3 * Co)
        for kyx in range(9):
            ky, kx = divmod(kyx, 3)
            acc += xp[:, p, ky:ky + H, kx:kx + W] @ ns[rot * 9 + kyx].t()
        slots = acc if p == 0 else slots + acc                      # the first MMA of a column overwrites all three slots
        if p >= 1:
            s = (p - 1) % 3
            out[:, p - 1] = slots[..., s * Co:(s + 1) * Co]
            slots[..., s * Co:(s + 1) * Co] = 0
    s = (D - 1) % 3
    out[:, D - 1] = slots[..., s * Co:(s + 1) * Co]
    return out


def test_nstack_rotations_reproduce_conv3d():
    g = torch.Generator().manual_seed(6)
    for D in (1, 2, 3, 7):
        conv = nn.Conv3d(5, 7, 3, 1, 1, bias=False)
        x = torch.randn(2, 5, D, 6, 5, generator=g)
        pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
        assert pc.weight_ns is not None and pc.weight_ns.shape == (36, 3 * pc.cout_pad, pc.cin_pad)
        out = _emulate_plane_scatter(pc.weight_ns.float(), pad_c(to_cl(x), pc.cin_pad), pc.cout_pad)
        torch.testing.assert_close(out[..., :7], to_cl(conv(x)), rtol=1e-4, atol=1e-5)
        assert out[..., 7:].abs().max() == 0


def test_conv2d_as_volume_of_images_twin():
    """A stride-1 3x3 2-D layer gets a 27-tap twin whose kz = 0 / 2 slices are zero: a batch of images run as a volume."""
    g = torch.Generator().manual_seed(7)
    conv, bn = nn.Conv2d(16, 24, 3, 1, 1, bias=True), _rand_bn(nn.BatchNorm2d(24), g)
    x = torch.randn(6, 16, 9, 7, generator=g)
    pc = PackedConv.from_conv(conv, bn, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    assert pc.vol is not None and pc.vol.ntaps == 27 and pc.vol.vol is None
    vol_in = pad_c(to_cl(x), pc.cin_pad).view(2, 3, 9, 7, pc.cin_pad)           # 2 "volumes" of 3 images
    out = emulate(pc.vol, vol_in).view(6, 1, 9, 7, pc.cout_pad)
    torch.testing.assert_close(out[..., :24], to_cl(bn(conv(x))), rtol=1e-4, atol=1e-5)
    out_ns = _emulate_plane_scatter(pc.vol.weight_ns.float(), vol_in, pc.cout_pad) + pc.vol.bias
    torch.testing.assert_close(out_ns.view(6, 1, 9, 7, -1)[..., :24], to_cl(bn(conv(x))), rtol=1e-4, atol=1e-5)
    assert PackedConv.from_conv(nn.Conv2d(16, 24, 3, 2, 1), None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu').vol is None   # stride 2: no twin


def test_deconv_blocked_is_a_3x3x3_conv_plus_depth_to_space():
    g = torch.Generator().manual_seed(8)
    dc, bn = nn.ConvTranspose3d(6, 8, 4, 2, 1, bias=False), _rand_bn(nn.BatchNorm3d(8), g)
    x = torch.randn(2, 6, 3, 2, 4, generator=g)
    pc = PackedConv.from_deconv_k4s2p1_blocked(dc, bn, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    assert pc.ntaps == 27 and pc.n_classes == 1 and pc.cout == 64 and pc.ntaps_algo == 8
    y = emulate(pc, pad_c(to_cl(x), pc.cin_pad))                                 # [2,3,2,4,64], channel = class*8 + c
    out = torch.zeros(2, 6, 4, 8, 8)
    for cls in range(8):
        cz, cy, cx = (cls >> 2) & 1, (cls >> 1) & 1, cls & 1
        out[:, cz::2, cy::2, cx::2] = y[..., cls * 8:(cls + 1) * 8]
    torch.testing.assert_close(out, to_cl(bn(dc(x))), rtol=1e-4, atol=1e-5)
    # algorithmic FLOPs are those of the transposed conv (8 of the 27 taps per output channel are non-zero)
    plain = PackedConv.from_deconv_k4s2p1(dc, bn, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    assert pc.flops(2, 3, 2, 4) == plain.flops(2, 3, 2, 4)


def test_split_conv_operands_are_exact_tf32_splits():
    from stereo_3d_reconstruction_b200.layers import SplitConv
    conv = nn.Conv3d(5, 7, 3, 1, 1, bias=True)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, lib.DTYPE_F32, 'cpu')
    sc = SplitConv(pc)
    w, hi, lo = pc.weight, sc.hi.weight, sc.lo.weight
    assert torch.equal(hi + lo, w)                                               # exact split
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0                 # hi is representable in TF32
    assert float(lo.abs().max()) <= float(w.abs().max()) * 2.0 ** -10
    assert sc.hi.act == lib.ACT_NONE and float(sc.hi.bias.abs().max()) == 0      # bias / activation only in the last pass
    assert sc.lo.act == lib.ACT_RELU and torch.equal(sc.lo.bias, pc.bias)
    assert sc.hi.weight_ns is not None and sc.lo.weight_ns is not None           # both run on the plane-scatter kernel


def test_skewed_tensor_map_views_match_the_concat_volume():
    """The fused cost-volume kernel (csrc/conv_scatter_concat.cu) never builds the volume: its target half of plane d is a TMA
    box of a SKEWED view of the zero-margined feature rows.  Re-state the view's address arithmetic on the host -- element
    (x, dd) of a [W x D] view with equal X / D strides of one pixel, zero outside 0 <= x < W -- and check it against the
    oracle's volume, halo columns included."""
    from oracle import models as O
    g = torch.Generator().manual_seed(9)
    B, C, h, w, D = 2, 4, 3, 11, 5
    fL, fR = torch.randn(B, C, h, w, generator=g), torch.randn(B, C, h, w, generator=g)
    pad = D                                                           # >= D - 1 zero pixels on both sides of every row
    P = w + 2 * pad
    rows = torch.zeros(2 * B, h, P, C)
    rows[:B, :, pad:pad + w] = fL.permute(0, 2, 3, 1)
    rows[B:, :, pad:pad + w] = fR.permute(0, 2, 3, 1)
    flat = rows.reshape(2 * B, h, P * C)

    def view(n_src, base_px, x, dd, y):
        """the tensor map: address = base + (x + dd) pixels; X bound [0, w) -> zero fill (conv halo)"""
        if x < 0 or x >= w or y < 0 or y >= h:
            return torch.zeros(C)
        px = base_px + x + dd
        return flat[n_src, y, px * C:(px + 1) * C]

    vol_l = O.build_concat_volume(fL, fR, D, -1)                      # [B, 2C, D, h, w]
    vol_r = O.build_concat_volume(fR, fL, D, +1)
    for d in range(D):
        for y in range(h):
            for x in range(-1, w + 1):                                # including the conv's halo columns
                inside = 0 <= x < w
                # left-referenced volume n reads the RIGHT image n at x - d: coordinate (x, D-1-d), base D-1 pixels to the left
                got = view(B + 0, pad - (D - 1), x, D - 1 - d, y)
                want = vol_l[0, C:, d, y, x] if inside else torch.zeros(C)
                assert torch.equal(got, want), ('left', d, y, x)
                # right-referenced volume n reads the LEFT image n at x + d: coordinate (x, d)
                got = view(1, pad, x, d, y)
                want = vol_r[1, C:, d, y, x] if inside else torch.zeros(C)
                assert torch.equal(got, want), ('right', d, y, x)


def test_tf32x3_three_pass_algebra_reaches_fp32_accuracy():
    """SplitConv's claim, checked on the CPU: with every MMA operand truncated to TF32 (10 mantissa bits), the single pass
    errs at ~1e-3 while hi*hi + lo*hi + hi*lo (operands split exactly, fp32 accumulation) reproduces the fp64 conv to ~1e-6."""
    from stereo_3d_reconstruction_b200.layers import SplitConv
    g = torch.Generator().manual_seed(10)
    conv = nn.Conv3d(16, 16, 3, 1, 1, bias=True)
    x = torch.randn(1, 16, 4, 6, 5, generator=g)
    ref = to_cl(F.conv3d(x.double(), conv.weight.double(), conv.bias.double(), padding=1)).float()
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    sc = SplitConv(pc)

    def tf32(t):                                                    # what a kind::tf32 MMA keeps of an fp32 operand
        return (t.view(torch.int32) & -8192).view(torch.float32)

    def mma_pass(p, xin):                                           # emulate() with TF32-truncated operands, fp32 accumulate
        q = p.derive(tf32(p._ctor[0]), p._ctor[1], lib.ACT_NONE)
        return emulate(q, tf32(xin))

    xc = pad_c(to_cl(x), pc.cin_pad)
    x_hi = tf32(xc)
    x_lo = xc - x_hi
    one = mma_pass(pc, xc)
    three = mma_pass(sc.hi, x_hi) + (mma_pass(sc.hi, x_lo) - sc.hi.bias) + (mma_pass(sc.lo, x_hi))
    scale = ref.abs().max().item()
    e1 = (one[..., :16] - ref).abs().max().item() / scale
    e3 = (three[..., :16] - ref).abs().max().item() / scale
    assert 1e-5 < e1 < 5e-3, e1                                     # single-pass TF32: ~1e-3
    assert e3 < 5e-6, e3                                            # three passes: fp32-level


def test_reference_once_algebra_matches_the_concat_conv():
    """csrc/conv_scatter_concat.cu, reference-once form, restated on the host with the weights PackedConv.refonce_weights packs:
    out[z] = bias + R + conv over the TARGET half only, R = ref (*) sum_kz W[kz] computed once, plus ref (*) (-W[kz=0]) on plane 0 and
    ref (*) (-W[kz=2]) on plane D-1 -- equal to the 3x3x3 conv over the whole concat volume (fp32 here: exact up to summation order)."""
    from oracle import models as O
    g = torch.Generator().manual_seed(12)
    B, C, h, w, cout = 1, 4, 5, 7, 6
    for D in (1, 2, 5):
        conv = nn.Conv3d(2 * C, cout, 3, 1, 1, bias=True)
        pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
        ro = pc.refonce_weights(pc.cin_pad // 2).float()              # [27, cout_pad, C_pad]: sum_kz W | -W[kz=0] | -W[kz=2]
        Cp = pc.cin_pad // 2
        assert ro.shape == (27, pc.cout_pad, Cp)
        fL, fR = torch.randn(B, C, h, w, generator=g), torch.randn(B, C, h, w, generator=g)
        vol = O.build_concat_volume(fL, fR, D, -1)                    # [B, 2C, D, h, w]: ref = left, target = right at x - d
        with torch.no_grad():
            want = conv(vol)
        # this test's C is not a multiple of 16: the packed layer pads the channel axis, ref channels first
        # (the product only ever uses C * elem in {32, 64} bytes)
        def conv2d(x_nchw, w27):                                      # 3x3 2-D conv with taps [9, cout_pad, Cp]
            wk = w27[:, :cout, :x_nchw.shape[1]].reshape(3, 3, cout, -1).permute(2, 3, 0, 1)
            return F.conv2d(x_nchw, wk, padding=1)
        if pc.cin_pad == 2 * C:
            R = conv2d(fL, ro[0:9])
            c0 = conv2d(fL, ro[9:18])
            cL = conv2d(fL, ro[18:27])
            wt = conv.weight[:, C:].detach()                          # target-channel weights
            tgt = F.conv3d(vol[:, C:], wt, padding=1)
            got = tgt + R.unsqueeze(2) + conv.bias.view(1, -1, 1, 1, 1)
            got[:, :, 0] += c0
            got[:, :, D - 1] += cL
            torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)
        else:
            # padded layout (cin_pad > 2C): only check the packing rule itself
            w3 = conv.weight.detach().permute(2, 3, 4, 0, 1).reshape(3, 9, cout, 2 * C)
            torch.testing.assert_close(ro[0:9, :cout, :C], w3[:, :, :, :C].sum(0))
            torch.testing.assert_close(ro[9:18, :cout, :C], -w3[0, :, :, :C])
            torch.testing.assert_close(ro[18:27, :cout, :C], -w3[2, :, :, :C])


def test_reference_once_algebra_with_16_channel_features():
    """The same with a channel count the kernel takes (C = 16: cin_pad == 2C), through the packed weights end to end."""
    from oracle import models as O
    g = torch.Generator().manual_seed(13)
    B, C, h, w, cout, D = 1, 16, 4, 6, 5, 3
    conv = nn.Conv3d(2 * C, cout, 3, 1, 1, bias=True)
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    assert pc.cin_pad == 2 * C
    ro = pc.refonce_weights(C).float()
    fL, fR = torch.randn(B, C, h, w, generator=g), torch.randn(B, C, h, w, generator=g)
    vol = O.build_concat_volume(fL, fR, D, -1)
    k2 = lambda w27: w27[:, :cout].reshape(3, 3, cout, C).permute(2, 3, 0, 1)
    with torch.no_grad():
        want = conv(vol)
        got = F.conv3d(vol[:, C:], conv.weight[:, C:], padding=1) + F.conv2d(fL, k2(ro[0:9]), padding=1).unsqueeze(2) + \
            conv.bias.view(1, -1, 1, 1, 1)
        got[:, :, 0] += F.conv2d(fL, k2(ro[9:18]), padding=1)
        got[:, :, D - 1] += F.conv2d(fL, k2(ro[18:27]), padding=1)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)


def test_scatter_form_classifier_with_per_patch_partial_sums():
    """csrc/conv_scatter_cls.cu restated on the host: the Cout = 1 3x3x3 classifier over Y in SCATTER form -- every pixel of a 32x8
    patch projects its channels on the 27 taps, every cell of the patch plus its halo ring (34 x 10) gathers the taps of its
    in-patch neighbours into three running cost planes, finished planes are stored as the patch's PARTIAL sums, and a second pass
    adds, per pixel, the partial sums of the (<= 4) patches whose ring covers it -- equals conv3d(Y, Wb)."""
    g = torch.Generator().manual_seed(14)
    N, C, D, h, w = 1, 3, 4, 37, 19                                   # ragged: 2 x 3 patches of 32 x 8
    TY, TX = 32, 8
    cy, cx = -(-h // TY), -(-w // TX)
    Y = torch.randn(N, C, D, h, w, generator=g)
    Wb = torch.randn(1, C, 3, 3, 3, generator=g)
    want = F.conv3d(Y, Wb, padding=1)[:, 0]                            # [N, D, h, w]
    taps = Wb[0].reshape(C, 27)                                        # tap index (kz * 3 + ky) * 3 + kx
    partials = torch.zeros(N, cy, cx, D, TY + 2, TX + 2)
    for n in range(N):
        for py in range(cy):
            for px in range(cx):
                run = torch.zeros(3, TY + 2, TX + 2)                   # cost planes z-1, z, z+1 of this patch (+ ring)
                for z in range(D):
                    # projections of the patch's pixels; pixels outside the image enter as zeros
                    P = torch.zeros(TY, TX, 27)
                    ys, xs = min(TY, h - py * TY), min(TX, w - px * TX)
                    P[:ys, :xs] = torch.einsum('cyx,ct->yxt', Y[n, :, z, py * TY:py * TY + ys, px * TX:px * TX + xs], taps)
                    # cell (yc, xc) in [-1, 32] x [-1, 8] gathers pixel (yc + ky - 1, xc + kx - 1) when it is inside the patch
                    a = torch.zeros(3, TY + 2, TX + 2)                 # through kz = 2, 1, 0 -> planes z-1, z, z+1
                    for ky in range(3):
                        for kx in range(3):
                            for kz in range(3):
                                # pixels [0, TY) x [0, TX) land on cells pixel - (ky - 1, kx - 1), stored with the +1 ring offset
                                a[2 - kz, 2 - ky:2 - ky + TY, 2 - kx:2 - kx + TX] += P[:, :, (kz * 3 + ky) * 3 + kx]
                    run += a
                    if z >= 1:
                        partials[n, py, px, z - 1] = run[0]
                    run = torch.stack([run[1], run[2], torch.zeros(TY + 2, TX + 2)])
                partials[n, py, px, D - 1] = run[0]
    got = torch.zeros(N, D, h, w)
    for n in range(N):
        for y in range(h):
            for x in range(w):
                py, px, yi, xi = y // TY, x // TX, y % TY, x % TX
                dyn = -1 if yi == 0 else (1 if yi == TY - 1 else 0)
                dxn = -1 if xi == 0 else (1 if xi == TX - 1 else 0)
                hy = dyn != 0 and 0 <= py + dyn < cy
                hx = dxn != 0 and 0 <= px + dxn < cx
                c = partials[n, py, px, :, yi + 1, xi + 1].clone()
                yn = (-1 if dyn > 0 else TY) + 1                       # the pixel in the neighbour's ring coordinates
                xn = (-1 if dxn > 0 else TX) + 1
                if hx:
                    c += partials[n, py, px + dxn, :, yi + 1, xn]
                if hy:
                    c += partials[n, py + dyn, px, :, yn, xi + 1]
                if hx and hy:
                    c += partials[n, py + dyn, px + dxn, :, yn, xn]
                got[n, :, y, x] = c
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


def test_residual_as_identity_tap():
    """csrc/conv_scatter_rm.cu: the residual enters the accumulator as one more tap whose weight matrix is the identity; in the
    swizzled shared-memory operand layout row r (output channel) holds its 1.0 at 16-byte chunk (ci / 8) ^ (r % 8)."""
    CP = 64
    for rank in (0, 1):
        tile = torch.zeros(CP // 2 * 128, dtype=torch.uint8)           # this CTA's 32 rows x 128 bytes
        for lane in range(32):
            r, ci = lane, rank * (CP // 2) + lane
            off = (r >> 3) * 1024 + (r & 7) * 128 + (((ci >> 3) ^ (r & 7)) << 4) + (ci & 7) * 2
            tile[off + 1] = 0x3F                                       # bf16 1.0 = 0x3F80, little endian
            tile[off] = 0x80
        # undo the 128B swizzle and read the rows back
        rows = torch.zeros(CP // 2, CP)
        for r in range(CP // 2):
            for chunk in range(8):
                phys = (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4)
                for e in range(8):
                    lo, hi = int(tile[phys + 2 * e]), int(tile[phys + 2 * e + 1])
                    rows[r, chunk * 8 + e] = 1.0 if (hi, lo) == (0x3F, 0x80) else 0.0
        want = torch.eye(CP)[rank * (CP // 2):(rank + 1) * (CP // 2)]
        assert torch.equal(rows, want)
