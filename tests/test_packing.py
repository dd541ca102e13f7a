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
