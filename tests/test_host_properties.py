"""Property tests (hypothesis) of the host logic: tile chooser, sharding, and PackedConv tap tables vs
torch.nn.functional on random small shapes (CPU emulation of the S3dConvParams contract)."""
import torch
import torch.nn as nn
import torch.nn.functional as F
from hypothesis import given, settings, strategies as st

from stereo_3d_reconstruction_b200 import lib
from stereo_3d_reconstruction_b200.core import test as T
from stereo_3d_reconstruction_b200.layers import PackedConv, _choose_tile
from tests.emulate import emulate, to_cl, pad_c


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 300), st.integers(1, 40), st.integers(1, 80), st.integers(1, 80), st.sampled_from([1, 2]))
def test_choose_tile_valid(N, D, H, W, s):
    tw, th, td, tn = _choose_tile(N, D, H, W, s, s, 1)
    assert tw * th * td * tn == 128
    for v in (tw, th, td, tn):
        assert v & (v - 1) == 0
    assert tw * s <= 256 and th * s <= 256


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 2000), st.integers(1, 16))
def test_shard_range_is_a_partition(n, world):
    rs = [T.shard_range(n, r, world) for r in range(world)]
    assert rs[0][0] == 0 and rs[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
    sizes = [hi - lo for lo, hi in rs]
    assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes


@settings(max_examples=12, deadline=None)
@given(st.integers(1, 9), st.integers(1, 9), st.integers(3, 9), st.integers(3, 9), st.sampled_from([1, 2]), st.integers(0, 10 ** 6))
def test_conv2d_packing_matches_torch(cin, cout, H, W, stride, seed):
    torch.manual_seed(seed)
    conv = nn.Conv2d(cin, cout, 3, stride, 1)
    x = torch.randn(2, cin, H, W)
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    out = emulate(pc, pad_c(to_cl(x), pc.cin_pad))
    torch.testing.assert_close(out[..., :cout], to_cl(conv(x)), rtol=1e-4, atol=1e-5)


@settings(max_examples=8, deadline=None)
@given(st.integers(1, 6), st.integers(1, 6), st.integers(1, 3), st.integers(1, 3), st.integers(1, 3), st.integers(0, 10 ** 6))
def test_deconv_packing_matches_torch(cin, cout, D, H, W, seed):
    torch.manual_seed(seed)
    dc = nn.ConvTranspose3d(cin, cout, 4, 2, 1, bias=False)
    x = torch.randn(1, cin, D, H, W)
    pc = PackedConv.from_deconv_k4s2p1(dc, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    out = emulate(pc, pad_c(to_cl(x), pc.cin_pad))
    torch.testing.assert_close(out[..., :cout], to_cl(dc(x)), rtol=1e-4, atol=1e-5)


def test_zstack_weights_layout():
    """weight_zs[sv*9+kyx] rows 0..63 = W[kz=sv], rows 64..127 = W[kz=sv-1], zeros out of range (conv_halo.cu)."""
    torch.manual_seed(0)
    conv = nn.Conv3d(16, 24, 3, 1, 1, bias=False)
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    assert pc.weight_zs.shape[0] == 38 and not pc.zs_ident and pc.weight_zs[36:].abs().sum() == 0      # Cin != Cout: no identity blocks
    zs = pc.weight_zs[:36].view(4, 9, 128, pc.cin_pad)
    w = pc.weight.view(3, 9, pc.cout_pad, pc.cin_pad)
    for sv in range(4):
        top = w[sv] if sv <= 2 else torch.zeros_like(w[0])
        bot = w[sv - 1] if sv >= 1 else torch.zeros_like(w[0])
        assert torch.equal(zs[sv, :, :pc.cout_pad], top) and torch.equal(zs[sv, :, 64:64 + pc.cout_pad], bot)
        assert zs[sv, :, pc.cout_pad:64].abs().sum() == 0 and zs[sv, :, 64 + pc.cout_pad:].abs().sum() == 0
    pc2 = PackedConv.from_conv(nn.Conv3d(32, 32, 3, 1, 1), None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    assert pc2.zs_ident and torch.equal(pc2.weight_zs[36, :32], torch.eye(32)) and torch.equal(pc2.weight_zs[37, 64:96], torch.eye(32))
    assert pc2.weight_zs[36, 32:].abs().sum() == 0 and pc2.weight_zs[37, :64].abs().sum() == 0
    conv2 = nn.Conv3d(16, 128, 3, 1, 1)
    assert PackedConv.from_conv(conv2, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu').weight_zs is None
