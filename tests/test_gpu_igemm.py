"""tcgen05 implicit-GEMM conv engine vs the torch fp32 reference of the same op (GPU)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from stereo_3d_reconstruction_b200 import lib
from stereo_3d_reconstruction_b200.layers import PackedConv
from tests.emulate import to_cl, pad_c

pytestmark = pytest.mark.gpu

# tolerances relative to the output's max magnitude: bf16 inputs are rounded once (2^-9 each),
# products accumulate in fp32; TF32 keeps 10 mantissa bits.
TOL = {'bf16': 2e-2, 'tf32': 3e-3}


def _check(pc, x_nc, ref_nc, prec, cout, **kw):
    dt = torch.bfloat16 if prec == 'bf16' else torch.float32
    x = pad_c(to_cl(x_nc), pc.cin_pad).to(dt).cuda()
    got = pc(x, engine='igemm', **kw).float().cpu()
    ref = to_cl(ref_nc)
    assert got.shape[:4] == ref.shape[:4]
    err = (got[..., :cout] - ref).abs().max().item()
    assert err <= TOL[prec] * (ref.abs().max().item() + 1e-6), (err, ref.abs().max().item())
    if got.shape[-1] > cout and pc.act != lib.ACT_SIGMOID:
        assert got[..., cout:].abs().max().item() == 0.0


def _code(prec):
    return lib.DTYPE_BF16 if prec == 'bf16' else lib.DTYPE_F32


@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
@pytest.mark.parametrize('cin,cout,H,W,stride', [(3, 32, 32, 32, 2), (32, 64, 35, 35, 1), (64, 64, 16, 16, 1),
                                                 (16, 32, 37, 41, 2), (128, 256, 9, 9, 2), (256, 320, 8, 8, 1)])
def test_conv2d(prec, cin, cout, H, W, stride):
    torch.manual_seed(0)
    conv = nn.Conv2d(cin, cout, 3, stride, 1)
    x = torch.randn(3, cin, H, W)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, _code(prec), 'cuda')
    _check(pc, x, F.relu(conv(x)), prec, cout)


@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
@pytest.mark.parametrize('cin,cout,D,H,W', [(64, 64, 8, 16, 16), (32, 32, 5, 9, 11), (9, 16, 8, 8, 8), (64, 1, 4, 8, 8)])
def test_conv3d(prec, cin, cout, D, H, W):
    torch.manual_seed(1)
    conv = nn.Conv3d(cin, cout, 3, 1, 1, bias=False)
    x = torch.randn(2, cin, D, H, W)
    pc = PackedConv.from_conv(conv, None, lib.ACT_LEAKY, _code(prec), 'cuda', act_param=0.2)
    _check(pc, x, F.leaky_relu(conv(x), 0.2), prec, cout)


@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
@pytest.mark.parametrize('cin,cout,S,N', [(64, 32, 2, 5), (32, 8, 8, 2), (128, 64, 4, 3)])
def test_deconv(prec, cin, cout, S, N):
    torch.manual_seed(2)
    dc = nn.ConvTranspose3d(cin, cout, 4, 2, 1, bias=False)
    x = torch.randn(N, cin, S, S, S)
    pc = PackedConv.from_deconv_k4s2p1(dc, None, lib.ACT_RELU, _code(prec), 'cuda')
    _check(pc, x, F.relu(dc(x)), prec, cout)


@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
@pytest.mark.parametrize('shape', [(2, 5, 6, 7), (3, 16, 16, 16), (1, 1, 1, 1)])
def test_deconv_blocked_plus_depth_to_space(prec, shape):
    """ConvTranspose3d(32, 8, 4, 2, 1) + BN + ReLU as ONE 3x3x3 conv with 8 parity classes x 8 channels on the
    plane-scatter kernel, then depth-to-space with the 1x1x1 + sigmoid projection in channel 8."""
    from stereo_3d_reconstruction_b200 import ops
    torch.manual_seed(5)
    N, d, h, w = shape
    dc = nn.ConvTranspose3d(32, 8, 4, 2, 1, bias=False)
    bn = nn.BatchNorm3d(8).eval()
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.1);  bn.running_var.uniform_(0.5, 1.5);  bn.weight.uniform_(0.8, 1.2);  bn.bias.normal_(0, 0.1)
    x = torch.randn(N, 32, d, h, w)
    with torch.no_grad():
        ref = to_cl(F.relu(bn(dc(x))))                                # [N,2d,2h,2w,8]
    pw = torch.randn(8) * 0.5
    dt = torch.bfloat16 if prec == 'bf16' else torch.float32
    pc = PackedConv.from_deconv_k4s2p1_blocked(dc, bn, lib.ACT_RELU, _code(prec), 'cuda')
    y = pc(pad_c(to_cl(x), pc.cin_pad).to(dt).cuda(), engine='igemm')     # [N,d,h,w,64]
    assert y.shape == (N, d, h, w, 64)
    out = ops.depth_to_space(y, 16, pw.cuda(), lib.ACT_SIGMOID).float().cpu()
    assert out.shape == (N, 2 * d, 2 * h, 2 * w, 16)
    scale = ref.abs().max().item() + 1e-6
    assert (out[..., :8] - ref).abs().max().item() <= TOL[prec] * scale
    proj_ref = torch.sigmoid((out[..., :8] * pw).sum(-1))                 # from the stored (rounded) features
    assert (out[..., 8] - proj_ref).abs().max().item() <= (1e-2 if prec == 'bf16' else 1e-5)
    assert out[..., 9:].abs().max().item() == 0.0
    # same layer through the generic engine's 8-class formulation
    pc8 = PackedConv.from_deconv_k4s2p1(dc, bn, lib.ACT_RELU, _code(prec), 'cuda')
    old = pc8(pad_c(to_cl(x), pc8.cin_pad).to(dt).cuda(), engine='igemm').float().cpu()
    assert (out[..., :8] - old[..., :8]).abs().max().item() <= TOL[prec] * scale


@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
def test_residual_and_linear(prec):
    torch.manual_seed(3)
    conv = nn.Conv2d(32, 32, 3, 1, 1)
    x = torch.randn(2, 32, 12, 12)
    dt = torch.bfloat16 if prec == 'bf16' else torch.float32
    res = torch.randn(2, 32, 12, 12).to(dt).float()
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, _code(prec), 'cuda')
    _check(pc, x, conv(x) + res, prec, 32, residual=to_cl(res).to(dt).cuda())
    fc = nn.Linear(32 * 4 * 4, 100)
    xf = torch.randn(7, 32, 4, 4)
    pc = PackedConv.from_linear_over_map(fc, 32, 4, 4, lib.ACT_TANH, _code(prec), 'cuda', act_param=0.5)
    ref = (0.5 * torch.tanh(fc(xf.flatten(1)))).view(7, 100, 1, 1)
    _check(pc, xf, ref, prec, 100)


def test_igemm_rejects_bad_shapes():
    conv = nn.Conv2d(16, 16, 3, 1, 1)
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, lib.DTYPE_BF16, 'cuda')
    x = torch.zeros(1, 1, 8, 8, 16, dtype=torch.bfloat16).cuda()
    p = pc.params(1, 1, 8, 8, (8 * 8 * 16, 8 * 8 * 16, 8 * 16, 16), lib.DTYPE_BF16, 16)
    import ctypes
    q = type(p).from_buffer_copy(p); q.tw = 3
    out = torch.zeros(1, 1, 8, 8, 16, dtype=torch.bfloat16).cuda()
    rc = lib.load().s3d_conv_igemm(ctypes.byref(q), x.data_ptr(), pc.weight.data_ptr(), None, None, out.data_ptr(), None)
    assert rc == -1 and b'tile' in lib.load().s3d_last_error()


@pytest.mark.parametrize('knob', [None, 'S3D_SCATTER_NO_PAIR', 'S3D_SCATTER_GENERIC', 'S3D_SCATTER_NO_TRANSPOSE',
                                  'S3D_SCATTER_RES_TRANSPOSE', 'S3D_SCATTER_TPS3', 'S3D_NO_SCATTER'])
@pytest.mark.parametrize('prec,N,cin,cout,D,H,W,res,act', [
    ('bf16', 1, 64, 64, 1, 8, 8, False, 'relu'),        # one column, one plane (no CTA pair possible)
    ('bf16', 1, 64, 64, 2, 33, 9, True, 'none'),        # ragged patches in y and x, residual, 2 planes
    ('bf16', 3, 64, 48, 3, 40, 20, False, 'relu'),      # 48 channels: 6 chunks per pixel, non-transposed epilogue
    ('bf16', 2, 16, 16, 7, 32, 32, False, 'leaky'),     # fusion-scorer shape (32-byte rows, 9-tap stages)
    ('bf16', 1, 32, 32, 4, 70, 70, True, 'relu'),       # 64-byte rows, odd number of columns
    ('bf16', 5, 64, 64, 6, 16, 24, True, 'none'),       # the residual layer's kernel
    ('bf16', 2, 64, 32, 5, 24, 16, False, 'none'),
    ('tf32', 2, 32, 32, 5, 20, 12, True, 'relu'),       # fp32 storage, 128-byte rows
    ('tf32', 1, 16, 64, 3, 9, 33, False, 'leaky'),
    ('tf32', 2, 64, 64, 5, 20, 12, True, 'relu'),       # 256-byte rows: two K chunks per tap (the tf32 aggregation layers)
    ('bf16', 1, 128, 64, 3, 16, 16, False, 'relu'),     # 128 bf16 channels: two K chunks as well
])
def test_conv3d_plane_scatter_shapes(monkeypatch, knob, prec, N, cin, cout, D, H, W, res, act):
    """conv_scatter.cu over edge shapes, in every mode its knobs select (CTA pairs / single CTA, lean per-shape kernels /
    all-in-one kernel, transposed / direct stores and residual reads, 9- / 3-tap stages) and the z-stacked fallback."""
    if knob:
        monkeypatch.setenv(knob, '1')
    torch.manual_seed(7)
    conv = nn.Conv3d(cin, cout, 3, 1, 1, bias=True)
    x = torch.randn(N, cin, D, H, W)
    dt = torch.bfloat16 if prec == 'bf16' else torch.float32
    code = {'relu': lib.ACT_RELU, 'none': lib.ACT_NONE, 'leaky': lib.ACT_LEAKY}[act]
    fn = {'relu': F.relu, 'none': lambda t: t, 'leaky': lambda t: F.leaky_relu(t, 0.2)}[act]
    pc = PackedConv.from_conv(conv, None, code, _code(prec), 'cuda', act_param=0.2)
    kw = {}
    ref = conv(x)
    if res:
        r = torch.randn(N, cout, D, H, W).to(dt).float()
        ref = ref + r
        kw['residual'] = pad_c(to_cl(r), pc.cout_pad).to(dt).cuda()
    _check(pc, x, fn(ref), prec, cout, **kw)


@pytest.mark.parametrize('knob', [None, 'S3D_SCATTER_NO_PAIR'])
@pytest.mark.parametrize('prec,B,C,cout,D,h,w', [
    ('bf16', 2, 32, 64, 8, 16, 16),       # the network's shape (64-byte halves, lean kernel)
    ('bf16', 1, 16, 32, 5, 9, 13),        # ragged patches, 32-byte halves
    ('bf16', 3, 32, 64, 32, 40, 64),      # D = 32 planes, two patch rows
    ('bf16', 1, 64, 48, 3, 33, 10),       # 128-byte halves, 48 output channels
    ('bf16', 2, 32, 64, 1, 8, 8),         # a single disparity plane
    ('bf16', 1, 32, 64, 12, 8, 8),        # more disparities than pixels in a row: planes d >= w have an all-zero target half
    ('tf32', 1, 32, 64, 4, 20, 12),       # fp32 storage: 128-byte halves
    ('tf32', 2, 16, 32, 6, 12, 24),
])
def test_conv_concat_volume_fused_is_bit_identical(monkeypatch, knob, prec, B, C, cout, D, h, w):
    """conv_scatter_concat.cu: cost volume (never written) + 3x3x3 conv == s3d_cost_volume_concat + the same conv, bit for
    bit (same MMA sequence), and == the oracle's volume + conv3d within the engine's tolerance."""
    from stereo_3d_reconstruction_b200 import ops
    from oracle import models as O
    dt = torch.bfloat16 if prec == 'bf16' else torch.float32
    if knob:
        if C * (2 if prec == 'bf16' else 4) == 128:
            pytest.skip('128-byte halves need CTA pairs (a single CTA cannot hold two weight stages of 36-48 KB)')
        monkeypatch.setenv(knob, '1')
    torch.manual_seed(9)
    conv = nn.Conv3d(2 * C, cout, 3, 1, 1, bias=True)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, _code(prec), 'cuda')
    f = torch.randn(2 * B, C, h, w)
    feat = to_cl(f).to(dt).cuda()                                         # [2B,1,h,w,C]
    pad, P = D, w + 2 * D
    featp = torch.zeros(2 * B, 1, h, P, C, dtype=dt, device='cuda')
    featp[:, :, :, pad:pad + w] = feat
    got = ops.conv_concat_volume(pc, featp, B, D, pad)
    vol = ops.cost_volume_concat(feat, B, D)
    two = pc(vol, engine='igemm')
    assert got.shape == two.shape == (2 * B, D, h, w, pc.cout_pad)
    assert torch.equal(got, two)
    fr = feat.float().cpu()[:, 0].permute(0, 3, 1, 2)                     # rounded features, NCHW
    ref_vol = torch.cat([O.build_concat_volume(fr[:B], fr[B:], D, -1), O.build_concat_volume(fr[B:], fr[:B], D, +1)], 0)
    ref = to_cl(F.relu(conv(ref_vol)))
    err = (got.float().cpu()[..., :cout] - ref).abs().max().item()
    assert err <= TOL[prec] * (ref.abs().max().item() + 1e-6)


@pytest.mark.parametrize('D,H,W', [(4, 16, 16), (5, 37, 19), (2, 64, 64)])
def test_conv3d_residual_on_tensor_core(D, H, W, monkeypatch):
    monkeypatch.setenv('S3D_NO_SCATTER', '1')           # the z-stacked kernel is the fallback of conv_scatter.cu now
    _conv3d_residual_on_tensor_core(D, H, W)


def _conv3d_residual_on_tensor_core(D, H, W):
    """64->64 3x3x3 bf16 with a residual: the halo kernel adds the residual as an identity 'tap' (TMA-staged
    residual plane x [I;0] / [0;I] weight blocks) instead of loading it in the epilogue."""
    torch.manual_seed(5)
    conv = nn.Conv3d(64, 64, 3, 1, 1, bias=True)
    x = torch.randn(2, 64, D, H, W)
    res = torch.randn(2, 64, D, H, W).to(torch.bfloat16).float()
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, lib.DTYPE_BF16, 'cuda')
    assert pc.zs_ident
    _check(pc, x, F.relu(conv(x) + res), 'bf16', 64, residual=to_cl(res).to(torch.bfloat16).cuda())


def test_zstack_without_host_stacked_weights():
    """A direct C caller may leave w_zstack NULL: the kernel then assembles each stacked weight stage from two
    TMA boxes of the plain [taps][Cout][Cin] tensor (out-of-range kz = TMA zero fill).  Same result."""
    import ctypes
    torch.manual_seed(6)
    conv = nn.Conv3d(64, 48, 3, 1, 1, bias=True)
    x = torch.randn(2, 64, 5, 19, 13)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, lib.DTYPE_BF16, 'cuda')
    xc = pad_c(to_cl(x), pc.cin_pad).to(torch.bfloat16).cuda()
    ref = pc(xc, engine='igemm')
    out = torch.zeros_like(ref)
    C = out.shape[-1]
    p = pc.params(2, 5, 19, 13, (5 * 19 * 13 * C, 19 * 13 * C, 13 * C, C), lib.DTYPE_BF16, C)
    q = type(p).from_buffer_copy(p)
    q.w_zstack = None
    q.w_zstack_ident = 0
    rc = lib.load().s3d_conv_igemm(ctypes.byref(q), xc.data_ptr(), pc.weight.data_ptr(), pc.bias.data_ptr(), None,
                                   out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.load().s3d_last_error()
    torch.cuda.synchronize()
    assert torch.equal(out, ref)
