"""tcgen05 implicit-GEMM conv engine vs the torch fp32 reference of the same op (GPU).

Inputs AND weights are rounded to the engine's operand precision BEFORE the fp32 reference is computed (bf16: round to
nearest; tf32: the low 13 mantissa bits cleared, which is what kind::tf32 reads), so the comparison sees only the
engine's own arithmetic (fp32 accumulation order, output rounding) and the tolerances can be tight: one wrong tap out of
27 is ~4 % of the output's max and fails by an order of magnitude."""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from stereo_3d_reconstruction_b200 import lib
from stereo_3d_reconstruction_b200.layers import PackedConv
from tests.emulate import to_cl, pad_c

pytestmark = pytest.mark.gpu

# relative to the output's max magnitude.  bf16: the output is STORED in bf16 (2^-9 = 2e-3 of each value); tf32: fp32
# storage, only the accumulation order differs.
TOL = {'bf16': 4e-3, 'tf32': 2e-4}


def _q(t, prec):
    """Round a tensor to the operand precision of the engine."""
    t = t.detach().float()
    if prec == 'bf16':
        return t.to(torch.bfloat16).float()
    return (t.contiguous().view(torch.int32) & -8192).view(torch.float32)


def _qmod(m, prec):
    """The layer with its weight rounded to the operand precision (bias stays fp32: it is added in the epilogue)."""
    m = copy.deepcopy(m)
    with torch.no_grad():
        m.weight.copy_(_q(m.weight, prec))
    return m


def _check(pc, x_nc, ref_nc, prec, cout, **kw):
    """x_nc must already be rounded (_q) and ref_nc computed from rounded weights (_qmod)."""
    dt = torch.bfloat16 if prec == 'bf16' else torch.float32
    x = pad_c(to_cl(x_nc), pc.cin_pad).to(dt).cuda()
    assert torch.equal(x.float().cpu()[..., :x_nc.shape[1]], to_cl(x_nc)), 'test inputs must be pre-rounded'
    got = pc(x, engine='igemm', **kw).float().cpu()
    ref = to_cl(ref_nc)
    assert got.shape[:4] == ref.shape[:4]
    err = (got[..., :cout] - ref).abs().max().item()
    assert err <= TOL[prec] * (ref.abs().max().item() + 1e-6), (err, ref.abs().max().item())
    if got.shape[-1] > cout and pc.act != lib.ACT_SIGMOID:
        assert got[..., cout:].abs().max().item() == 0.0


def _code(prec):
    return lib.DTYPE_BF16 if prec == 'bf16' else lib.DTYPE_F32


@pytest.mark.parametrize('ts1', [0, 1])
@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
@pytest.mark.parametrize('cin,cout,H,W,stride', [(3, 32, 32, 32, 2), (32, 64, 35, 35, 1), (64, 64, 16, 16, 1),
                                                 (16, 32, 37, 41, 2), (128, 256, 9, 9, 2), (256, 320, 8, 8, 1)])
def test_conv2d(knobs, ts1, prec, cin, cout, H, W, stride):
    """Generic engine, several taps per pipeline stage (default) and one (knob igemm_ts1)."""
    knobs('igemm_ts1', ts1)
    torch.manual_seed(0)
    conv = _qmod(nn.Conv2d(cin, cout, 3, stride, 1), prec)
    x = _q(torch.randn(3, cin, H, W), prec)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, _code(prec), 'cuda')
    with torch.no_grad():
        _check(pc, x, F.relu(conv(x)), prec, cout)


@pytest.mark.parametrize('halo', [1, 0])
@pytest.mark.parametrize('cin,cout,N,H,W,act,res', [
    (32, 32, 4, 64, 64, 'relu', False),     # conv1 of the feature encoder
    (64, 64, 3, 64, 64, 'relu', False),     # conv3
    (64, 64, 2, 35, 37, 'none', True),      # conv4: residual, no activation, ragged tiles
    (64, 32, 5, 19, 8, 'none', False),      # conv5
    (32, 64, 1, 7, 5, 'relu', True),        # smaller than one tile
    (64, 64, 200, 16, 8, 'relu', False),    # more tiles than SMs x 2: several tiles per CTA, both accumulator buffers reused
])
def test_conv2d_halo_engine(knobs, halo, cin, cout, N, H, W, act, res):
    """Stride-1 3x3 2-D layers with 32 / 64 channels on the halo-once engine (csrc/conv2d_halo.cu) and, with knob
    no_conv2d_halo, on the engines it replaces: conv2d on bf16-rounded operands at the output's own storage rounding; the two
    engines agree to one bf16 rounding of the output; a launch into a channel / column slice of a wider buffer (the zero-margined
    feature rows of the cost-volume kernels) leaves the rest of the buffer untouched."""
    knobs('no_conv2d_halo', 1 - halo)
    torch.manual_seed(3)
    conv = _qmod(nn.Conv2d(cin, cout, 3, 1, 1), 'bf16')
    x_nc = _q(torch.randn(N, cin, H, W), 'bf16')
    r_nc = _q(torch.randn(N, cout, H, W), 'bf16') if res else None
    code = lib.ACT_RELU if act == 'relu' else lib.ACT_NONE
    pc = PackedConv.from_conv(conv, None, code, lib.DTYPE_BF16, 'cuda')
    with torch.no_grad():
        ref = conv(x_nc) + (r_nc if res else 0)
        ref = F.relu(ref) if act == 'relu' else ref
    x = to_cl(x_nc).to(torch.bfloat16).cuda()
    kw = {'residual': to_cl(r_nc).to(torch.bfloat16).cuda()} if res else {}
    n0 = lib.launches()
    got = pc(x, **kw)
    assert lib.launches() - n0 == 1
    refc = to_cl(ref)
    err = (got.float().cpu() - refc).abs().max().item()
    assert err <= TOL['bf16'] * (refc.abs().max().item() + 1e-6), err
    if not res:
        # into a slice of a wider, padded buffer: [N,1,H,W+6,cout] real pixels at columns 3 .. W+2 (as enc5 -> featp)
        P = W + 6
        buf = torch.full((N, 1, H, P, cout), 7.0, dtype=torch.bfloat16, device='cuda')
        pc(x, out=buf, out_view=(3 * cout, (H * P * cout, H * P * cout, P * cout, cout)), cout_store=cout)
        assert torch.equal(buf[:, :, :, 3:3 + W], got)
        assert (buf[:, :, :, :3] == 7).all() and (buf[:, :, :, 3 + W:] == 7).all()


@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
@pytest.mark.parametrize('cin,cout,D,H,W', [(64, 64, 8, 16, 16), (32, 32, 5, 9, 11), (9, 16, 8, 8, 8), (64, 1, 4, 8, 8)])
def test_conv3d(prec, cin, cout, D, H, W):
    torch.manual_seed(1)
    conv = _qmod(nn.Conv3d(cin, cout, 3, 1, 1, bias=False), prec)
    x = _q(torch.randn(2, cin, D, H, W), prec)
    pc = PackedConv.from_conv(conv, None, lib.ACT_LEAKY, _code(prec), 'cuda', act_param=0.2)
    with torch.no_grad():
        _check(pc, x, F.leaky_relu(conv(x), 0.2), prec, cout)


@pytest.mark.parametrize('ts1', [0, 1])
@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
@pytest.mark.parametrize('cin,cout,S,N', [(64, 32, 2, 5), (32, 8, 8, 2), (128, 64, 4, 3), (512, 128, 4, 2)])
def test_deconv(knobs, ts1, prec, cin, cout, S, N):
    knobs('igemm_ts1', ts1)
    torch.manual_seed(2)
    dc = _qmod(nn.ConvTranspose3d(cin, cout, 4, 2, 1, bias=False), prec)
    x = _q(torch.randn(N, cin, S, S, S), prec)
    pc = PackedConv.from_deconv_k4s2p1(dc, None, lib.ACT_RELU, _code(prec), 'cuda')
    with torch.no_grad():
        _check(pc, x, F.relu(dc(x)), prec, cout)


@pytest.mark.parametrize('prec', ['bf16', 'tf32'])
@pytest.mark.parametrize('shape', [(2, 5, 6, 7), (3, 16, 16, 16), (1, 1, 1, 1)])
def test_deconv_blocked_plus_depth_to_space(prec, shape):
    """ConvTranspose3d(32, 8, 4, 2, 1) + BN + ReLU as ONE 3x3x3 conv with 8 parity classes x 8 channels on the
    plane-scatter kernel, then depth-to-space with the 1x1x1 + sigmoid projection in channel 8."""
    from stereo_3d_reconstruction_b200 import ops
    torch.manual_seed(5)
    N, d, h, w = shape
    dc = nn.ConvTranspose3d(32, 8, 4, 2, 1, bias=True)
    bn0 = nn.BatchNorm3d(8).eval()
    with torch.no_grad():
        bn0.running_mean.normal_(0, 0.1);  bn0.running_var.uniform_(0.5, 1.5);  bn0.weight.uniform_(0.8, 1.2);  bn0.bias.normal_(0, 0.1)
        # fold the BN by hand and round the FOLDED weights (what the engine multiplies with); BN folding itself is
        # covered on CPU (tests/test_packing.py) and by the model tests
        scale = bn0.weight / torch.sqrt(bn0.running_var + bn0.eps)
        dc.weight.copy_(_q(dc.weight * scale.view(1, -1, 1, 1, 1), prec))
        dc.bias.copy_(bn0.bias - bn0.running_mean * scale)
    bn = None
    x = _q(torch.randn(N, 32, d, h, w), prec)
    with torch.no_grad():
        ref = to_cl(F.relu(dc(x)))                                    # [N,2d,2h,2w,8]
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
    conv = _qmod(nn.Conv2d(32, 32, 3, 1, 1), prec)
    x = _q(torch.randn(2, 32, 12, 12), prec)
    dt = torch.bfloat16 if prec == 'bf16' else torch.float32
    res = torch.randn(2, 32, 12, 12).to(dt).float()
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, _code(prec), 'cuda')
    with torch.no_grad():
        _check(pc, x, conv(x) + res, prec, 32, residual=to_cl(res).to(dt).cuda())
    fc = _qmod(nn.Linear(32 * 4 * 4, 100), prec)
    xf = _q(torch.randn(7, 32, 4, 4), prec)
    pc = PackedConv.from_linear_over_map(fc, 32, 4, 4, lib.ACT_TANH, _code(prec), 'cuda', act_param=0.5)
    with torch.no_grad():
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


@pytest.mark.parametrize('knob', [None, 'scatter_no_pair', 'scatter_generic', 'scatter_no_transpose',
                                  'scatter_res_transpose', 'scatter_tps3', 'no_scatter', 'scatter_no_rm', 'scatter_one_cta'])
@pytest.mark.parametrize('prec,N,cin,cout,D,H,W,res,act', [
    ('bf16', 1, 64, 64, 1, 8, 8, False, 'relu'),        # one column, one plane (no CTA pair possible)
    ('bf16', 3, 64, 64, 9, 40, 20, True, 'relu'),       # residual by identity-tap MMA (conv_scatter_rm.cu) + ReLU, ragged patches
    ('bf16', 40, 64, 64, 4, 32, 16, True, 'none'),      # 160 columns on 148 SMs: two columns per CTA, a phantom column
    ('bf16', 1, 64, 64, 2, 33, 9, True, 'none'),        # ragged patches in y and x, residual, 2 planes
    ('bf16', 3, 64, 48, 3, 40, 20, False, 'relu'),      # 48 channels: 6 chunks per pixel, non-transposed epilogue
    ('bf16', 2, 16, 16, 7, 32, 32, False, 'leaky'),     # fusion-scorer shape (32-byte rows, 9-tap stages)
    ('bf16', 40, 16, 16, 5, 32, 32, False, 'leaky'),    # the same with 160 columns: two CTAs per SM (kTwo), phantom columns
    ('bf16', 1, 32, 32, 4, 70, 70, True, 'relu'),       # 64-byte rows, odd number of columns
    ('bf16', 5, 64, 64, 6, 16, 24, True, 'none'),       # the residual layer's kernel
    ('bf16', 2, 64, 32, 5, 24, 16, False, 'none'),
    ('tf32', 2, 32, 32, 5, 20, 12, True, 'relu'),       # fp32 storage, 128-byte rows
    ('tf32', 1, 16, 64, 3, 9, 33, False, 'leaky'),
    ('tf32', 2, 64, 64, 5, 20, 12, True, 'relu'),       # 256-byte rows: two K chunks per tap (the tf32 aggregation layers)
    ('bf16', 1, 128, 64, 3, 16, 16, False, 'relu'),     # 128 bf16 channels: two K chunks as well
])
def test_conv3d_plane_scatter_shapes(knobs, knob, prec, N, cin, cout, D, H, W, res, act):
    """conv_scatter.cu over edge shapes, in every mode its knobs select (CTA pairs / single CTA, lean per-shape kernels /
    all-in-one kernel, transposed / direct stores and residual reads, 9- / 3-tap stages) and through the generic per-tap
    engine (no_scatter)."""
    if knob:
        knobs(knob, 1)
    torch.manual_seed(7)
    conv = _qmod(nn.Conv3d(cin, cout, 3, 1, 1, bias=True), prec)
    x = _q(torch.randn(N, cin, D, H, W), prec)
    dt = torch.bfloat16 if prec == 'bf16' else torch.float32
    code = {'relu': lib.ACT_RELU, 'none': lib.ACT_NONE, 'leaky': lib.ACT_LEAKY}[act]
    fn = {'relu': F.relu, 'none': lambda t: t, 'leaky': lambda t: F.leaky_relu(t, 0.2)}[act]
    pc = PackedConv.from_conv(conv, None, code, _code(prec), 'cuda', act_param=0.2)
    kw = {}
    with torch.no_grad():
        ref = conv(x)
    if res:
        r = torch.randn(N, cout, D, H, W).to(dt).float()
        ref = ref + r
        kw['residual'] = pad_c(to_cl(r), pc.cout_pad).to(dt).cuda()
    _check(pc, x, fn(ref), prec, cout, **kw)


@pytest.mark.parametrize('knob', [None, 'scatter_no_pair'])
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
def test_conv_concat_volume_fused_is_bit_identical(knobs, knob, prec, B, C, cout, D, h, w):
    """conv_scatter_concat.cu: cost volume (never written) + 3x3x3 conv == s3d_cost_volume_concat + the same conv, bit for
    bit (same MMA sequence), and == the oracle's volume + conv3d within the engine's tolerance."""
    from stereo_3d_reconstruction_b200 import ops
    from oracle import models as O
    dt = torch.bfloat16 if prec == 'bf16' else torch.float32
    if knob:
        if C * (2 if prec == 'bf16' else 4) == 128:
            pytest.skip('128-byte halves need CTA pairs (a single CTA cannot hold two weight stages of 36-48 KB)')
        knobs(knob, 1)
    torch.manual_seed(9)
    conv = _qmod(nn.Conv3d(2 * C, cout, 3, 1, 1, bias=True), prec)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, _code(prec), 'cuda')
    f = _q(torch.randn(2 * B, C, h, w), prec)
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
    with torch.no_grad():
        ref = to_cl(F.relu(conv(ref_vol)))
    err = (got.float().cpu()[..., :cout] - ref).abs().max().item()
    assert err <= TOL[prec] * (ref.abs().max().item() + 1e-6)


@pytest.mark.parametrize('B,C,D,h,w', [
    (2, 32, 8, 16, 16),        # the network's shape (64-byte halves)
    (3, 32, 32, 40, 64),       # D = 32 planes, two patch rows, several columns per CTA
    (1, 16, 5, 9, 13),         # ragged patches, 32-byte halves
    (2, 32, 1, 8, 8),          # one plane: both border corrections land on it
    (2, 32, 2, 8, 8),          # two planes: out[0] and out[D-1] are neighbours
    (1, 32, 12, 8, 8),         # more disparities than pixels in a row
    (80, 32, 4, 8, 8),         # 160 columns on 148 SMs: some CTAs walk two columns (R rewritten per column), phantom column
])
def test_conv_concat_volume_ref_once(B, C, D, h, w):
    """Reference-once form of the fused cost volume + first aggregation layer (s3d_conv_concat_volume_ro): same result as the
    volume + conv3d on bf16-rounded operands within the bf16 engine tolerance, and as the bit-exact fused kernel to the
    same bound (the summed reference weights are rounded to bf16 once: not bit-identical by design)."""
    from stereo_3d_reconstruction_b200 import ops
    from oracle import models as O
    cout = 64
    torch.manual_seed(11)
    conv = _qmod(nn.Conv3d(2 * C, cout, 3, 1, 1, bias=True), 'bf16')
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, _code('bf16'), 'cuda')
    f = _q(torch.randn(2 * B, C, h, w), 'bf16')
    feat = to_cl(f).to(torch.bfloat16).cuda()
    pad, P = D, w + 2 * D
    featp = torch.zeros(2 * B, 1, h, P, C, dtype=torch.bfloat16, device='cuda')
    featp[:, :, :, pad:pad + w] = feat
    got = ops.conv_concat_volume(pc, featp, B, D, pad, ref_once=True)
    exact = ops.conv_concat_volume(pc, featp, B, D, pad)
    torch.cuda.synchronize()
    assert got.shape == exact.shape
    scale = exact.float().abs().max().item() + 1e-6
    # two bf16 roundings of the output apart at most, plus the rounding of the summed weights
    assert (got.float() - exact.float()).abs().max().item() <= 2 * TOL['bf16'] * scale
    fr = feat.float().cpu()[:, 0].permute(0, 3, 1, 2)
    ref_vol = torch.cat([O.build_concat_volume(fr[:B], fr[B:], D, -1), O.build_concat_volume(fr[B:], fr[:B], D, +1)], 0)
    with torch.no_grad():
        ref = to_cl(F.relu(conv(ref_vol)))
    err = (got.float().cpu()[..., :cout] - ref).abs().max().item()
    # the summed reference weights are rounded to bf16 once more than in the reference: allow twice the engine's bound
    assert err <= 2 * TOL['bf16'] * (ref.abs().max().item() + 1e-6)
    # every plane on its own (a wrong border correction would hide behind the max over the whole volume otherwise)
    for z in sorted({0, 1, D // 2, D - 2, D - 1} & set(range(D))):
        ez = (got.float().cpu()[:, z, ..., :cout] - ref[:, z]).abs().max().item()
        assert ez <= 2 * TOL['bf16'] * (ref[:, z].abs().max().item() + 1e-6), (z, ez)


@pytest.mark.parametrize('B,C,D,h,w', [
    (1, 32, 32, 64, 64),       # one stereo pair at the benchmark shape
    (2, 32, 8, 16, 24),
    (1, 16, 5, 9, 13),         # ragged row blocks, 16 feature channels
    (2, 32, 2, 8, 8),          # two planes: both border corrections on neighbouring planes
    (1, 32, 12, 8, 8),         # more disparities than pixels in a row: most target reads fall into the zero margin
    (3, 32, 16, 33, 40),       # rows longer than one 32-pixel block, odd height
])
@pytest.mark.parametrize('map_engine', ['map_conv', 'generic'])
def test_conv_concat_volume_sheared(B, C, D, h, w, map_engine, knobs):
    """SHEARED form of the cost volume + first aggregation layer (ops.conv_concat_volume_sheared: 2-D map convolutions on the
    generic engine + s3d_concat_gonce_assemble): same result as conv3d over the oracle's concat volume on bf16-rounded
    operands, and as the bit-exact fused kernel, within the bound of the reference-once form (weights summed before the bf16
    rounding); every border plane and both edge columns checked on their own."""
    from stereo_3d_reconstruction_b200 import ops
    from oracle import models as O
    knobs('no_map_conv', int(map_engine == 'generic'))      # maps by csrc/map_conv.cu (C = 32) or by the generic engine
    cout = 64
    torch.manual_seed(12)
    conv = _qmod(nn.Conv3d(2 * C, cout, 3, 1, 1, bias=True), 'bf16')
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, _code('bf16'), 'cuda')
    f = _q(torch.randn(2 * B, C, h, w), 'bf16')
    feat = to_cl(f).to(torch.bfloat16).cuda()
    pad, P = max(D, 4), w + 2 * max(D, 4)
    featp = torch.zeros(2 * B, 1, h, P, C, dtype=torch.bfloat16, device='cuda')
    featp[:, :, :, pad:pad + w] = feat
    got = ops.conv_concat_volume_sheared(pc, featp, B, D, pad)
    exact = ops.conv_concat_volume(pc, featp, B, D, pad)
    torch.cuda.synchronize()
    assert got.shape == exact.shape
    scale = exact.float().abs().max().item() + 1e-6
    assert (got.float() - exact.float()).abs().max().item() <= 2 * TOL['bf16'] * scale
    fr = feat.float().cpu()[:, 0].permute(0, 3, 1, 2)
    ref_vol = torch.cat([O.build_concat_volume(fr[:B], fr[B:], D, -1), O.build_concat_volume(fr[B:], fr[:B], D, +1)], 0)
    with torch.no_grad():
        ref = to_cl(F.relu(conv(ref_vol)))
    g = got.float().cpu()[..., :cout]
    assert (g - ref).abs().max().item() <= 2 * TOL['bf16'] * (ref.abs().max().item() + 1e-6)
    for z in sorted({0, 1, D // 2, D - 2, D - 1} & set(range(D))):
        ez = (g[:, z] - ref[:, z]).abs().max().item()
        assert ez <= 2 * TOL['bf16'] * (ref[:, z].abs().max().item() + 1e-6), (z, ez)
    for x in (0, 1, w - 2, w - 1):                      # the edge-column maps (x = w-1 left-referenced, x = 0 right-referenced)
        ex = (g[:, :, :, x] - ref[:, :, :, x]).abs().max().item()
        assert ex <= 2 * TOL['bf16'] * (ref[:, :, :, x].abs().max().item() + 1e-6), (x, ex)


@pytest.mark.parametrize('nz', [0, 2, 3, 5, 8])
@pytest.mark.parametrize('prec,N,cin,cout,D,H,W,res', [
    ('bf16', 1, 64, 64, 32, 64, 64, False),      # the aggregation layer at batch 1: 32 columns -> 4 chunks of 8 planes
    ('bf16', 2, 64, 64, 13, 33, 9, True),        # ragged last chunk, residual
    ('bf16', 2, 16, 16, 32, 32, 32, False),      # fusion scorer at batch 1
    ('tf32', 1, 32, 32, 9, 20, 12, True),
])
def test_plane_scatter_z_split(knobs, nz, prec, N, cin, cout, D, H, W, res):
    """Small batches: a column is cut into z-chunks that march over 2 extra planes and store only their own planes
    (conv_scatter.cuh, ScArgs::nz).  nz = 0: the launcher's own choice; the result must equal the unsplit kernel's BIT FOR BIT
    (same MMA sequence per output plane)."""
    torch.manual_seed(11)
    conv = _qmod(nn.Conv3d(cin, cout, 3, 1, 1, bias=True), prec)
    x = _q(torch.randn(N, cin, D, H, W), prec)
    dt = torch.bfloat16 if prec == 'bf16' else torch.float32
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, _code(prec), 'cuda')
    kw = {}
    with torch.no_grad():
        ref = conv(x)
    if res:
        r = torch.randn(N, cout, D, H, W).to(dt).float()
        ref = ref + r
        kw['residual'] = pad_c(to_cl(r), pc.cout_pad).to(dt).cuda()
    xc = pad_c(to_cl(x), pc.cin_pad).to(dt).cuda()
    knobs('scatter_zsplit', -1)
    whole = pc(xc, engine='igemm', **kw).clone()
    knobs('scatter_zsplit', nz)
    got = pc(xc, engine='igemm', **kw)
    assert torch.equal(got, whole)
    _check(pc, x, F.relu(ref), prec, cout, **kw)
