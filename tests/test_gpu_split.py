"""'bf16x3' precision (S3D_DTYPE_BF16X2): every value is a bf16 pair hi + lo and every tensor-core product is
hi*hi + lo*hi + hi*lo on kind::f16 MMAs into one fp32 accumulator.  Each kernel of the split path is compared with the
torch fp32 op on UNROUNDED fp32 inputs; the stated tolerance is 1e-4 of the output's max (expected ~2^-16 per product)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import models as O
from stereo_3d_reconstruction_b200 import lib, models as M, ops
from stereo_3d_reconstruction_b200.layers import PackedConv, to_storage, from_storage
from stereo_3d_reconstruction_b200.utils import synthetic
from tests.common import small_cfg
from tests.emulate import to_cl, pad_c

pytestmark = pytest.mark.gpu
X2 = lib.DTYPE_BF16X2
TOL = 1e-4


def _split(x_cl):                 # fp32 [..., C] -> bf16 [..., hi(C) | lo(C)]
    return to_storage(x_cl, X2)


def _unsplit(t):
    return from_storage(t.cpu(), X2)


def _check(pc, x_nc, ref_nc, cout, **kw):
    x = _split(pad_c(to_cl(x_nc), pc.cin_pad)).cuda()
    got = pc(x, engine='igemm', **kw)
    assert got.dtype == torch.bfloat16 and got.shape[-1] == 2 * pc.cout_pad
    got = _unsplit(got)
    ref = to_cl(ref_nc)
    assert got.shape[:4] == ref.shape[:4]
    scale = ref.abs().max().item() + 1e-6
    err = (got[..., :cout] - ref).abs().max().item()
    assert err <= TOL * scale, (err, scale)
    if got.shape[-1] > cout:
        assert got[..., cout:].abs().max().item() == 0.0


def test_split_roundtrip_and_storage():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(3, 5, 7, 24, generator=g) * 3
    s = ops.split_bf16(x.cuda())
    assert torch.equal(s.cpu(), _split(x))                                     # kernel == the host packing rule
    back = ops.unsplit_bf16(s).cpu()
    assert (back - x).abs().max().item() <= 2.0 ** -16 * x.abs().max().item()


@pytest.mark.parametrize('cin,cout,H,W,stride', [(32, 64, 35, 35, 2), (16, 32, 37, 41, 2), (128, 256, 9, 9, 2), (256, 320, 8, 8, 1)])
def test_generic_engine_conv2d(cin, cout, H, W, stride):
    torch.manual_seed(0)
    conv = nn.Conv2d(cin, cout, 3, stride, 1)
    x = torch.randn(3, cin, H, W)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, X2, 'cuda')
    with torch.no_grad():
        _check(pc, x, F.relu(conv(x)), cout)


@pytest.mark.parametrize('cin,cout,S,N', [(64, 32, 2, 5), (128, 64, 4, 3), (2048, 512, 2, 2)])
def test_generic_engine_deconv_and_linear(cin, cout, S, N):
    torch.manual_seed(2)
    dc = nn.ConvTranspose3d(cin, cout, 4, 2, 1, bias=False)
    x = torch.randn(N, cin, S, S, S)
    pc = PackedConv.from_deconv_k4s2p1(dc, None, lib.ACT_RELU, X2, 'cuda')
    with torch.no_grad():
        _check(pc, x, F.relu(dc(x)), cout)
    fc = nn.Linear(32 * 4 * 4, 100)
    xf = torch.randn(7, 32, 4, 4)
    pc = PackedConv.from_linear_over_map(fc, 32, 4, 4, lib.ACT_TANH, X2, 'cuda', act_param=0.5)
    with torch.no_grad():
        ref = (0.5 * torch.tanh(fc(xf.flatten(1)))).view(7, 100, 1, 1)
    _check(pc, xf, ref, 100)
    # split in -> fp32 out (the point decoder's last layer, the classifier's tap projections)
    x = _split(pad_c(to_cl(xf), pc.cin_pad)).cuda()
    out32 = torch.empty(7, 1, 1, 1, pc.cout_pad, dtype=torch.float32, device='cuda')
    pc(x, out=out32)
    assert (out32.cpu()[..., :100] - to_cl(ref)).abs().max().item() <= TOL


@pytest.mark.parametrize('knob', [None, 'scatter_no_pair', 'scatter_generic', 'no_scatter'])
@pytest.mark.parametrize('N,cin,cout,D,H,W,res,act', [
    (1, 64, 64, 1, 8, 8, False, 'relu'),        # one column, one plane
    (2, 64, 64, 5, 33, 9, True, 'none'),        # the aggregation shapes: 256-byte [hi | lo] rows = two K chunks; residual
    (2, 64, 64, 6, 16, 24, False, 'relu'),
    (3, 64, 32, 3, 40, 20, False, 'none'),      # enc5's shape
    (2, 16, 16, 7, 32, 32, False, 'leaky'),     # fusion scorer: 64-byte rows, 9-tap stages
    (1, 32, 32, 4, 70, 70, True, 'relu'),       # 128-byte rows, one chunk
    (2, 32, 64, 5, 16, 16, False, 'relu'),      # the blocked last deconv's shape
    (2, 9, 16, 4, 12, 12, False, 'leaky'),      # padded input channels
])
def test_plane_scatter_split(knobs, knob, N, cin, cout, D, H, W, res, act):
    if knob:
        knobs(knob, 1)
    torch.manual_seed(7)
    conv = nn.Conv3d(cin, cout, 3, 1, 1, bias=True)
    x = torch.randn(N, cin, D, H, W)
    code = {'relu': lib.ACT_RELU, 'none': lib.ACT_NONE, 'leaky': lib.ACT_LEAKY}[act]
    fn = {'relu': F.relu, 'none': lambda t: t, 'leaky': lambda t: F.leaky_relu(t, 0.2)}[act]
    pc = PackedConv.from_conv(conv, None, code, X2, 'cuda', act_param=0.2)
    kw = {}
    with torch.no_grad():
        ref = conv(x)
    if res:
        r = torch.randn(N, cout, D, H, W)
        ref = ref + r
        kw['residual'] = _split(pad_c(to_cl(r), pc.cout_pad)).cuda()
    _check(pc, x, fn(ref), cout, **kw)


@pytest.mark.parametrize('nz', [0, 3])
def test_plane_scatter_split_z_split(knobs, nz):
    """bf16x3 kernels with z-chunks (small batches): bit-identical to the unsplit march."""
    torch.manual_seed(12)
    conv = nn.Conv3d(64, 64, 3, 1, 1, bias=True)
    x = torch.randn(1, 64, 16, 40, 24)
    r = torch.randn(1, 64, 16, 40, 24)
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, X2, 'cuda')
    xs, rs = _split(to_cl(x)).cuda(), _split(to_cl(r)).cuda()
    knobs('scatter_zsplit', -1)
    whole = pc(xs, residual=rs).clone()
    knobs('scatter_zsplit', nz)
    assert torch.equal(pc(xs, residual=rs), whole)
    with torch.no_grad():
        _check(pc, x, conv(x) + r, 64, residual=rs)


def test_conv2d_as_volume_split():
    """stride-1 3x3 2-D layers run as volumes of images on the plane-scatter kernel (encoder layers 1, 3, 4, 5)."""
    torch.manual_seed(8)
    conv = nn.Conv2d(64, 64, 3, 1, 1)
    x = torch.randn(6, 64, 20, 28)
    r = torch.randn(6, 64, 20, 28)
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, X2, 'cuda')
    assert pc.vol is not None
    with torch.no_grad():
        _check(pc, x, conv(x) + r, 64, residual=_split(to_cl(r)).cuda())


@pytest.mark.parametrize('with_disp,u8,H,W,cout', [(False, False, 16, 20, 32), (True, False, 9, 11, 16), (True, True, 7, 5, 32)])
def test_conv_first_split(with_disp, u8, H, W, cout):
    g = torch.Generator().manual_seed(21)
    cin = 4 if with_disp else 3
    conv = nn.Conv2d(cin, cout, 3, 2, 1)
    if u8:
        raw = torch.randint(0, 256, (2, H, W, 3), generator=g, dtype=torch.uint8)
        img = (raw.float() * (1.0 / 255.0)).permute(0, 3, 1, 2).contiguous()
    else:
        raw = img = torch.rand(2, 3, H, W, generator=g)
    disp = torch.rand(2, H, W, generator=g) * 20 if with_disp else None
    x = img if disp is None else torch.cat([img, (disp * 0.05).unsqueeze(1)], 1)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, X2, 'cuda')
    with torch.no_grad():
        ref = F.relu(conv(x))
    got = ops.conv_first(raw.cuda(), pc, None if disp is None else disp.cuda(), 0.05)
    assert got.shape == (2, 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1, 2 * pc.cout_pad)
    err = (_unsplit(got)[:, 0, :, :, :cout] - ref.permute(0, 2, 3, 1)).abs().max().item()
    assert err <= TOL * ref.abs().max().item()


def test_glue_kernels_split():
    g = torch.Generator().manual_seed(6)
    # concat volume: the hi halves and the lo halves each form the plain volume
    B, C, h, w, D = 2, 16, 9, 21, 8
    f = torch.randn(2 * B, C, h, w, generator=g)
    fs = _split(to_cl(f))
    vol = ops.cost_volume_concat(fs.cuda(), B, D, split=True).cpu()
    assert vol.shape == (2 * B, D, h, w, 4 * C)
    hi = ops.cost_volume_concat(fs[..., :C].contiguous().cuda(), B, D).cpu()
    lo = ops.cost_volume_concat(fs[..., C:].contiguous().cuda(), B, D).cpu()
    assert torch.equal(vol[..., :2 * C], hi) and torch.equal(vol[..., 2 * C:], lo)
    fr = _unsplit(fs)[:, 0].permute(0, 3, 1, 2)
    ref = torch.cat([O.build_concat_volume(fr[:B], fr[B:], D, -1), O.build_concat_volume(fr[B:], fr[:B], D, +1)], 0)
    assert torch.equal(_unsplit(vol), ref.permute(0, 2, 3, 4, 1).contiguous())
    # pooling / latent re-indexing
    N, C, H, W, L = 3, 16, 5, 7, 2
    x = torch.randn(N, C, H, W, generator=g)
    pooled = F.adaptive_avg_pool2d(x, L)
    ref_vox = pooled.reshape(N, C * L * L // 8, 2, 2, 2).permute(0, 2, 3, 4, 1).contiguous()
    xs = _split(to_cl(x)).cuda()
    torch.testing.assert_close(_unsplit(ops.latent_to_vox(xs, L, split=True)), ref_vox, rtol=3e-5, atol=3e-5)
    torch.testing.assert_close(_unsplit(ops.avg_pool(xs, L, split=True)), to_cl(pooled), rtol=3e-5, atol=3e-5)
    # depth-to-space + projection
    N, d = 2, 3
    y = torch.randn(N, d, d, d, 64, generator=g)
    pw = torch.randn(8, generator=g) * 0.5
    out = _unsplit(ops.depth_to_space(_split(y).cuda(), 16, pw.cuda(), lib.ACT_SIGMOID, split=True))
    assert out.shape == (N, 2 * d, 2 * d, 2 * d, 16)
    yv = _unsplit(_split(y)).view(N, d, d, d, 2, 2, 2, 8)
    ref = yv.permute(0, 1, 4, 2, 5, 3, 6, 7).reshape(N, 2 * d, 2 * d, 2 * d, 8)
    assert torch.equal(out[..., :8], ref)
    torch.testing.assert_close(out[..., 8], torch.sigmoid((ref * pw).sum(-1)), rtol=3e-5, atol=3e-5)
    assert out[..., 9:].abs().max().item() == 0
    # fusion + IoU on split tensors
    B, V, nv = 2, 2, 32 ** 3
    score = torch.randn(V * B, nv, 16, generator=g)
    volm = torch.rand(V * B, nv, 16, generator=g)
    gt = synthetic.gt_volume(B, seed=7)
    th = [0.2, 0.3, 0.4, 0.5]
    ss, vs = _split(score), _split(volm)
    s = _unsplit(ss)[..., 0].view(V, B, nv)
    v = _unsplit(vs)[..., 8].view(V, B, nv)
    ref = torch.clamp((F.softmax(s, 0) * v).sum(0), 0, 1)
    iou = torch.zeros(B, 4, 2, dtype=torch.int64).cuda()
    got = ops.fuse_views(ss.cuda(), 0, 32, vs.cuda(), 8, 32, B, V, nv, gt=gt.view(B, -1).cuda(), thresholds=th, iou=iou,
                         score_lo=16, vol_lo=16).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    assert torch.equal(iou.cpu(), O.iou_counts(got.view(B, 32, 32, 32), gt, th))


@pytest.mark.parametrize('cv', ['concat', 'corr'])
@pytest.mark.parametrize('name', ['Stereo2Voxel', 'Stereo2Point'])
def test_small_model_bf16x3_matches_oracle(name, cv):
    """The whole forward in 'bf16x3' against the CPU fp32 oracle: the north_star's 1e-3 (stated here: 5e-4)."""
    cfg = small_cfg(NETWORK__PRECISION='bf16x3', NETWORK__COST_VOLUME=cv)
    oracle = O.make_model(name, cfg, seed=0)
    model = M.build_model(name, cfg)
    model.load_state_dict(oracle.state_dict())
    model.cuda().pack()
    left, right, _ = synthetic.stereo_pair(3, 64, 64, 16, seed=0)
    with torch.no_grad():
        r = oracle(left, right)
        o = model(left.cuda(), right.cuda())
    dmax = r[0].abs().max().item()
    assert (o[0].cpu() - r[0]).abs().max().item() <= 5e-4 * dmax
    assert (o[1].cpu() - r[1]).abs().max().item() <= 5e-4 * dmax
    assert (o[2].cpu() - r[2]).abs().max().item() <= 5e-4


@pytest.mark.parametrize('H,W,B', [(70, 50, 2), (37, 101, 1)])
def test_odd_input_sizes_bf16x3(H, W, B):
    cfg = small_cfg(NETWORK__PRECISION='bf16x3', CONST__IMG_H=H, CONST__IMG_W=W)
    oracle = O.make_model('Stereo2Voxel', cfg, seed=1)
    model = M.build_model('Stereo2Voxel', cfg)
    model.load_state_dict(oracle.state_dict())
    model.cuda().pack()
    left, right, _ = synthetic.stereo_pair(B, H, W, 16, seed=0)
    gt = synthetic.gt_volume(B)
    with torch.no_grad():
        rdl, rdr, rvox = oracle(left, right)
        dl, dr, vox, iou = model(left.cuda(), right.cuda(), gt.cuda())
    assert (dl.cpu() - rdl).abs().max().item() <= 5e-4 * rdl.abs().max().item()
    assert (vox.cpu() - rvox).abs().max().item() <= 5e-4
    assert torch.equal(iou.cpu(), O.iou_counts(vox.cpu(), gt, cfg.TEST.VOXEL_THRESH))


@pytest.mark.parametrize('B,D,h,w', [
    (2, 8, 16, 16),        # the network's shape, 8 columns
    (3, 32, 40, 64),       # D = 32 planes, two patch rows, 96 columns
    (1, 5, 9, 13),         # ragged patches
    (2, 1, 8, 8),          # one plane: both border corrections land on it
    (2, 2, 8, 8),          # two planes
    (80, 4, 8, 8),         # 160 columns on 148 SMs: two columns per CTA, phantom column
])
def test_conv_concat_volume_ref_once_split(B, D, h, w):
    """Reference-once cost volume + first aggregation layer on split (BF16X2) operands (conv_scatter_concat_ros_kernel) against
    Conv3d + ReLU over the oracle's concat volume on UNROUNDED fp32 features, plane by plane: 1e-4 of the output's max."""
    C, cout = 32, 64
    torch.manual_seed(21)
    conv = nn.Conv3d(2 * C, cout, 3, 1, 1, bias=True)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, X2, 'cuda')
    f = torch.randn(2 * B, C, h, w)
    feat = _split(to_cl(f)).cuda()                                      # [2B,1,h,w, hi(C) | lo(C)]
    pad, P = D, w + 2 * D
    featp = torch.zeros(2 * B, 1, h, P, 2 * C, dtype=torch.bfloat16, device='cuda')
    featp[:, :, :, pad:pad + w] = feat
    got = ops.conv_concat_volume(pc, featp, B, D, pad, ref_once=True)
    torch.cuda.synchronize()
    assert got.shape == (2 * B, D, h, w, 2 * cout)
    got = _unsplit(got)
    ref_vol = torch.cat([O.build_concat_volume(f[:B], f[B:], D, -1), O.build_concat_volume(f[B:], f[:B], D, +1)], 0)
    with torch.no_grad():
        ref = to_cl(F.relu(conv(ref_vol)))
    scale = ref.abs().max().item() + 1e-6
    for z in range(D):
        ez = (got[:, z] - ref[:, z]).abs().max().item()
        assert ez <= TOL * scale, (z, ez, scale)


@pytest.mark.parametrize('B,D,h,w', [(1, 8, 16, 24), (2, 4, 9, 13), (1, 32, 64, 64), (2, 2, 8, 8)])
def test_conv_concat_volume_sheared_split(B, D, h, w):
    """SHEARED form of the cost volume + first aggregation layer on split (BF16X2) operands: map convolutions with three MMAs per
    product (map_conv_kernel<.., kSplit>) + the streaming pass writing bf16 pairs, against Conv3d + ReLU over the oracle's concat
    volume on UNROUNDED fp32 features, plane by plane and at the edge columns: 1e-4 of the output's max."""
    C, cout = 32, 64
    torch.manual_seed(22)
    conv = nn.Conv3d(2 * C, cout, 3, 1, 1, bias=True)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, X2, 'cuda')
    f = torch.randn(2 * B, C, h, w)
    feat = _split(to_cl(f)).cuda()                                      # [2B,1,h,w, hi(C) | lo(C)]
    pad = max(D, 4)
    P = w + 2 * pad
    featp = torch.zeros(2 * B, 1, h, P, 2 * C, dtype=torch.bfloat16, device='cuda')
    featp[:, :, :, pad:pad + w] = feat
    got = ops.conv_concat_volume_sheared(pc, featp, B, D, pad)
    torch.cuda.synchronize()
    assert got.shape == (2 * B, D, h, w, 2 * cout)
    got = _unsplit(got)
    ref_vol = torch.cat([O.build_concat_volume(f[:B], f[B:], D, -1), O.build_concat_volume(f[B:], f[:B], D, +1)], 0)
    with torch.no_grad():
        ref = to_cl(F.relu(conv(ref_vol)))
    scale = ref.abs().max().item() + 1e-6
    for z in range(D):
        ez = (got[:, z] - ref[:, z]).abs().max().item()
        assert ez <= TOL * scale, (z, ez, scale)
    for x in (0, w - 1):
        ex = (got[:, :, :, x] - ref[:, :, :, x]).abs().max().item()
        assert ex <= TOL * scale, (x, ex, scale)
