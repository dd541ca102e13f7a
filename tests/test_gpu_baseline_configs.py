"""Parity AT the BASELINE.json configurations themselves (not a reduced network):

  (a) the default config.py Stereo2Voxel (256x256, D=32, C=32, A=64, 2048->512->... decoder, 73 M parameters;
      BASELINE configs[0..2], `runner.py --test`, /root/reference/README.md:91 and :68-78) against the CPU oracle in
      every precision mode, tolerances stated here;
  (b) the default Stereo2Point (N_POINTS = 2048; configs[3]);
  (c) Chamfer at 32 x 2048 x 16384 (configs[3]) bit-exact in indices AND distances against oracle/chamfer_ref.c,
      duplicated blocks included (ties -> lowest index);
  (d) the dominant kernels at the exact bench shape (64->64 3x3x3 over D=32, 64x64 planes, residual and not; the
      fused classifier + soft-argmin) against F.conv3d on operand-rounded inputs;
  (e) the corners of the cost-volume sweep (configs[4]: D=128, C=64, 64x64).

The oracle is this repo's restatement of the north_star (PARITY UNPINNED: the reference's model source is not on disk).
The CPU oracle takes ~0.25 s per pair at the default size; the whole file runs in about a minute."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from config import cfg as default_cfg
from oracle import chamfer as OC
from oracle import models as O
from stereo_3d_reconstruction_b200 import lib, models as M, ops
from stereo_3d_reconstruction_b200.layers import PackedConv
from stereo_3d_reconstruction_b200.utils import synthetic
from tests.emulate import to_cl

pytestmark = pytest.mark.gpu

# (disparity relative to the oracle's max disparity, occupancy absolute).  The north_star asks for 1e-3 in fp32: the
# exact SIMT mode ('fp32'), the three-pass split modes on the tensor cores ('bf16x3': bf16 hi/lo operand pairs on
# kind::f16; 'tf32x3': three kind::tf32 passes) are held to it.  Single-pass 'tf32' and 'bf16' state their own wider
# tolerances (10-bit / 8-bit operands through 12 stacked convs).
TOLS = {'fp32': (1e-3, 1e-3), 'bf16x3': (1e-3, 1e-3), 'tf32x3': (1e-3, 1e-3), 'tf32': (1e-2, 2e-2), 'bf16': (3e-2, 6e-2)}
HAVE = getattr(M, 'PRECISIONS', ('bf16', 'tf32', 'tf32x3', 'fp32'))
MEASURED = {}           # printed at the end of the module (pytest -s) and usable by scripts/parity_report.py


# B = 2: 64 columns per aggregation layer, fewer than half the SMs -> the small-batch paths (materialised volume, z-split columns,
# separate classifier kernel).  B = 3: 96 columns -> the paths of the benchmark (reference-once first layer, residual on the tensor
# core, last layer + classifier chain; for 'bf16x3' the split reference-once kernel).
@pytest.fixture(scope='module', params=[2, 3], ids=['B2-small-batch-paths', 'B3-batched-paths'])
def voxel_case(request):
    cfg = default_cfg.clone()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    oracle = O.make_model('Stereo2Voxel', cfg, seed=0)
    B = request.param
    left, right, _ = synthetic.stereo_pair(B, cfg.CONST.IMG_H, cfg.CONST.IMG_W, 2 * cfg.NETWORK.MAX_DISP, seed=11)
    gt = synthetic.gt_volume(B, seed=12)
    with torch.no_grad():
        rdl, rdr, rvox = oracle(left, right)
    assert sum(p.numel() for p in oracle.parameters()) > 70e6                  # the 73 M-parameter network, not a toy
    return cfg, oracle.state_dict(), left, right, gt, rdl, rdr, rvox


@pytest.mark.parametrize('prec', ['fp32', 'bf16x3', 'tf32x3', 'tf32', 'bf16'])
def test_default_stereo2voxel_matches_oracle(voxel_case, prec):
    if prec not in HAVE:
        pytest.skip('precision mode %s not built' % prec)
    cfg0, sd, left, right, gt, rdl, rdr, rvox = voxel_case
    cfg = cfg0.clone()
    cfg.NETWORK.PRECISION = prec
    model = M.build_model('Stereo2Voxel', cfg)
    model.load_state_dict(sd)
    model.cuda().pack()
    with torch.no_grad():
        dl, dr, vox, iou = model(left.cuda(), right.cuda(), gt.cuda())
    dl, dr, vox, iou = dl.cpu(), dr.cpu(), vox.cpu(), iou.cpu()
    td, tv = TOLS[prec]
    dmax = max(rdl.abs().max().item(), rdr.abs().max().item())
    e_d = max((dl - rdl).abs().max().item(), (dr - rdr).abs().max().item()) / dmax
    e_v = (vox - rvox).abs().max().item()
    B = left.shape[0]
    MEASURED['stereo2voxel/%s/B%d' % (prec, B)] = (e_d, e_v, (vox - rvox).abs().mean().item())
    assert dl.shape == (B, 1, 256, 256) and vox.shape == (B, 32, 32, 32)
    assert rdl.std() > 1.0 and rvox.std() > 0.05                               # a non-degenerate target
    assert e_d <= td, ('disparity', prec, e_d)
    assert e_v <= tv, ('occupancy', prec, e_v)
    for t in cfg.TEST.VOXEL_THRESH:                                            # thresholded voxels: only boundary flips
        flips = (vox >= t) != (rvox >= t)
        assert ((rvox - t).abs()[flips] <= tv).all()
    assert torch.equal(iou, O.iou_counts(vox, gt, cfg.TEST.VOXEL_THRESH))      # integer stats exact


def test_default_stereo2voxel_reference_once_first_layer(voxel_case, knobs):
    """A/B path of the first aggregation layer (knob `no_sheared`: the reference-once tensor-core kernel instead of the sheared
    form -- 2-D map convolutions + one streaming pass -- that the bf16 forward takes by default) inside the default
    73 M-parameter network, held to the bf16 tolerances; the launch counts tell the two paths apart."""
    cfg0, sd, left, right, gt, rdl, rdr, rvox = voxel_case
    if left.shape[0] < 3:
        pytest.skip('batch 2 takes the small-batch paths (materialised volume)')
    cfg = cfg0.clone()
    cfg.NETWORK.PRECISION = 'bf16'
    from stereo_3d_reconstruction_b200 import lib as _l
    counts = {}
    for no_sheared in (0, 1):
        knobs('no_sheared', no_sheared)
        model = M.build_model('Stereo2Voxel', cfg)
        model.load_state_dict(sd)
        model.cuda().pack()
        n0 = _l.launches()
        with torch.no_grad():
            dl, dr, vox, iou = model(left.cuda(), right.cuda(), gt.cuda())
        counts[no_sheared] = _l.launches() - n0
        td, tv = TOLS['bf16']
        dmax = max(rdl.abs().max().item(), rdr.abs().max().item())
        e_d = max((dl.cpu() - rdl).abs().max().item(), (dr.cpu() - rdr).abs().max().item()) / dmax
        e_v = (vox.cpu() - rvox).abs().max().item()
        MEASURED['stereo2voxel/bf16%s/B%d' % ('+reference-once' if no_sheared else '', left.shape[0])] = (e_d, e_v, (vox.cpu() - rvox).abs().mean().item())
        assert e_d <= td and e_v <= tv, (no_sheared, e_d, e_v)
    assert counts[0] == counts[1] + 4, counts               # four map convolutions + the streaming pass replace one kernel


@pytest.mark.parametrize('prec', ['fp32', 'bf16x3', 'bf16'])
def test_default_stereo2point_matches_oracle(prec):
    if prec not in HAVE:
        pytest.skip('precision mode %s not built' % prec)
    cfg = default_cfg.clone()
    cfg.NETWORK.PRECISION = prec
    assert cfg.CONST.N_POINTS == 2048
    oracle = O.make_model('Stereo2Point', cfg, seed=0)
    model = M.build_model('Stereo2Point', cfg)
    model.load_state_dict(oracle.state_dict())
    model.cuda().pack()
    left, right, _ = synthetic.stereo_pair(2, cfg.CONST.IMG_H, cfg.CONST.IMG_W, 2 * cfg.NETWORK.MAX_DISP, seed=13)
    with torch.no_grad():
        rdl, _, rp = oracle(left, right)
        dl, _, p = model(left.cuda(), right.cuda())
    assert p.shape == rp.shape == (2, 2048, 3)
    e_p = (p.cpu() - rp).abs().max().item()                                    # points live in [-0.5, 0.5]^3
    e_d = (dl.cpu() - rdl).abs().max().item() / rdl.abs().max().item()
    MEASURED['stereo2point/' + prec] = (e_d, e_p)
    assert rp.std() > 0.05
    assert e_p <= {'fp32': 1e-3, 'bf16x3': 1e-3, 'bf16': 4e-2}[prec], ('points', prec, e_p)
    assert e_d <= TOLS[prec][0]


def test_chamfer_config4_bit_exact():
    """32 x 2048 predicted vs 16384 GT points: indices and distances array_equal with the C oracle."""
    a, b = synthetic.point_clouds(32, 2048, 16384, seed=21, duplicates=True)
    rd1, rd2, ri1, ri2 = OC.chamfer_c(a.numpy(), b.numpy())
    d1, d2, i1, i2 = ops.chamfer_forward(a.cuda(), b.cuda())
    assert np.array_equal(i1.cpu().numpy(), ri1) and np.array_equal(i2.cpu().numpy(), ri2)
    assert np.array_equal(d1.cpu().numpy(), rd1) and np.array_equal(d2.cpu().numpy(), rd2)
    # the duplicated blocks did produce exact ties that resolved to the lower index
    k = 16384 // 8
    assert (ri1 < 16384 - k).all()


def _bf16(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize('res', [False, True])
def test_aggregation_layer_at_bench_shape(res):
    """The dominant kernel (conv_scatter, 64->64 3x3x3, D=32, 64x64 planes -- the bench shape on a 2-volume slice)
    against F.conv3d on bf16-rounded inputs and weights: <= 2e-3 of the output's max BEFORE the bf16 store rounding
    is added (2^-9 of each value), i.e. 4e-3 in total."""
    torch.manual_seed(31)
    conv = nn.Conv3d(64, 64, 3, 1, 1, bias=True)
    with torch.no_grad():
        conv.weight.copy_(_bf16(conv.weight))
    x = _bf16(torch.randn(2, 64, 32, 64, 64))
    act = lib.ACT_NONE if res else lib.ACT_RELU
    pc = PackedConv.from_conv(conv, None, act, lib.DTYPE_BF16, 'cuda')
    kw = {}
    with torch.no_grad():
        ref = conv(x)
    if res:
        r = _bf16(torch.randn(2, 64, 32, 64, 64))
        ref = ref + r
        kw['residual'] = to_cl(r).to(torch.bfloat16).cuda()
    else:
        ref = F.relu(ref)
    got = pc(to_cl(x).to(torch.bfloat16).cuda(), engine='igemm', **kw).float().cpu()
    ref = to_cl(ref)
    scale = ref.abs().max().item()
    err = (got - ref).abs()
    assert err.max().item() <= 4e-3 * scale
    # against the fp32 value rounded the way the kernel stores it only the accumulation order is left: at most ONE
    # bf16 ulp (a value on a rounding boundary; ulp <= 2^-7 of the value), and that for a small minority of outputs
    assert (got - _bf16(ref)).abs().max().item() <= scale * 2.0 ** -7
    assert (got == _bf16(ref)).float().mean().item() > 0.98


def test_fused_classifier_at_bench_shape():
    """cls_fused at D=32, 64x64, C=64 (bench shape, 2 volumes) == Conv3d(64,1,3,1,1) + soft-argmin on rounded inputs."""
    g = torch.Generator().manual_seed(32)
    N, C, D, h, w = 2, 64, 32, 64, 64
    x = torch.randn(N, C, D, h, w, generator=g).to(torch.bfloat16)
    wt = (torch.randn(1, C, 3, 3, 3, generator=g) * 0.15).to(torch.bfloat16)
    ref = O.soft_argmin(F.conv3d(x.float(), wt.float(), padding=1).squeeze(1))
    w_taps = torch.zeros(32, C, dtype=torch.bfloat16)
    w_taps[:27] = wt[0].reshape(C, 27).t()
    got = ops.cls_soft_argmin(x.permute(0, 2, 3, 4, 1).contiguous().cuda(), w_taps.cuda(), -1.0).cpu()
    assert ref.std() > 1.0
    assert (got - ref).abs().max().item() <= 2e-3 * (D - 1)                     # <= 2e-3 relative to the disparity range
    torch.testing.assert_close(got, ref, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize('C,D', [(64, 128), (32, 128), (64, 32)])
def test_costvolume_sweep_corners(C, D):
    """configs[4] corners at 1/4 resolution (64x64): concat volume exact, fused correlation + soft-argmax and the
    stand-alone soft-argmin against the oracle's functions."""
    B, h, w = 1, 64, 64
    g = torch.Generator().manual_seed(41)
    f = _bf16(torch.randn(2 * B, C, h, w, generator=g) * 1.5)
    feat = to_cl(f).to(torch.bfloat16).cuda()
    vol = ops.cost_volume_concat(feat, B, D)
    ref_vol = torch.cat([O.build_concat_volume(f[:B], f[B:], D, -1), O.build_concat_volume(f[B:], f[:B], D, +1)], 0)
    assert torch.equal(vol.float().cpu(), ref_vol.permute(0, 2, 3, 4, 1).contiguous())
    cost_ref = torch.cat([O.build_corr_volume(f[:B], f[B:], D, -1), O.build_corr_volume(f[B:], f[:B], D, +1)], 0)
    ref = O.soft_argmax(cost_ref)
    got = ops.corr_soft_argmin(feat, B, D).cpu()                                # tensor-core path (bf16, w <= 64)
    torch.testing.assert_close(got, ref, rtol=1e-3, atol=1e-3 * D)
    got32, cost = ops.corr_soft_argmin(to_cl(f).cuda(), B, D, want_cost=True)   # exact fp32 SIMT path
    torch.testing.assert_close(cost.cpu(), cost_ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(got32.cpu(), ref, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(ops.soft_argmin(-cost_ref.cuda(), -1.0).cpu(), ref, rtol=1e-4, atol=1e-3)


def test_zz_report_measured(capsys):
    """Not a check: prints the measured errors of this module (pytest -s) so that profiles/ can quote them."""
    with capsys.disabled():
        for k in sorted(MEASURED):
            print('\n[measured] %-24s %s' % (k, ' '.join('%.3g' % v for v in MEASURED[k])), end='')
        print()
