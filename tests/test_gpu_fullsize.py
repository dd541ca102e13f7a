"""BASELINE.json's full sizes and LARGE batches, checked through size-independent properties: structure of the concat
volume, invariances of soft-argmin, batch-composition invariance and determinism of the whole forward, exactness of the
integer IoU counts.  (Value parity with the CPU oracle at the default configuration itself is
tests/test_gpu_baseline_configs.py; these tests cover what a 2-pair oracle run cannot: B = 16..128, micro-batching.)"""
import pytest
import torch

from config import cfg as default_cfg
from oracle import models as O
from stereo_3d_reconstruction_b200 import models as M, ops
from stereo_3d_reconstruction_b200.utils import synthetic

pytestmark = pytest.mark.gpu


def test_concat_volume_structure_full_size():
    """B=64, C=32, D=32, 64x64 (bench shape): ref half is constant over d, target half is the shifted map."""
    B, C, D, h, w = 64, 32, 32, 64, 64
    g = torch.Generator(device='cuda').manual_seed(0)
    feat = torch.randn(2 * B, 1, h, w, C, device='cuda', generator=g).to(torch.bfloat16)
    vol = ops.cost_volume_concat(feat, B, D)
    assert vol.shape == (2 * B, D, h, w, 2 * C)
    f = feat[:, 0]
    for d in (0, 1, 17, D - 1):
        assert torch.equal(vol[:, d, :, :, :C], f)                                   # reference half
        assert torch.equal(vol[:B, d, :, d:, C:], f[B:, :, :w - d])                  # left-ref: target at x-d
        assert vol[:B, d, :, :d, C:].abs().sum() == 0
        assert torch.equal(vol[B:, d, :, :w - d, C:], f[:B, :, d:])                  # right-ref: target at x+d
        assert vol[B:, d, :, w - d:, C:].abs().sum() == 0


def test_soft_argmin_properties_full_size():
    N, D, h, w = 128, 32, 64, 64
    g = torch.Generator(device='cuda').manual_seed(1)
    cost = torch.randn(N, D, h, w, device='cuda', generator=g) * 4
    d0 = ops.soft_argmin(cost, -1.0).clone()
    assert d0.min() >= 0 and d0.max() <= D - 1
    d1 = ops.soft_argmin(cost + 3.25, -1.0)                                          # shift invariance
    torch.testing.assert_close(d1, d0, rtol=1e-4, atol=1e-4)
    idx = torch.randint(0, D, (N, h, w), device='cuda', generator=g)                 # (near) one-hot -> index
    onehot = torch.full((N, D, h, w), 60.0, device='cuda').scatter_(1, idx.unsqueeze(1), 0.0)
    torch.testing.assert_close(ops.soft_argmin(onehot, -1.0), idx.float(), rtol=0, atol=1e-4)
    rev = ops.soft_argmin(cost.flip(1), -1.0)                                        # reversing d mirrors the result
    torch.testing.assert_close(rev, (D - 1) - d0, rtol=1e-4, atol=1e-3)


def test_forward_full_size_invariances():
    """Default config (256x256, D=32, 73 M parameters), bf16: determinism, batch-composition invariance,
    micro-batch invariance, IoU counts exact for the produced voxels."""
    cfg = default_cfg.clone()
    cfg.NETWORK.PRECISION = 'bf16'
    cfg.CONST.MICRO_BATCH = 64
    model = M.build_model('Stereo2Voxel', cfg, seed=0).cuda().pack()
    B = 16
    left, right, _ = synthetic.stereo_pair(B, 256, 256, 64, seed=3, device='cuda')
    gt = synthetic.gt_volume(B, seed=4, device='cuda')
    with torch.no_grad():
        a = [t.clone() for t in model(left, right, gt)]
        b = [t.clone() for t in model(left, right, gt)]
        for x, y in zip(a, b):
            assert torch.equal(x, y)                                                   # deterministic
        perm = torch.randperm(B, device='cuda')
        c = [t.clone() for t in model(left[perm], right[perm], gt[perm])]
        for x, y in zip(a, c):
            assert torch.equal(x[perm], y)                                             # samples are independent
        cfg.CONST.MICRO_BATCH = 4
        d = [t.clone() for t in model(left, right, gt)]
        for x, y in zip(a, d):
            assert torch.equal(x, y)                                                   # micro-batching changes nothing
    dl, dr, vox, iou = a
    assert dl.shape == (B, 1, 256, 256) and vox.shape == (B, 32, 32, 32)
    assert vox.min() >= 0 and vox.max() <= 1 and torch.isfinite(dl).all() and dl.min() >= 0 and dl.max() <= 4 * 31
    assert torch.equal(iou.cpu(), O.iou_counts(vox.cpu(), gt.cpu(), cfg.TEST.VOXEL_THRESH))


def test_corr_soft_argmin_recovers_shift_full_size():
    """configs[4] shape (D=64, C=32): target features = reference shifted by d  =>  disparity d."""
    B, C, h, w, D, d = 8, 32, 64, 64, 64, 11
    g = torch.Generator(device='cuda').manual_seed(5)
    fl = torch.randn(B, 1, h, w, C, device='cuda', generator=g) * 4
    fr = torch.zeros_like(fl)
    fr[:, :, :, :w - d] = fl[:, :, :, d:]
    feat = torch.cat([fl, fr], 0).to(torch.bfloat16)
    disp = ops.corr_soft_argmin(feat, B, D)
    assert (disp[:B, :, 2 * d:].round() == d).float().mean() > 0.95
    assert (disp[B:, :, :w - 2 * d].round() == d).float().mean() > 0.95
