"""Parity of the bandwidth-bound kernels and the SIMT conv against the oracle (GPU)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import models as O
from oracle import chamfer as OC
from stereo_3d_reconstruction_b200 import lib, ops
from stereo_3d_reconstruction_b200.layers import PackedConv
from stereo_3d_reconstruction_b200.utils import synthetic
from tests.emulate import to_cl, pad_c

pytestmark = pytest.mark.gpu


def cl_feat(f):            # [N,C,h,w] -> [N,1,h,w,C]
    return to_cl(f)


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('B,C,h,w,D', [(2, 32, 16, 16, 8), (1, 16, 9, 21, 32), (3, 64, 5, 40, 16)])
def test_cost_volume_concat(dtype, B, C, h, w, D):
    g = torch.Generator().manual_seed(0)
    f = torch.randn(2 * B, C, h, w, generator=g).to(dtype).float()
    ref = torch.cat([O.build_concat_volume(f[:B], f[B:], D, -1), O.build_concat_volume(f[B:], f[:B], D, +1)], 0)
    got = ops.cost_volume_concat(cl_feat(f).to(dtype).cuda(), B, D)
    assert torch.equal(got.float().cpu(), ref.permute(0, 2, 3, 4, 1).contiguous())      # pure data movement: exact


@pytest.mark.parametrize('N,D,h,w', [(2, 8, 16, 16), (3, 32, 7, 13), (1, 128, 4, 33), (2, 5, 3, 3)])
def test_soft_argmin(N, D, h, w):
    g = torch.Generator().manual_seed(1)
    cost = torch.randn(N, D, h, w, generator=g) * 3
    ref = O.soft_argmin(cost)
    got = ops.soft_argmin(cost.cuda(), -1.0).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('N,C,D,h,w', [(2, 16, 8, 9, 13), (1, 64, 32, 16, 16), (1, 32, 1, 3, 3), (1, 16, 2, 1, 5),
                                       (3, 64, 5, 40, 21), (2, 32, 9, 64, 64)])
def test_cls_soft_argmin_fused_equals_cout1_conv(N, C, D, h, w):
    """One-pass classifier (cls_fused.cu): tcgen05 projections + in-kernel gather + online softmax
    == Conv3d(C,1,3,1,1) -> soft-argmin on the same bf16-rounded inputs, and == the unfused two-kernel path."""
    g = torch.Generator().manual_seed(13)
    x = torch.randn(N, C, D, h, w, generator=g).to(torch.bfloat16)
    wt = (torch.randn(1, C, 3, 3, 3, generator=g) * 0.3).to(torch.bfloat16)
    cost_ref = F.conv3d(x.float(), wt.float(), padding=1).squeeze(1)
    ref = O.soft_argmin(cost_ref)
    w_taps = torch.zeros(32, C, dtype=torch.bfloat16)
    w_taps[:27] = wt[0].reshape(C, 27).t()
    xc = x.permute(0, 2, 3, 4, 1).contiguous().cuda()                              # [N,D,h,w,C]
    got = ops.cls_soft_argmin(xc, w_taps.cuda(), -1.0)
    torch.testing.assert_close(got.cpu(), ref, rtol=2e-4, atol=2e-4 * max(D, 1))
    # the two-kernel path on the same inputs
    taps = torch.einsum('ncdhw,tc->ndhtw', x.float(), w_taps.float()).contiguous()
    two = ops.tap_gather_soft_argmin(taps.cuda(), -1.0)
    torch.testing.assert_close(got, two, rtol=1e-4, atol=1e-4 * max(D, 1))


@pytest.mark.parametrize('N,D,h,w', [
    (2, 8, 32, 16),       # 4 columns: patch borders in x, one CTA pair per two columns
    (1, 5, 40, 21),       # ragged patches: pixels outside the image must enter the classifier as zeros
    (2, 1, 8, 8),         # a single plane
    (2, 2, 33, 9),        # two planes, one-pixel ragged edges
    (3, 32, 64, 64),      # the network's volume shape, 48 columns
    (20, 4, 64, 64),      # 320 columns on 148 SMs: several columns per CTA, phantom columns in the last round
])
def test_conv_cls_chain_equals_conv_then_classifier(N, D, h, w):
    """conv_scatter_cls.cu: cls_a (3x3x3 64 -> 64 + ReLU) + Cout = 1 classifier + soft-argmin in one march (the layer's output never
    written) == the plane-scatter conv followed by the one-pass classifier on its STORED bf16 output (same rounding of Y; only the
    fp32 summation order of the 27 taps differs), and == Conv3d -> ReLU -> bf16 -> Conv3d(64,1) -> soft-argmin in torch."""
    import torch.nn as nn
    from stereo_3d_reconstruction_b200 import lib
    from stereo_3d_reconstruction_b200.layers import PackedConv
    g = torch.Generator().manual_seed(17)
    conv = nn.Conv3d(64, 64, 3, 1, 1, bias=True)
    with torch.no_grad():
        conv.weight.copy_((torch.randn(conv.weight.shape, generator=g) * 0.04).to(torch.bfloat16).float())
        conv.bias.copy_(torch.randn(64, generator=g) * 0.1)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, lib.DTYPE_BF16, 'cuda')
    x = torch.randn(N, 64, D, h, w, generator=g).to(torch.bfloat16)
    wt = (torch.randn(1, 64, 3, 3, 3, generator=g) * 0.3).to(torch.bfloat16)
    w_taps = torch.zeros(32, 64, dtype=torch.bfloat16)
    w_taps[:27] = wt[0].reshape(64, 27).t()
    xc = x.permute(0, 2, 3, 4, 1).contiguous().cuda()
    got = ops.conv_cls_soft_argmin(pc, xc, w_taps.cuda(), -1.0)
    y = pc(xc, engine='igemm')                                               # the stored bf16 volume
    two = ops.cls_soft_argmin(y, w_taps.cuda(), -1.0)
    torch.cuda.synchronize()
    torch.testing.assert_close(got, two, rtol=2e-4, atol=2e-4 * max(D, 1))
    with torch.no_grad():
        yr = F.relu(conv(x.float())).to(torch.bfloat16).float()
        ref = O.soft_argmin(F.conv3d(yr, wt.float(), padding=1).squeeze(1))
    # Y itself carries the conv engine's bf16 rounding (one ulp flips against torch's fp32 accumulation order): looser than above
    torch.testing.assert_close(got.cpu(), ref, rtol=5e-3, atol=5e-3 * max(D, 1))
    # deterministic: no atomics, fixed summation order
    again = ops.conv_cls_soft_argmin(pc, xc, w_taps.cuda(), -1.0)
    assert torch.equal(got, again)


@pytest.mark.parametrize('dtype,tol', [(torch.float32, 1e-3), (torch.bfloat16, 1e-3)])
@pytest.mark.parametrize('B,C,h,w,D', [(2, 32, 8, 64, 32), (1, 16, 5, 21, 8), (1, 64, 3, 35, 64), (1, 32, 2, 40, 128)])
def test_corr_soft_argmin(dtype, tol, B, C, h, w, D):
    g = torch.Generator().manual_seed(2)
    f = torch.randn(2 * B, C, h, w, generator=g).to(dtype).float()     # same rounded inputs for both sides
    cost_ref = torch.cat([O.build_corr_volume(f[:B], f[B:], D, -1), O.build_corr_volume(f[B:], f[:B], D, +1)], 0)
    ref = O.soft_argmax(cost_ref)
    got, cost = ops.corr_soft_argmin(cl_feat(f).to(dtype).cuda(), B, D, want_cost=True)
    torch.testing.assert_close(cost.cpu(), cost_ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(got.cpu(), ref, rtol=tol, atol=tol)


@pytest.mark.parametrize('N,h,w,H,W', [(2, 16, 16, 64, 64), (1, 35, 35, 137, 137), (3, 8, 12, 32, 48)])
def test_upsample_disp(N, h, w, H, W):
    g = torch.Generator().manual_seed(3)
    q = torch.rand(N, h, w, generator=g) * 8
    ref = F.interpolate(q.unsqueeze(1) * 4.0, size=(H, W), mode='bilinear', align_corners=False).squeeze(1)
    got = ops.upsample_disp(q.cuda(), H, W, 4.0).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('B,N,M,dup', [(2, 300, 1000, True), (1, 1, 1, False), (3, 17, 5, True), (2, 2048, 4096, True),
                                       (1, 1025, 1023, False)])
def test_chamfer_bit_exact(B, N, M, dup):
    a, b = synthetic.point_clouds(B, N, M, seed=5, duplicates=dup)
    rd1, rd2, ri1, ri2 = OC.chamfer_c(a.numpy(), b.numpy())
    d1, d2, i1, i2 = ops.chamfer_forward(a.cuda(), b.cuda())
    assert np.array_equal(i1.cpu().numpy(), ri1) and np.array_equal(i2.cpu().numpy(), ri2)          # bit-exact indices
    assert np.array_equal(d1.cpu().numpy(), rd1) and np.array_equal(d2.cpu().numpy(), rd2)          # bit-exact distances


@pytest.mark.parametrize('B,N,M,dup', [(2, 300, 1000, True), (1, 1, 1, False), (3, 17, 5, True), (2, 2048, 4096, True),
                                       (1, 1025, 1023, False), (2, 4100, 2500, True), (1, 257, 33, True), (4, 31, 513, False)])
@pytest.mark.parametrize('R', [4, 8])
def test_chamfer_symmetric_one_pass_bit_exact(B, N, M, dup, R):
    """The symmetric kernel (every pair evaluated once for both directions; packed fp32 ops, value-only minima, index by
    chunk rescan, 64-bit atomicMin keys) forced at every size: bit-identical to the C oracle, whichever set is the larger,
    with ragged tails, several shared-memory tiles of the streamed set (> 2048 points) and exact ties."""
    a, b = synthetic.point_clouds(B, N, M, seed=7, duplicates=dup)
    rd1, rd2, ri1, ri2 = OC.chamfer_c(a.numpy(), b.numpy())
    lib.set_knob('chamfer_sym', 1)
    lib.set_knob('chamfer_sym_r', R)                        # resident points per lane: both instantiations
    try:
        assert lib.load().s3d_chamfer_workspace_bytes(B, N, M) == 8 * B * min(N, M)
        d1, d2, i1, i2 = ops.chamfer_forward(a.cuda(), b.cuda())
    finally:
        lib.set_knob('chamfer_sym', 0)
        lib.set_knob('chamfer_sym_r', 0)
    assert np.array_equal(i1.cpu().numpy(), ri1) and np.array_equal(i2.cpu().numpy(), ri2)
    assert np.array_equal(d1.cpu().numpy(), rd1) and np.array_equal(d2.cpu().numpy(), rd2)


def test_chamfer_symmetric_ties_and_non_finite():
    """All points identical (every distance ties: index 0 everywhere), NaN points (never win; an all-NaN set leaves +inf /
    index 0), coordinates whose squared distance overflows to +inf (never 'less than' the initial +inf: index 0)."""
    lib.set_knob('chamfer_sym', 1)
    try:
        a = torch.zeros(2, 300, 3);  b = torch.zeros(2, 700, 3)
        cases = [(a.clone(), b.clone())]
        a2, b2 = synthetic.point_clouds(2, 300, 700, seed=9, duplicates=True)
        a2[0, 5] = float('nan');  b2[1, 100:400] = float('nan');  cases.append((a2, b2))
        a3, b3 = synthetic.point_clouds(1, 64, 600, seed=10)
        a3[:] = float('nan');  cases.append((a3, b3))
        a4, b4 = synthetic.point_clouds(1, 40, 300, seed=11)
        a4[0, :20] = 3e19;  b4[0, :10] = -3e19;  cases.append((a4, b4))
        for x, y in cases:
            rd1, rd2, ri1, ri2 = OC.chamfer_c(x.numpy(), y.numpy())
            d1, d2, i1, i2 = ops.chamfer_forward(x.cuda(), y.cuda())
            assert np.array_equal(i1.cpu().numpy(), ri1) and np.array_equal(i2.cpu().numpy(), ri2)
            assert np.array_equal(d1.cpu().numpy(), rd1, equal_nan=True) and np.array_equal(d2.cpu().numpy(), rd2, equal_nan=True)
    finally:
        lib.set_knob('chamfer_sym', 0)


def test_chamfer_rejects_empty():
    with pytest.raises(lib.S3dError):
        ops.chamfer_forward(torch.zeros(1, 0, 3).cuda(), torch.zeros(1, 4, 3).cuda())


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_fuse_views_and_iou(dtype):
    g = torch.Generator().manual_seed(4)
    B, V, nv = 3, 2, 32 ** 3
    score = torch.randn(V * B, nv, 16, generator=g).to(dtype)
    vol = torch.rand(V * B, nv, 16, generator=g).to(dtype)
    gt = synthetic.gt_volume(B, seed=7)
    th = [0.2, 0.3, 0.4, 0.5]
    s = score[..., 0].float().view(V, B, nv)
    v = vol[..., 8].float().view(V, B, nv)
    ref = torch.clamp((F.softmax(s, 0) * v).sum(0), 0, 1)
    iou = torch.zeros(B, 4, 2, dtype=torch.int64).cuda()
    got = ops.fuse_views(score.cuda(), 0, 16, vol.cuda(), 8, 16, B, V, nv, gt=gt.view(B, -1).cuda(), thresholds=th,
                         iou=iou).cpu()
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    # integer stats must be exact for the fused values the kernel itself produced
    ref_iou = O.iou_counts(got.view(B, 32, 32, 32), gt, th)
    assert torch.equal(iou.cpu(), ref_iou)


def test_latent_to_vox_and_pool():
    g = torch.Generator().manual_seed(6)
    N, C, H, W, L = 3, 16, 5, 7, 2
    x = torch.randn(N, C, H, W, generator=g)
    pooled = F.adaptive_avg_pool2d(x, L)
    ref_vox = pooled.reshape(N, C * L * L // 8, 2, 2, 2).permute(0, 2, 3, 4, 1).contiguous()
    got = ops.latent_to_vox(to_cl(x).cuda(), L).cpu()
    torch.testing.assert_close(got, ref_vox, rtol=1e-5, atol=1e-6)
    got2 = ops.avg_pool(to_cl(x).cuda(), L).cpu()
    torch.testing.assert_close(got2, to_cl(pooled), rtol=1e-5, atol=1e-6)


def test_pack_image():
    g = torch.Generator().manual_seed(8)
    img = torch.rand(2, 3, 9, 11, generator=g)
    disp = torch.rand(2, 9, 11, generator=g) * 30
    got = ops.pack_image(img.cuda(), disp.cuda(), 0.25, dtype=torch.float32).cpu()
    assert got.shape == (2, 1, 9, 11, 16)
    torch.testing.assert_close(got[:, 0, :, :, :3], img.permute(0, 2, 3, 1))
    torch.testing.assert_close(got[:, 0, :, :, 3], disp * 0.25)
    assert got[..., 4:].abs().max() == 0


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_pack_image_u8(dtype):
    g = torch.Generator().manual_seed(9)
    img = torch.randint(0, 256, (2, 9, 11, 3), generator=g, dtype=torch.uint8)       # decoded PNG layout (HWC)
    disp = torch.rand(2, 9, 11, generator=g) * 30
    got = ops.pack_image(img.cuda(), disp.cuda(), 0.25, dtype=dtype).cpu()
    assert got.shape == (2, 1, 9, 11, 16)
    ref = (img.float() * (1.0 / 255.0)).to(dtype)
    assert torch.equal(got[:, 0, :, :, :3], ref)
    assert torch.equal(got[:, 0, :, :, 3], (disp * 0.25).to(dtype))
    assert got[..., 4:].abs().max() == 0


@pytest.mark.parametrize('simt', [0, 1])
@pytest.mark.parametrize('dtype_code,dtype,tol', [(lib.DTYPE_F32, torch.float32, 1e-5), (lib.DTYPE_BF16, torch.bfloat16, 2e-2)])
@pytest.mark.parametrize('with_disp,u8,H,W,cout', [(False, False, 16, 20, 32), (True, False, 9, 11, 16), (False, True, 12, 12, 64),
                                                    (True, True, 7, 5, 32), (False, True, 131, 70, 32), (True, False, 70, 131, 32)])
def test_conv_first_direct_from_raw_image(knobs, simt, dtype_code, dtype, tol, with_disp, u8, H, W, cout):
    """conv_first.cu (SIMT) / conv_first_tc.cu (bf16, 32 channels: im2col rows built by the threads + tcgen05; knob
    no_conv_first_tc selects the SIMT kernel) == Conv2d(3|4, cout, 3, stride 2, pad 1) + bias + ReLU on the raw image (fp32 NCHW or
    uint8 HWC)."""
    knobs('no_conv_first_tc', simt)
    g = torch.Generator().manual_seed(21)
    cin = 4 if with_disp else 3
    conv = nn.Conv2d(cin, cout, 3, 2, 1)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * 0.3)
        conv.bias.copy_(torch.randn(cout, generator=g) * 0.1)
    if u8:
        raw = torch.randint(0, 256, (2, H, W, 3), generator=g, dtype=torch.uint8)
        img = (raw.float() * (1.0 / 255.0)).permute(0, 3, 1, 2).contiguous()
    else:
        raw = img = torch.rand(2, 3, H, W, generator=g)
    disp = torch.rand(2, H, W, generator=g) * 20 if with_disp else None
    x = img if disp is None else torch.cat([img, (disp * 0.05).unsqueeze(1)], 1)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, dtype_code, 'cuda')
    if dtype == torch.bfloat16:           # the kernel rounds inputs and weights to the storage type, like the engine it replaces
        x = x.to(dtype).float()
        ref = F.relu(F.conv2d(x, conv.weight.to(dtype).float(), conv.bias, 2, 1))
    else:
        ref = F.relu(conv(x))
    got = ops.conv_first(raw.cuda(), pc, None if disp is None else disp.cuda(), 0.05)
    assert got.shape == (2, 1, (H - 1) // 2 + 1, (W - 1) // 2 + 1, pc.cout_pad) and got.dtype == dtype
    torch.testing.assert_close(got[:, 0, :, :, :cout].float().cpu(), ref.permute(0, 2, 3, 1), rtol=tol, atol=tol)
    # and against the path it replaces (staged copy + implicit-GEMM / SIMT engine), same rounding points
    staged = ops.pack_image(raw.cuda(), None if disp is None else disp.cuda(), 0.05, dtype=dtype)
    old = pc(staged, engine='igemm' if dtype == torch.bfloat16 else 'direct')
    torch.testing.assert_close(got.float(), old.float(), rtol=tol, atol=tol)


# ---- SIMT conv (exact fp32 engine) vs torch.nn.functional --------------------------------------
def _run(pc, x_nc, engine, dtype=torch.float32):
    x = pad_c(to_cl(x_nc), pc.cin_pad).to(dtype).cuda()
    return pc(x, engine=engine).float().cpu()


def test_direct_conv_family_fp32():
    torch.manual_seed(0)
    conv = nn.Conv2d(3, 20, 3, 2, 1)
    x = torch.randn(2, 3, 11, 13)
    pc = PackedConv.from_conv(conv, None, lib.ACT_RELU, lib.DTYPE_F32, 'cuda')
    torch.testing.assert_close(_run(pc, x, 'direct')[..., :20], to_cl(F.relu(conv(x))), rtol=1e-4, atol=1e-5)
    conv3 = nn.Conv3d(5, 7, 3, 1, 1)
    x3 = torch.randn(1, 5, 4, 6, 5)
    pc = PackedConv.from_conv(conv3, None, lib.ACT_NONE, lib.DTYPE_F32, 'cuda')
    torch.testing.assert_close(_run(pc, x3, 'direct')[..., :7], to_cl(conv3(x3)), rtol=1e-4, atol=1e-5)
    dc = nn.ConvTranspose3d(6, 5, 4, 2, 1, bias=False)
    xd = torch.randn(2, 6, 2, 3, 4)
    pc = PackedConv.from_deconv_k4s2p1(dc, None, lib.ACT_NONE, lib.DTYPE_F32, 'cuda')
    torch.testing.assert_close(_run(pc, xd, 'direct')[..., :5], to_cl(dc(xd)), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('N,C,D,h,w', [(2, 16, 8, 9, 13), (1, 64, 32, 16, 16), (1, 8, 1, 3, 3), (1, 8, 2, 1, 5)])
def test_tap_gather_soft_argmin_equals_cout1_conv(N, C, D, h, w):
    """cls_b path: pointwise per-tap projections + gather == Conv3d(C,1,3,1,1) -> soft-argmin."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, C, D, h, w, generator=g)
    wt = torch.randn(1, C, 3, 3, 3, generator=g) * 0.3
    cost_ref = F.conv3d(x, wt, padding=1).squeeze(1)
    ref = O.soft_argmin(cost_ref)
    taps = torch.einsum('ncdhw,tc->ndhtw', x, wt[0].reshape(C, 27).t())          # line-planar [N,D,h,27,w]
    taps = torch.cat([taps, torch.zeros(N, D, h, 5, w)], 3).contiguous()        # 32 tap rows per line
    got, cost = ops.tap_gather_soft_argmin(taps.cuda(), -1.0, want_cost=True)
    torch.testing.assert_close(cost.cpu(), cost_ref, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize('B,C,h,w,D', [(2, 32, 8, 64, 32), (1, 16, 5, 21, 8), (1, 64, 3, 35, 64), (1, 32, 2, 40, 128),
                                       (3, 32, 7, 64, 24), (1, 128, 4, 64, 48)])
def test_corr_soft_argmin_tensor_core_path(B, C, h, w, D):
    """bf16, w <= 64, no cost output -> corr_tc.cu (tcgen05 Gram matrix + band soft-argmax in the epilogue)."""
    g = torch.Generator().manual_seed(12)
    f = (torch.randn(2 * B, C, h, w, generator=g) * 1.5).to(torch.bfloat16).float()
    cost_ref = torch.cat([O.build_corr_volume(f[:B], f[B:], D, -1), O.build_corr_volume(f[B:], f[:B], D, +1)], 0)
    ref = O.soft_argmax(cost_ref)
    got = ops.corr_soft_argmin(cl_feat(f).to(torch.bfloat16).cuda(), B, D)
    torch.testing.assert_close(got.cpu(), ref, rtol=1e-3, atol=1e-3)
