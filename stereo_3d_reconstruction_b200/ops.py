# -*- coding: utf-8 -*-
"""Python wrappers over the non-conv C-ABI entry points (include/s3d.h).  Every wrapper validates
device / dtype / contiguity, allocates outputs with torch, passes raw pointers and the current
CUDA stream, and raises on a non-zero return code.  None of them computes anything in Python."""
import ctypes

import torch

from . import lib as _lib


def _code(t):
    if t.dtype == torch.bfloat16:
        return _lib.DTYPE_BF16
    if t.dtype == torch.float32:
        return _lib.DTYPE_F32
    raise TypeError('unsupported dtype %s' % t.dtype)


def _code_like(t, split):
    """dtype code of an activation tensor; split=True: bf16 pairs [hi(C) | lo(C)] (include/s3d.h, S3D_DTYPE_BF16X2)."""
    if split:
        assert t.dtype == torch.bfloat16 and t.shape[-1] % 2 == 0
        return _lib.DTYPE_BF16X2
    return _code(t)


def _chk(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.S3dError('s3d ops need CUDA tensors (there is no CPU fallback)')
        if not t.is_contiguous():
            raise ValueError('tensor must be contiguous')


def _stream():
    return torch.cuda.current_stream().cuda_stream


def pack_image(img, disp=None, disp_scale=1.0, cpad=16, dtype=torch.bfloat16, out=None):
    """img fp32 NCHW [B,3,H,W], or uint8 HWC [B,H,W,3] (decoded PNG; scaled by 1/255), (+ disp fp32 [B,H,W])
    -> channels-last [B,1,H,W,cpad]."""
    _chk(img, disp, out)
    if img.dtype == torch.uint8:
        B, H, W, C = img.shape
        assert C == 3
        if out is None:
            out = torch.empty((B, 1, H, W, cpad), dtype=dtype, device=img.device)
        rc = _lib.load().s3d_pack_image_u8(img.data_ptr(), disp.data_ptr() if disp is not None else None,
                                           float(disp_scale), 1.0 / 255.0, out.data_ptr(), B, H, W, cpad, _code(out),
                                           _stream())
        _lib.check(rc, 's3d_pack_image_u8')
        _lib.count_launch()
        return out
    B, C, H, W = img.shape
    assert C == 3 and img.dtype == torch.float32
    if out is None:
        out = torch.empty((B, 1, H, W, cpad), dtype=dtype, device=img.device)
    rc = _lib.load().s3d_pack_image(img.data_ptr(), disp.data_ptr() if disp is not None else None, float(disp_scale),
                                    out.data_ptr(), B, H, W, cpad, _code(out), _stream())
    _lib.check(rc, 's3d_pack_image')
    _lib.count_launch()
    return out


def cost_volume_concat(feat, B, D, C=None, out=None, split=False):
    """feat [2B,1,h,w,C] -> vol [2B,D,h,w,2C].  split: feat [.., hi(C) | lo(C)] -> vol [.., hi(2C) | lo(2C)]."""
    _chk(feat, out)
    n2, one, h, w, Cf = feat.shape
    C = Cf if C is None else C
    assert n2 == 2 * B and one == 1 and C == Cf
    if out is None:
        out = torch.empty((2 * B, D, h, w, 2 * C), dtype=feat.dtype, device=feat.device)
    assert out.shape == (2 * B, D, h, w, 2 * C) and out.dtype == feat.dtype
    rc = _lib.load().s3d_cost_volume_concat(feat.data_ptr(), out.data_ptr(), B, h, w, C // 2 if split else C, D,
                                            _code_like(feat, split), _stream())
    _lib.check(rc, 's3d_cost_volume_concat')
    _lib.count_launch()
    return out


def soft_argmin(cost, sign=-1.0, out=None):
    """cost fp32 [N,D,h,w] -> disp fp32 [N,h,w] = sum_d d * softmax_d(sign*cost)."""
    _chk(cost, out)
    assert cost.dtype == torch.float32
    N, D, h, w = cost.shape
    if out is None:
        out = torch.empty((N, h, w), dtype=torch.float32, device=cost.device)
    rc = _lib.load().s3d_soft_argmin(cost.data_ptr(), out.data_ptr(), N, D, h, w, float(sign), _stream())
    _lib.check(rc, 's3d_soft_argmin')
    _lib.count_launch()
    return out


def tap_gather_soft_argmin(taps, sign=-1.0, out=None, want_cost=False):
    """taps fp32, line-planar [N,D,h,S>=27,w] (per-tap projections) -> disp fp32 [N,h,w]; see include/s3d.h."""
    _chk(taps, out)
    assert taps.dtype == torch.float32 and taps.dim() == 5
    N, D, h, S, w = taps.shape
    if out is None:
        out = torch.empty((N, h, w), dtype=torch.float32, device=taps.device)
    cost = torch.empty((N, D, h, w), dtype=torch.float32, device=taps.device) if want_cost else None
    rc = _lib.load().s3d_tap_gather_soft_argmin(taps.data_ptr(), out.data_ptr(), cost.data_ptr() if want_cost else None,
                                                N, D, h, w, S, float(sign), _stream())
    _lib.check(rc, 's3d_tap_gather_soft_argmin')
    _lib.count_launch()
    return (out, cost) if want_cost else out


def split_tf32(x, hi=None, lo=None):
    """fp32 x -> (hi, lo): hi representable in TF32 (low 13 mantissa bits cleared), lo = x - hi (include/s3d.h)."""
    _chk(x, hi, lo)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() % 4 == 0
    hi = torch.empty_like(x) if hi is None else hi
    lo = torch.empty_like(x) if lo is None else lo
    rc = _lib.load().s3d_split_tf32(x.data_ptr(), hi.data_ptr(), lo.data_ptr(), x.numel(), _stream())
    _lib.check(rc, 's3d_split_tf32')
    _lib.count_launch()
    return hi, lo


def split_bf16(x, out=None):
    """fp32 [..., C] -> bf16 pairs [..., hi(C) | lo(C)] (include/s3d.h, S3D_DTYPE_BF16X2)."""
    _chk(x, out)
    assert x.dtype == torch.float32 and x.is_contiguous()
    C = x.shape[-1]
    if out is None:
        out = torch.empty(x.shape[:-1] + (2 * C,), dtype=torch.bfloat16, device=x.device)
    rc = _lib.load().s3d_split_bf16(x.data_ptr(), out.data_ptr(), x.numel() // C, C, _stream())
    _lib.check(rc, 's3d_split_bf16')
    _lib.count_launch()
    return out


def unsplit_bf16(x, out=None):
    """bf16 pairs [..., hi(C) | lo(C)] -> fp32 [..., C] = hi + lo."""
    _chk(x, out)
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[-1] % 2 == 0
    C = x.shape[-1] // 2
    if out is None:
        out = torch.empty(x.shape[:-1] + (C,), dtype=torch.float32, device=x.device)
    rc = _lib.load().s3d_unsplit_bf16(x.data_ptr(), out.data_ptr(), x.numel() // (2 * C), C, _stream())
    _lib.check(rc, 's3d_unsplit_bf16')
    _lib.count_launch()
    return out


def depth_to_space(x, cpad=16, proj_w=None, proj_act=0, out=None, split=False):
    """x [N,d,h,w,64] (channel = parity class * 8 + c, from PackedConv.from_deconv_k4s2p1_blocked) -> [N,2d,2h,2w,cpad];
    optional projection of the 8 features into channel 8 (include/s3d.h, s3d_depth_to_space)."""
    _chk(x, proj_w, out)
    m = 2 if split else 1                                       # split: [.., hi(64) | lo(64)] -> [.., hi(cpad) | lo(cpad)]
    assert x.dim() == 5 and x.shape[-1] == 64 * m and x.is_contiguous()
    N, d, h, w, _ = x.shape
    if out is None:
        out = torch.empty((N, 2 * d, 2 * h, 2 * w, cpad * m), dtype=x.dtype, device=x.device)
    assert out.dtype == x.dtype and out.is_contiguous() and out.shape == (N, 2 * d, 2 * h, 2 * w, cpad * m)
    if proj_w is not None:
        assert proj_w.dtype == torch.float32 and proj_w.numel() >= 8
    rc = _lib.load().s3d_depth_to_space(x.data_ptr(), out.data_ptr(), proj_w.data_ptr() if proj_w is not None else None,
                                        int(proj_act), N, d, h, w, cpad, _code_like(x, split), _stream())
    _lib.check(rc, 's3d_depth_to_space')
    _lib.count_launch()
    return out


def conv_first(img, pc, disp=None, disp_scale=1.0, out=None):
    """First encoder layer straight from the raw image: img fp32 NCHW [B,3,H,W] or uint8 HWC [B,H,W,3] (+ disp fp32 [B,H,W]
    as a 4th channel) through PackedConv `pc` (3x3, stride 2, pad 1) -> channels-last [B,1,oH,oW,cout_pad]."""
    _chk(img, disp, out)
    u8 = img.dtype == torch.uint8
    if u8:
        B, H, W, C = img.shape
    else:
        B, C, H, W = img.shape
        assert img.dtype == torch.float32
    assert C == 3 and img.is_contiguous()
    assert pc.ksize == (1, 3, 3) and pc.stride == (1, 2, 2) and pc.pad == (0, 1, 1) and pc.n_classes == 1
    cin = 4 if disp is not None else 3
    assert pc.cin == cin, (pc.cin, cin)
    oH, oW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dt = pc.weight.dtype
    co = pc.cout_pad * (2 if pc.dtype_code == _lib.DTYPE_BF16X2 else 1)      # split layer: [.., hi | lo]
    if out is None:
        out = torch.empty((B, 1, oH, oW, co), dtype=dt, device=img.device)
    assert out.dtype == dt and out.is_contiguous() and out.shape[-1] == co
    rc = _lib.load().s3d_conv_first(img.data_ptr(), 1 if u8 else 0, disp.data_ptr() if disp is not None else None,
                                    float(disp_scale), pc.weight.data_ptr(), pc.bias.data_ptr(), out.data_ptr(), B, H, W, cin,
                                    pc.cin_pad, pc.cout_pad, pc.dtype_code, pc.act, pc.act_param, _stream())
    _lib.check(rc, 's3d_conv_first')
    _lib.count_launch()
    return out


def conv_concat_volume(pc, featp, B, D, pad, out=None, ref_once=False):
    """Concat cost volume + the 3x3x3 conv `pc` (PackedConv over 2C channels) in one kernel, the volume never written.
    featp: [2B,1,h,pitch,C] feature maps stored with `pad` >= D-1 ZERO pixels on both sides of every row (real pixels in
    columns pad .. pitch-pad-1), left images first.  -> [2B,D,h,w,cout_pad].  include/s3d.h, s3d_conv_concat_volume.
    ref_once: the reference-once form (s3d_conv_concat_volume_ro: half the tensor-core work, not bit-identical)."""
    import ctypes
    from .layers import is_split
    _chk(featp, out)
    split = is_split(pc.dtype_code)              # 'bf16x3': featp / out rows are [hi(C) | lo(C)] bf16 pairs, reference-once form only
    cm = 2 if split else 1
    n2, one, h, pitch, C = featp.shape
    C //= cm
    assert n2 == 2 * B and one == 1 and featp.is_contiguous() and pc.cin_pad == 2 * C and pad >= D - 1
    assert not split or ref_once, 'split (BF16X2) operands: reference-once form only'
    w = pitch - 2 * pad
    assert w > 0
    Co = cm * pc.cout_pad
    if out is None:
        out = torch.empty((2 * B, D, h, w, Co), dtype=featp.dtype, device=featp.device)
    assert out.is_contiguous() and out.shape == (2 * B, D, h, w, Co)
    p = pc.params(2 * B, D, h, w, (D * h * w * Co, h * w * Co, w * Co, Co), _code_like(out, split), pc.cout_pad)
    if ref_once:
        rc = _lib.load().s3d_conv_concat_volume_ro(ctypes.byref(p), featp.data_ptr(), pitch, pad, pc.refonce_weights(C).data_ptr(),
                                                   pc.bias.data_ptr(), out.data_ptr(), _stream())
        _lib.check(rc, 's3d_conv_concat_volume_ro')
    else:
        rc = _lib.load().s3d_conv_concat_volume(ctypes.byref(p), featp.data_ptr(), pitch, pad, pc.bias.data_ptr(), out.data_ptr(),
                                                _stream())
        _lib.check(rc, 's3d_conv_concat_volume')
    _lib.count_launch()
    return out


def conv_concat_volume_sheared(pc, featp, B, D, pad, out=None, bufs=None):
    """Concat cost volume + the 3x3x3 layer `pc` in SHEARED form (bf16, Cout 64, ReLU): four 2-D map convolutions on the
    tensor cores (1/14 of the layer's MMAs) + one streaming pass that writes the volume (include/s3d.h,
    s3d_concat_gonce_assemble; layers.py, PackedConv.gonce_convs).  featp as for conv_concat_volume.  bufs: optional dict of
    persistent fp32 workspaces {'maps_l', 'maps_r': [B,1,h,w+4,384], 'edge_l', 'edge_r': [B,1,h,D,256]}."""
    from .layers import is_split
    _chk(featp, out)
    split = is_split(pc.dtype_code)              # 'bf16x3': featp rows and the output are bf16 pairs [hi | lo], three MMAs per product
    cm = 2 if split else 1
    n2, one, h, pitch, C = featp.shape
    C //= cm
    assert n2 == 2 * B and one == 1 and featp.is_contiguous() and featp.dtype == torch.bfloat16
    assert pc.cin_pad == 2 * C and pc.cout_pad == 64 and pc.act == _lib.ACT_RELU and pad >= D - 1 and D >= 2
    w = pitch - 2 * pad
    g = pc.gonce_convs(C, pad, w, D)
    bufs = {} if bufs is None else bufs
    res = {}
    use_mc = C == 32 and not _lib.KNOBS['no_map_conv']       # the halo-once map engine (csrc/map_conv.cu): 32 feature channels
    for name, key, sl in (('left', 'maps_l', slice(0, B)), ('right', 'maps_r', slice(B, 2 * B)),
                          ('edge_left', 'edge_l', slice(0, B)), ('edge_right', 'edge_r', slice(B, 2 * B))):
        conv = g[name]
        off, ntx, ow = g[name + '_geom']
        o = bufs.get(key)
        if o is None:
            o = torch.empty((B, 1, h, ow, conv.cout_pad), dtype=torch.float32, device=featp.device)
        assert o.shape == (B, 1, h, ow, conv.cout_pad) and o.dtype == torch.float32 and o.is_contiguous()
        if use_mc:
            x = featp[sl]
            rc = _lib.load().s3d_map_conv(x.data_ptr(), conv.weight.data_ptr(), o.data_ptr(), B, h, pitch, ow, off, ntx,
                                          conv.cout_pad, pc.dtype_code, _stream())
            _lib.check(rc, 's3d_map_conv')
            _lib.count_launch()
            res[key] = o
        else:
            res[key] = conv(featp[sl], out=o)
    Co = cm * 64
    if out is None:
        out = torch.empty((2 * B, D, h, w, Co), dtype=torch.bfloat16, device=featp.device)
    assert out.is_contiguous() and out.shape == (2 * B, D, h, w, Co) and out.dtype == torch.bfloat16
    rc = _lib.load().s3d_concat_gonce_assemble(res['maps_l'].data_ptr(), res['maps_r'].data_ptr(), res['edge_l'].data_ptr(),
                                               res['edge_r'].data_ptr(), pc.bias.data_ptr(), out.data_ptr(), B, D, h, w, w + 4,
                                               _lib.DTYPE_BF16X2 if split else _lib.DTYPE_BF16, _stream())
    _lib.check(rc, 's3d_concat_gonce_assemble')
    _lib.count_launch()
    return out


def cls_soft_argmin(x, w_taps, sign=-1.0, out=None):
    """x bf16 [N,D,h,w,C] (aggregated volume), w_taps bf16 [32,C] (27 classifier taps) -> disp fp32 [N,h,w]: the Cout=1
    3x3x3 classifier and the soft-argmin in one pass (include/s3d.h, s3d_cls_soft_argmin)."""
    _chk(x, w_taps, out)
    assert x.dtype == torch.bfloat16 and w_taps.dtype == torch.bfloat16 and x.dim() == 5 and x.is_contiguous()
    N, D, h, w, C = x.shape
    assert w_taps.numel() == 32 * C and w_taps.is_contiguous()
    if out is None:
        out = torch.empty((N, h, w), dtype=torch.float32, device=x.device)
    rc = _lib.load().s3d_cls_soft_argmin(x.data_ptr(), w_taps.data_ptr(), out.data_ptr(), N, D, h, w, C, float(sign),
                                         _stream())
    _lib.check(rc, 's3d_cls_soft_argmin')
    _lib.count_launch()
    return out


def conv_cls_workspace_bytes(N, D, h, w):
    return int(_lib.load().s3d_conv_cls_workspace_bytes(N, D, h, w))


def conv_cls_soft_argmin(pc, x, w_taps, sign=-1.0, out=None, workspace=None):
    """The last aggregation layer `pc` (PackedConv, bf16 3x3x3 64 -> 64 + ReLU) over x [N,D,h,w,64], the Cout = 1 classifier
    (w_taps bf16 [32,64]) and the soft-argmin, the layer's output volume never written (include/s3d.h,
    s3d_conv_cls_soft_argmin; csrc/conv_scatter_cls.cu).  workspace: >= conv_cls_workspace_bytes(N, D, h, w) bytes.
    -> disp fp32 [N,h,w]."""
    import ctypes
    _chk(x, w_taps, out, workspace)
    assert x.dtype == torch.bfloat16 and w_taps.dtype == torch.bfloat16 and x.dim() == 5 and x.is_contiguous()
    N, D, h, w, C = x.shape
    assert C == 64 and pc.cin_pad == 64 and pc.cout_pad == 64 and w_taps.numel() == 32 * 64 and w_taps.is_contiguous()
    need = conv_cls_workspace_bytes(N, D, h, w)
    if workspace is None:
        workspace = torch.empty(need, dtype=torch.uint8, device=x.device)
    assert workspace.is_contiguous() and workspace.numel() * workspace.element_size() >= need
    if out is None:
        out = torch.empty((N, h, w), dtype=torch.float32, device=x.device)
    p = pc.params(N, D, h, w, (D * h * w * 64, h * w * 64, w * 64, 64), _lib.DTYPE_BF16, 64)
    rc = _lib.load().s3d_conv_cls_soft_argmin(ctypes.byref(p), x.data_ptr(), pc.bias.data_ptr(), w_taps.data_ptr(),
                                              workspace.data_ptr(), out.data_ptr(), float(sign), _stream())
    _lib.check(rc, 's3d_conv_cls_soft_argmin')
    _lib.count_launch(2)
    return out


def corr_soft_argmin(feat, B, D, out=None, want_cost=False, c_real=None):
    """feat [2B,1,h,w,C] -> disp fp32 [2B,h,w] (fused correlation + soft-argmax).  c_real: number of REAL feature
    channels when C is a padded pitch (the correlation is a mean over the real channels; padded ones are zero)."""
    _chk(feat, out)
    n2, one, h, w, C = feat.shape
    assert n2 == 2 * B and one == 1
    if out is None:
        out = torch.empty((2 * B, h, w), dtype=torch.float32, device=feat.device)
    cost = torch.empty((2 * B, D, h, w), dtype=torch.float32, device=feat.device) if want_cost else None
    rc = _lib.load().s3d_corr_soft_argmin(feat.data_ptr(), out.data_ptr(), cost.data_ptr() if want_cost else None,
                                          B, h, w, C, int(c_real or 0), D, _code(feat), _stream())
    _lib.check(rc, 's3d_corr_soft_argmin')
    _lib.count_launch()
    return (out, cost) if want_cost else out


def upsample_disp(disp_q, H, W, scale, out=None):
    _chk(disp_q, out)
    assert disp_q.dtype == torch.float32
    N, h, w = disp_q.shape
    if out is None:
        out = torch.empty((N, H, W), dtype=torch.float32, device=disp_q.device)
    rc = _lib.load().s3d_upsample_disp(disp_q.data_ptr(), out.data_ptr(), N, h, w, H, W, float(scale), _stream())
    _lib.check(rc, 's3d_upsample_disp')
    _lib.count_launch()
    return out


def latent_to_vox(x, L, out=None, split=False):
    """[N,1,H,W,C] -> pooled to LxL -> [N,2,2,2,C*L*L/8] (oracle's NCHW .view order).  split: hi | lo pairs in and out."""
    _chk(x, out)
    N, one, H, W, C = x.shape
    if out is None:
        out = torch.empty((N, 2, 2, 2, C * L * L // 8), dtype=x.dtype, device=x.device)
    assert out.shape[-1] == C * L * L // 8, 'latent_to_vox writes K = C*L*L/8 channels densely'
    rc = _lib.load().s3d_latent_to_vox(x.data_ptr(), out.data_ptr(), N, H, W, C // 2 if split else C, L,
                                       _code_like(x, split), _stream())
    _lib.check(rc, 's3d_latent_to_vox')
    _lib.count_launch()
    return out


def avg_pool(x, L, out=None, split=False):
    _chk(x, out)
    N, one, H, W, C = x.shape
    if out is None:
        out = torch.empty((N, 1, L, L, C), dtype=x.dtype, device=x.device)
    rc = _lib.load().s3d_avg_pool(x.data_ptr(), out.data_ptr(), N, H, W, C // 2 if split else C, L, _code_like(x, split),
                                  _stream())
    _lib.check(rc, 's3d_avg_pool')
    _lib.count_launch()
    return out


def fuse_views(score, score_off, score_stride, vol, vol_off, vol_stride, B, V, nvox, gt=None, thresholds=None,
               iou=None, out=None, score_lo=0, vol_lo=0):
    """Context-aware fusion epilogue (+ IoU counts).  score/vol: tensors holding [V*B, nvox] planes
    at element offset *_off with element stride *_stride between voxels.  score_lo / vol_lo > 0: split (hi | lo) tensors,
    the lo part of a value sits that many elements after its hi part."""
    _chk(score, vol, gt, iou, out)
    assert score.dtype == vol.dtype
    if out is None:
        out = torch.empty((B, nvox), dtype=torch.float32, device=score.device)
    T = 0
    th = None
    if gt is not None:
        assert gt.dtype == torch.uint8 and iou is not None and iou.dtype == torch.int64
        T = len(thresholds)
        th = (ctypes.c_float * T)(*[float(t) for t in thresholds])
    esz = score.element_size()
    rc = _lib.load().s3d_fuse_views(score.data_ptr() + score_off * esz, score_stride,
                                    vol.data_ptr() + vol_off * esz, vol_stride,
                                    _lib.DTYPE_BF16X2 if score_lo else _code(score), out.data_ptr(), B, V,
                                    nvox, gt.data_ptr() if gt is not None else None, th, T,
                                    iou.data_ptr() if iou is not None else None, int(score_lo), int(vol_lo), _stream())
    _lib.check(rc, 's3d_fuse_views')
    _lib.count_launch()
    return out


def chamfer_forward(xyz1, xyz2, workspace=None):
    """xyz1 [B,N,3], xyz2 [B,M,3] fp32 -> dist1 [B,N], dist2 [B,M], idx1 [B,N], idx2 [B,M] (int32).
    Large problems take the symmetric one-pass kernel (s3d_chamfer_forward_ws: every pair evaluated once for both
    directions), which needs a device workspace of s3d_chamfer_workspace_bytes(B, N, M) bytes; it is allocated here
    unless the caller passes one (uint8 tensor)."""
    _chk(xyz1, xyz2)
    assert xyz1.dtype == torch.float32 and xyz2.dtype == torch.float32
    B, N, three = xyz1.shape
    B2, M, three2 = xyz2.shape
    if three != 3 or three2 != 3 or B != B2:
        raise ValueError('chamfer: expected [B,N,3] and [B,M,3]')
    dev = xyz1.device
    dist1 = torch.empty((B, N), dtype=torch.float32, device=dev)
    dist2 = torch.empty((B, M), dtype=torch.float32, device=dev)
    idx1 = torch.empty((B, N), dtype=torch.int32, device=dev)
    idx2 = torch.empty((B, M), dtype=torch.int32, device=dev)
    L = _lib.load()
    need = int(L.s3d_chamfer_workspace_bytes(B, N, M)) if min(B, N, M) > 0 else 0
    if need and (workspace is None or workspace.numel() * workspace.element_size() < need):
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    rc = L.s3d_chamfer_forward_ws(xyz1.data_ptr(), xyz2.data_ptr(), dist1.data_ptr(), idx1.data_ptr(),
                                  dist2.data_ptr(), idx2.data_ptr(), B, N, M,
                                  workspace.data_ptr() if need else None, need, _stream())
    _lib.check(rc, 's3d_chamfer_forward_ws')
    _lib.count_launch(2)
    return dist1, dist2, idx1, idx2
