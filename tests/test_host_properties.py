"""Property tests (hypothesis) of the host logic: tile chooser, sharding, and PackedConv tap tables vs
torch.nn.functional on random small shapes (CPU emulation of the S3dConvParams contract)."""
import pytest
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


def test_nstack_weights_layout():
    """weight_ns[r*9+kyx] block s = W[kz = (r+1-s) mod 3] (rotation 3 = rotation 0 with block 2 zeroed): include/s3d.h."""
    torch.manual_seed(0)
    conv = nn.Conv3d(16, 24, 3, 1, 1, bias=False)
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    co = pc.cout_pad
    ns = pc.weight_ns.view(4, 9, 3, co, pc.cin_pad)
    w = pc.weight.view(3, 9, co, pc.cin_pad)
    for r in range(4):
        for s_ in range(3):
            want = torch.zeros_like(w[0]) if (r == 3 and s_ == 2) else w[((r % 3) + 1 - s_) % 3]
            assert torch.equal(ns[r, :, s_], want)
    conv2 = nn.Conv3d(16, 128, 3, 1, 1)
    assert PackedConv.from_conv(conv2, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu').weight_ns is None


def test_model_repacks_when_its_parameters_change():
    """load_state_dict / .to() / init_synthetic_weights drop the folded weights, workspaces and graphs."""
    from stereo_3d_reconstruction_b200 import models
    from tests.common import small_cfg
    m = models.build_model('Stereo2Voxel', small_cfg(), seed=0)
    m._packed, m._ws, m._graphs = {'stale': 1}, {'x': 1}, {'g': 1}
    m.load_state_dict(m.state_dict())
    assert m._packed is None and not m._ws and not m._graphs
    m._packed = {'stale': 1}
    m.float()
    assert m._packed is None
    m._packed = {'stale': 1}
    models.init_synthetic_weights(m, 1)
    assert m._packed is None
    c = small_cfg(NETWORK__DEC_CHANNELS=[24, 32, 16, 16, 8])
    with pytest.raises(ValueError):
        models.build_model('Stereo2Voxel', c, seed=0).pack()


def test_runner_flattens_checkpoints(tmp_path):
    """runner.py --weights: DataParallel prefixes, {'model': sd}, one state_dict per sub-network; clear error otherwise."""
    import runner
    from stereo_3d_reconstruction_b200 import models
    from tests.common import small_cfg
    m = models.build_model('Stereo2Voxel', small_cfg(), seed=0)
    sd = m.state_dict()
    ck = {'epoch_idx': 3}
    for sub in ('dispnet', 'rgbd_encoder', 'decoder', 'merger'):
        ck[sub + '_state_dict'] = {'module.' + k[len(sub) + 1:]: v for k, v in sd.items() if k.startswith(sub + '.')}
    assert set(runner.flatten_checkpoint(ck)) == set(sd)
    assert set(runner.flatten_checkpoint({'model': {'module.' + k: v for k, v in sd.items()}})) == set(sd)
    path = tmp_path / 'w.pth'
    torch.save(ck, path)
    runner.load_checkpoint(m, str(path))
    bad = dict(sd)
    bad.pop(next(iter(bad)))
    bad['encoder.upstream_name.weight'] = torch.zeros(1)
    torch.save(bad, path)
    with pytest.raises(SystemExit) as e:
        runner.load_checkpoint(m, str(path))
    assert 'missing' in str(e.value) and 'unexpected' in str(e.value)


def test_split_storage_and_packing():
    """'bf16x3' storage (S3D_DTYPE_BF16X2): v = hi + lo to 2^-16; packed weights are [.., hi(Cin_pad) | lo(Cin_pad)]."""
    from stereo_3d_reconstruction_b200.layers import to_storage, from_storage
    torch.manual_seed(0)
    x = torch.randn(4, 7, 24) * 5
    s = to_storage(x, lib.DTYPE_BF16X2)
    assert s.dtype == torch.bfloat16 and s.shape == (4, 7, 48)
    assert torch.equal(s[..., :24], x.to(torch.bfloat16))
    assert (from_storage(s, lib.DTYPE_BF16X2) - x).abs().max() <= 2.0 ** -16 * x.abs().max()
    conv = nn.Conv3d(16, 24, 3, 1, 1, bias=False)
    pc = PackedConv.from_conv(conv, None, lib.ACT_NONE, lib.DTYPE_BF16X2, 'cpu')
    ref = PackedConv.from_conv(conv, None, lib.ACT_NONE, lib.DTYPE_F32, 'cpu')
    assert pc.weight.shape == (27, pc.cout_pad, 2 * pc.cin_pad) and pc.weight_ns.shape == (36, 3 * pc.cout_pad, 2 * pc.cin_pad)
    torch.testing.assert_close(from_storage(pc.weight, lib.DTYPE_BF16X2), ref.weight, rtol=2.0 ** -15, atol=1e-7)
    torch.testing.assert_close(from_storage(pc.weight_ns, lib.DTYPE_BF16X2), ref.weight_ns, rtol=2.0 ** -15, atol=1e-7)
    conv2 = nn.Conv2d(32, 32, 3, 1, 1)
    assert PackedConv.from_conv(conv2, None, lib.ACT_RELU, lib.DTYPE_BF16X2, 'cpu').vol is not None     # 2-D layers as volumes
