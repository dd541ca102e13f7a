"""Per-kernel time breakdown of one Stereo2Voxel forward (CUDA events around every C-ABI call).
usage: layer_times.py [batch] [precision] [u8]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from config import cfg
from stereo_3d_reconstruction_b200 import models, ops, lib
from stereo_3d_reconstruction_b200.utils import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg.NETWORK.PRECISION = sys.argv[2] if len(sys.argv) > 2 else 'bf16'
cfg.CONST.MICRO_BATCH = max(B, 64)
model = models.build_model('Stereo2Voxel', cfg, seed=0).cuda().pack()
l, r, _ = synthetic.stereo_pair(B, cfg.CONST.IMG_H, cfg.CONST.IMG_W, 64, seed=0)
if len(sys.argv) > 3 and sys.argv[3] == 'u8':          # decoded 8-bit HWC inputs (the e2e path)
    l = (l.clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    r = (r.clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
l, r = l.cuda(), r.cuda()
gt = synthetic.gt_volume(B).cuda()
for _ in range(2):
    model(l, r, gt)
torch.cuda.synchronize()
events = []
def wrap(name, fn):
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(*a, **k); e1.record()
        events.append((name, e0, e1)); return out
    return f
oc = model._conv
model._conv = lambda name, x, **kw: wrap(name, oc)(name, x, **kw)
for n in ('pack_image', 'conv_first', 'conv_concat_volume', 'depth_to_space', 'cost_volume_concat', 'soft_argmin', 'tap_gather_soft_argmin', 'cls_soft_argmin', 'corr_soft_argmin', 'upsample_disp', 'latent_to_vox', 'fuse_views'):
    setattr(ops, n, wrap(n, getattr(ops, n)))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); model(l, r, gt); t1.record(); torch.cuda.synchronize()
tot = t0.elapsed_time(t1)
shapes = None
try:
    import bench
    shapes = bench.model_flop_shapes(cfg, B)
except Exception as e:
    print('no flops', e)
acc = 0.0
print('%-22s %9s %7s %10s' % ('kernel', 'ms', '%', 'TFLOP/s'))
for name, e0, e1 in events:
    ms = e0.elapsed_time(e1); acc += ms
    tf = ''
    if shapes and name in shapes and name in model._packed:
        tf = '%.0f' % (model._packed[name].flops(*shapes[name]) / ms / 1e9)
    print('%-22s %9.3f %6.1f%% %10s' % (name, ms, 100 * ms / tot, tf))
print('sum of kernels %.3f ms, forward %.3f ms, %.1f pairs/s' % (acc, tot, B / tot * 1e3))
