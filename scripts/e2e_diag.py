"""Where the e2e step loses time against the resident-input step: the same forward timed with (a) fp32 NCHW inputs resident in HBM
(bench.py's `value`), (b) uint8 HWC inputs resident, (c) uint8 HWC inputs + a concurrent H2D stream that is NOT consumed, (d) (b) +
the D2H of the results.  usage: e2e_diag.py [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from config import cfg
from stereo_3d_reconstruction_b200 import models
from stereo_3d_reconstruction_b200.utils import synthetic

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device('cuda', 0)
B, H, W, D = 64, cfg.CONST.IMG_H, cfg.CONST.IMG_W, cfg.NETWORK.MAX_DISP
cfg.NETWORK.PRECISION = 'bf16'
model = models.build_model('Stereo2Voxel', cfg, seed=0).to(dev).pack()
f32, u8, host = [], [], []
for i in range(4):
    l, r, _ = synthetic.stereo_pair(B, H, W, 2 * D, seed=i)
    g = synthetic.gt_volume(B, cfg.CONST.N_VOX, seed=50 + i)
    l8 = (l.clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    r8 = (r.clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    f32.append((l.to(dev), r.to(dev), g.to(dev)))
    u8.append((l8.to(dev), r8.to(dev), g.to(dev)))
    host.append((l8.pin_memory(), r8.pin_memory(), g.pin_memory()))
scratch = [torch.empty_like(t) for t in u8[0]]
copy_stream = torch.cuda.Stream(device=dev)
h_vox = torch.empty((B, 32, 32, 32), dtype=torch.float32).pin_memory()


def timed(fn):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def a(i):
    model(*f32[i % 4])


def b(i):
    model(*u8[i % 4])


def c(i):
    with torch.cuda.stream(copy_stream):
        for d, s in zip(scratch, host[i % 4]):
            d.copy_(s, non_blocking=True)
    model(*u8[i % 4])


def d(i):
    _, _, vox, _ = model(*u8[i % 4])
    h_vox.copy_(vox, non_blocking=True)


with torch.no_grad():
    for name, fn in (('fp32 resident', a), ('uint8 resident', b), ('uint8 resident + idle H2D stream', c), ('uint8 resident + D2H', d),
                     ('fp32 resident', a), ('uint8 resident', b)):
        print('%-36s %.3f ms/step' % (name, timed(fn)), flush=True)
