"""Secondary BASELINE configs: tf32 Stereo2Voxel (configs[1] 'fp32'), Stereo2Point + chamfer (configs[3]),
batch 512 on one GPU through micro-batches (configs[2] at N=1)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from config import cfg
from stereo_3d_reconstruction_b200 import models, ops
from stereo_3d_reconstruction_b200.extensions.chamfer_dist import chamfer_per_sample
from stereo_3d_reconstruction_b200.utils import synthetic


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


what = sys.argv[1] if len(sys.argv) > 1 else 'all'
H, W, D = cfg.CONST.IMG_H, cfg.CONST.IMG_W, cfg.NETWORK.MAX_DISP
if what in ('all', 'tf32'):
    c = cfg.clone(); c.NETWORK.PRECISION = 'tf32'
    m = models.build_model('Stereo2Voxel', c, seed=0).cuda().pack()
    l, r, _ = synthetic.stereo_pair(64, H, W, 2 * D, seed=0, device='cuda')
    ms = timed(lambda: m(l, r))
    print(json.dumps({'config': 'Stereo2Voxel fwd, batch 64, tf32 (fp32 storage, kind::tf32)', 'ms': ms, 'pairs_per_s': 64 / ms * 1e3}))
    del m
    torch.cuda.empty_cache()
if what in ('all', 'point'):
    c = cfg.clone(); c.NETWORK.PRECISION = 'bf16'
    m = models.build_model('Stereo2Point', c, seed=0).cuda().pack()
    l, r, _ = synthetic.stereo_pair(32, H, W, 2 * D, seed=0, device='cuda')
    _, gt = synthetic.point_clouds(32, 1, c.CONST.N_GT_POINTS, seed=3, device='cuda')

    def step():
        _, _, pts = m(l, r)
        return chamfer_per_sample(pts.contiguous(), gt)
    ms = timed(step)
    ms_ch = timed(lambda: chamfer_per_sample(m(l, r)[2].contiguous(), gt) if False else ops.chamfer_forward(torch.rand(32, 2048, 3, device='cuda'), gt))
    print(json.dumps({'config': 'Stereo2Point fwd + chamfer_dist, batch 32, 2048 vs 16384 pts, bf16', 'ms': ms,
                      'pairs_per_s': 32 / ms * 1e3, 'chamfer_only_ms': ms_ch}))
    del m
    torch.cuda.empty_cache()
if what in ('all', 'b512'):
    c = cfg.clone(); c.NETWORK.PRECISION = 'bf16'; c.CONST.MICRO_BATCH = 64
    m = models.build_model('Stereo2Voxel', c, seed=0).cuda().pack()
    l, r, _ = synthetic.stereo_pair(512, H, W, 2 * D, seed=0, device='cuda')
    gt = synthetic.gt_volume(512, device='cuda')
    ms = timed(lambda: m(l, r, gt), iters=2, warm=1)
    print(json.dumps({'config': 'Stereo2Voxel fwd, batch 512 on ONE GPU (8 micro-batches of 64), bf16', 'ms': ms,
                      'pairs_per_s': 512 / ms * 1e3, 'max_mem_GB': torch.cuda.max_memory_allocated() / 2 ** 30}))
