"""Eager vs CUDA-graph replay latency of the Stereo2Voxel forward at small batch sizes.  usage: graph_latency.py [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from config import cfg
from stereo_3d_reconstruction_b200 import models
from stereo_3d_reconstruction_b200.utils import synthetic

Bs = [int(a) for a in sys.argv[1:]] or [1, 2, 8, 64]
model = models.build_model('Stereo2Voxel', cfg, seed=0).cuda().pack()
for B in Bs:
    l, r, _ = synthetic.stereo_pair(B, cfg.CONST.IMG_H, cfg.CONST.IMG_W, 64, seed=B)
    l, r = l.cuda(), r.cuda()
    res = {}
    for mode, fn in (('eager', model), ('graph', model.graphed)):
        with torch.no_grad():
            for _ in range(5):
                fn(l, r)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 50 if B <= 8 else 10
            e0.record()
            for _ in range(n):
                fn(l, r)
            e1.record()
            torch.cuda.synchronize()
            res[mode] = e0.elapsed_time(e1) / n
    print('B=%3d  eager %.3f ms (%.0f pairs/s)   graph %.3f ms (%.0f pairs/s)' %
          (B, res['eager'], B / res['eager'] * 1e3, res['graph'], B / res['graph'] * 1e3))
