"""B = 1 / 8 latency (CUDA-graph replay) with the sheared first aggregation layer and with knob no_sheared."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from config import cfg
from stereo_3d_reconstruction_b200 import lib, models
dev = torch.device('cuda', 0)
cfg.NETWORK.PRECISION = 'bf16'
for knob in (0, 1, 0, 1):
    lib.set_knob('no_sheared', knob)
    model = models.build_model('Stereo2Voxel', cfg, seed=0).to(dev).pack()
    print('no_sheared =', knob, {k: round(v, 4) for k, v in bench.extra_latency(model, cfg, dev).items() if k.endswith('_ms')}, flush=True)
    del model
