"""Chamfer forward at BASELINE configs[3] (32 x 2048 x 16384): the symmetric one-pass kernel with 8 and 4 resident points per
lane against one search per direction; CUDA events, L2 flushed between launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stereo_3d_reconstruction_b200 import lib, ops
from stereo_3d_reconstruction_b200.utils import synthetic
B, N, M = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (32, 2048, 16384)
a, b = synthetic.point_clouds(B, N, M, seed=2, device='cuda')
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device='cuda')
for name, sym, r in (('two searches', -1, 0), ('symmetric R=8', 1, 8), ('symmetric R=4', 1, 4), ('automatic', 0, 0)):
    lib.set_knob('chamfer_sym', sym); lib.set_knob('chamfer_sym_r', r)
    ms = bench._events_ms(lambda: ops.chamfer_forward(a, b), 10, flush=flush)
    print('%-16s %.4f ms' % (name, ms), flush=True)
