"""BASELINE configs[4]: cost-volume + soft-argmin sweep (max disparity 32/64/128 x feature channels 32/64 at
1/4 resolution) and configs[3]: Chamfer (B=32, 2048 vs 16384 points).  Prints one JSON object per line.

Algorithmic bytes: each unique input read once + each output written once.
  concat build       2*(2B)*C*h*w*e / 2  (both feature maps, read once)  +  2B*2C*D*h*w*e written
  fused corr+softarg 2B*C*h*w*e read + 2B*h*w*4 written
  standalone softarg N*D*h*w*4 read + N*h*w*4 written
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stereo_3d_reconstruction_b200 import ops

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else 6650.0
flush = torch.empty(256 * 2 ** 20, dtype=torch.uint8, device='cuda')


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                                  # evict L2 (256 MB > 126 MB)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


B, h, w = 64, 64, 64
for dt, e in ((torch.bfloat16, 2), (torch.float32, 4)):
    for C in (32, 64):
        feat = torch.randn(2 * B, 1, h, w, C, device='cuda').to(dt)
        for D in (32, 64, 128):
            rec = {'B': B, 'h': h, 'w': w, 'C': C, 'D': D, 'dtype': str(dt).split('.')[1]}
            vol_bytes = 2 * B * D * h * w * 2 * C * e
            if vol_bytes < 40e9:
                vol = torch.empty(2 * B, D, h, w, 2 * C, device='cuda', dtype=dt)
                ms = timeit(lambda: ops.cost_volume_concat(feat, B, D, out=vol), 5)
                by = 2 * B * C * h * w * e + vol_bytes
                rec['concat_ms'] = ms; rec['concat_GBs'] = by / ms / 1e6; rec['concat_frac_hbm'] = by / ms / 1e6 / PEAK
                del vol
            disp = torch.empty(2 * B, h, w, device='cuda')
            ms = timeit(lambda: ops.corr_soft_argmin(feat, B, D, out=disp))
            by = 2 * B * C * h * w * e + 2 * B * h * w * 4
            fl = 2.0 * 2 * B * C * D * h * w
            rec['corr_ms'] = ms; rec['corr_GBs'] = by / ms / 1e6; rec['corr_frac_hbm'] = by / ms / 1e6 / PEAK
            rec['corr_TFLOPs_simt'] = fl / ms / 1e9
            if dt == torch.float32 and C == 32:
                cost = torch.randn(2 * B, D, h, w, device='cuda')
                ms = timeit(lambda: ops.soft_argmin(cost, -1.0, out=disp))
                by = 2 * B * D * h * w * 4 + 2 * B * h * w * 4
                rec['softargmin_ms'] = ms; rec['softargmin_GBs'] = by / ms / 1e6; rec['softargmin_frac_hbm'] = by / ms / 1e6 / PEAK
            print(json.dumps(rec), flush=True)

# Chamfer, BASELINE configs[3]
from stereo_3d_reconstruction_b200.utils import synthetic
a, b = synthetic.point_clouds(32, 2048, 16384, seed=2, device='cuda')
ms = timeit(lambda: ops.chamfer_forward(a, b))
pairs = 2.0 * 32 * 2048 * 16384
print(json.dumps({'chamfer_B32_2048x16384_ms': ms, 'pair_evals_per_s': pairs / ms * 1e3,
                  'fp32_instr_per_pair': 11, 'simt_issue_frac_of_peak': pairs * 11 / (ms / 1e3) / (148 * 128 * 1.9e9)}))
