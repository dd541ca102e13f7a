"""Generates tests/golden/*.npz from THIS repo's oracle (oracle/models.py, oracle/chamfer_ref.c).

The reference has no source, tests or golden vectors on disk (/root/reference = README.md +
requirements.txt), so these fixtures do NOT pin the oracle to upstream -- "parity unpinned".  They
freeze the oracle's own behaviour (layer table, synthetic-weight recipe, arithmetic order) so a
later edit cannot silently move the parity target, and they travel to the GPU box where the CUDA
path is compared against them without needing the oracle's code path to be identical.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import models as O          # noqa: E402
from oracle import chamfer as OC        # noqa: E402
from stereo_3d_reconstruction_b200.utils import synthetic   # noqa: E402
from tests.common import small_cfg      # noqa: E402


def main():
    torch.set_num_threads(1)            # deterministic summation order
    for cv in ('concat', 'corr'):
        cfg = small_cfg(NETWORK__COST_VOLUME=cv)
        m = O.make_model('Stereo2Voxel', cfg, seed=0)
        left, right, _ = synthetic.stereo_pair(2, 64, 64, 16, seed=0)
        with torch.no_grad():
            dl, dr, vox = m(left, right)
        gt = synthetic.gt_volume(2)
        iou = O.iou_counts(vox, gt, cfg.TEST.VOXEL_THRESH)
        np.savez_compressed(os.path.join(HERE, 'stereo2voxel_small_%s.npz' % cv),
                            disp_left=dl.numpy().astype(np.float16), disp_right=dr.numpy().astype(np.float16),
                            voxels=vox.numpy().astype(np.float16), iou=iou.numpy())
    cfg = small_cfg()
    m = O.make_model('Stereo2Point', cfg, seed=0)
    left, right, _ = synthetic.stereo_pair(2, 64, 64, 16, seed=0)
    with torch.no_grad():
        _, _, pts = m(left, right)
    np.savez_compressed(os.path.join(HERE, 'stereo2point_small.npz'), points=pts.numpy())
    a, b = synthetic.point_clouds(2, 300, 1000, seed=5, duplicates=True)
    d1, d2, i1, i2 = OC.chamfer_c(a.numpy(), b.numpy())
    np.savez_compressed(os.path.join(HERE, 'chamfer_300x1000.npz'), dist1=d1, dist2=d2, idx1=i1, idx2=i2)
    print('wrote', sorted(f for f in os.listdir(HERE) if f.endswith('.npz')))


if __name__ == '__main__':
    main()
