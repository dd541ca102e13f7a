# -*- coding: utf-8 -*-
"""`extensions.chamfer_dist` -- the import path of the reference's Chamfer extension (/root/reference/README.md:62-65).

    from extensions.chamfer_dist import ChamferDistance
    cd = ChamferDistance()(pred_points, gt_points)

The implementation lives in stereo_3d_reconstruction_b200/extensions/chamfer_dist (forward only; sm_100a kernel in
libs3d_b200.so).  After `python setup.py build_ext --inplace` (or `install`) in this directory the compiled torch C++
shim `chamfer` is used for the call; without it the same kernel is reached through the ctypes binding of the same
library.  Neither has a CPU path.
"""
from stereo_3d_reconstruction_b200.extensions.chamfer_dist import (ChamferDistance, ChamferFunction,  # noqa: F401
                                                                   chamfer_per_sample, backend)
