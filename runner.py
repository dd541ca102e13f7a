#!/usr/bin/env python3
# -*- coding: utf-8 -*-
"""CLI entry, same contract as the reference's runner.py (README.md:85,91):

    python3 runner.py --test --weights=/path/to/pretrained/model.pth

Only the inference (`--test`) path is built (training is out of scope).
StereoShapeNet and the pretrained .pth files are unavailable offline, so without --weights the
model gets seeded synthetic weights and the test set is synthetic (config.py shapes).
Multi-GPU: launch with torchrun (one process per GPU); shards are contiguous by rank and the
IoU / Chamfer statistics are reduced with one NCCL all-reduce.
"""
import argparse
import json
import os
import sys

import torch


def get_args():
    p = argparse.ArgumentParser(description='Stereo 3D reconstruction (B200 inference path)')
    p.add_argument('--test', action='store_true', help='run the test (inference) path')
    p.add_argument('--weights', default=None, help='checkpoint (.pth) with a state_dict or {"model": state_dict}')
    p.add_argument('--model', default='Stereo2Voxel', choices=['Stereo2Voxel', 'Stereo2Point'])
    p.add_argument('--precision', default=None, choices=['bf16', 'tf32', 'tf32x3', 'fp32'])
    p.add_argument('--batch-size', type=int, default=None)
    p.add_argument('--n-samples', type=int, default=None, help='synthetic test-set size (default: one batch per rank)')
    return p.parse_args()


def main():
    args = get_args()
    from config import cfg
    if not args.test:
        sys.exit('Only `runner.py --test` is implemented: this is the inference-tier build (training is out of scope).')
    if not torch.cuda.is_available():
        sys.exit('runner.py --test needs a CUDA device (sm_100a); there is no CPU fallback.')
    from stereo_3d_reconstruction_b200 import models
    from stereo_3d_reconstruction_b200.core import test as T
    import torch.distributed as dist
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    if args.precision:
        cfg.NETWORK.PRECISION = args.precision
    bs = args.batch_size or cfg.CONST.BATCH_SIZE
    n = args.n_samples or bs * world
    model = models.build_model(args.model, cfg, seed=None if args.weights else cfg.CONST.SEED)
    if args.weights:
        sd = torch.load(args.weights, map_location='cpu')
        model.load_state_dict(sd.get('model', sd) if isinstance(sd, dict) and 'model' in sd else sd)
    model.cuda().pack()
    if args.model == 'Stereo2Voxel':
        res = T.iou_summary(T.test_voxel(cfg, model, n, bs, rank, world).cpu(), cfg.TEST.VOXEL_THRESH)
    else:
        res = T.chamfer_summary(T.test_point(cfg, model, n, bs, rank, world).cpu())
    if rank == 0:
        res.update(model=args.model, precision=cfg.NETWORK.PRECISION, world_size=world, data='synthetic')
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
