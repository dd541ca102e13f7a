# -*- coding: utf-8 -*-
"""Test driver: loop batches -> model forward -> IoU (Stereo2Voxel) / Chamfer (Stereo2Point)
statistics, sharded across ranks with ONE all-reduce of a small stats vector at the end
(the reference's `runner.py --test` path, README.md:91 -- its core/test.py is not
on disk).  Statistics are integer sums so that an N-GPU run is bit-identical to the 1-GPU run:
    per threshold t:  sum_b intersection(b,t), sum_b union(b,t)   (int64)
    n_samples                                                    (int64)
Chamfer: sum_b CD(b) is accumulated in fp64 and reduced as a fixed-point int64 (2^-40 units).
"""
import torch
import torch.distributed as dist

from ..utils import synthetic

_FX = float(1 << 40)


def shard_range(n_total, rank, world):
    """Contiguous shard [lo, hi) of n_total samples for `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_stats(stats):
    """SUM all-reduce of an int64 stats tensor over the default process group (no-op without one)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def iou_summary(stats, thresholds):
    """stats int64 [2T+1] -> dict.  IoU(t) = sum intersection / sum union."""
    T = len(thresholds)
    inter, union, n = stats[:T].double(), stats[T:2 * T].double(), int(stats[2 * T].item())
    return {'n_samples': n, 'thresholds': list(thresholds),
            'iou': [(i / u).item() if u > 0 else 0.0 for i, u in zip(inter, union)],
            'intersection': [int(v) for v in stats[:T].tolist()], 'union': [int(v) for v in stats[T:2 * T].tolist()]}


def test_voxel(cfg, model, n_samples, batch_size, rank=0, world=1, device='cuda'):
    """Synthetic StereoShapeNet stand-in: sample i is generated from seed i, so every rank can
    materialise exactly its own shard and the union over ranks is independent of `world`."""
    th = list(cfg.TEST.VOXEL_THRESH)
    T = len(th)
    stats = torch.zeros(2 * T + 1, dtype=torch.int64, device=device)
    lo, hi = shard_range(n_samples, rank, world)
    H, W = cfg.CONST.IMG_H, cfg.CONST.IMG_W
    with torch.no_grad():
        for s in range(lo, hi, batch_size):
            ids = range(s, min(s + batch_size, hi))
            pairs = [synthetic.stereo_pair(1, H, W, 2 * cfg.NETWORK.MAX_DISP, seed=i) for i in ids]
            left = torch.cat([p[0] for p in pairs]).to(device)
            right = torch.cat([p[1] for p in pairs]).to(device)
            gt = torch.cat([synthetic.gt_volume(1, cfg.CONST.N_VOX, seed=100000 + i) for i in ids]).to(device)
            _, _, _, iou = model(left, right, gt)
            stats[:T] += iou[:, :, 0].sum(0)
            stats[T:2 * T] += iou[:, :, 1].sum(0)
            stats[2 * T] += len(ids)
    return reduce_stats(stats)


def test_point(cfg, model, n_samples, batch_size, rank=0, world=1, device='cuda'):
    from ..extensions.chamfer_dist import chamfer_per_sample
    stats = torch.zeros(2, dtype=torch.int64, device=device)
    lo, hi = shard_range(n_samples, rank, world)
    H, W = cfg.CONST.IMG_H, cfg.CONST.IMG_W
    with torch.no_grad():
        for s in range(lo, hi, batch_size):
            ids = range(s, min(s + batch_size, hi))
            pairs = [synthetic.stereo_pair(1, H, W, 2 * cfg.NETWORK.MAX_DISP, seed=i) for i in ids]
            left = torch.cat([p[0] for p in pairs]).to(device)
            right = torch.cat([p[1] for p in pairs]).to(device)
            gt = torch.cat([synthetic.point_clouds(1, 1, cfg.CONST.N_GT_POINTS, seed=200000 + i)[1] for i in ids]).to(device)
            _, _, pts = model(left, right)
            cd = chamfer_per_sample(pts.contiguous(), gt)
            stats[0] += torch.round(cd.sum() * _FX).to(torch.int64)
            stats[1] += len(ids)
    return reduce_stats(stats)


def chamfer_summary(stats):
    n = int(stats[1].item())
    return {'n_samples': n, 'chamfer_distance': (stats[0].item() / _FX) / max(n, 1)}
