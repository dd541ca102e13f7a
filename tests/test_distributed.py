"""N>1 host logic on CPU: contiguous sharding + the single SUM all-reduce of the integer stats
vector (gloo, world_size 2) must reproduce the 1-process result bit for bit."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereo_3d_reconstruction_b200.core import test as T


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            rs = [T.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in rs) - min(hi - lo for lo, hi in rs) <= 1


def _fake_counts(i, T_):
    g = torch.Generator().manual_seed(1000 + i)
    inter = torch.randint(0, 2000, (T_,), generator=g)
    return inter, inter + torch.randint(1, 3000, (T_,), generator=g)


def _local_stats(n, rank, world, T_):
    s = torch.zeros(2 * T_ + 1, dtype=torch.int64)
    lo, hi = T.shard_range(n, rank, world)
    for i in range(lo, hi):
        a, u = _fake_counts(i, T_)
        s[:T_] += a; s[T_:2 * T_] += u; s[2 * T_] += 1
    return s


def _worker(rank, world, port, n, T_, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    s = T.reduce_stats(_local_stats(n, rank, world, T_))
    if rank == 0:
        q.put(s.tolist())
    dist.destroy_process_group()


def test_gloo_world2_matches_single_process():
    n, T_ = 37, 4
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0)); port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n, T_, q)) for r in range(2)]
    [p.start() for p in ps]
    got = q.get(timeout=120)
    [p.join(60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    ref = _local_stats(n, 0, 1, T_)
    assert got == ref.tolist()
    summ = T.iou_summary(torch.tensor(got), [0.2, 0.3, 0.4, 0.5])
    assert summ['n_samples'] == n and all(0 <= v <= 1 for v in summ['iou'])
