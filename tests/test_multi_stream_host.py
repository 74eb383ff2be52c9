"""Host-side logic of the multi-GPU path on CPU: stream sharding, and the post-run gather / max-over-ranks over
torch.distributed with the gloo backend, world_size 2 (streams shard, nothing else does)."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_streams_partitions_every_stream_once():
    from multi_stream import shard_streams
    for n, world in ((64, 8), (63, 4), (7, 2), (3, 8), (0, 2)):
        parts = [shard_streams(n, world, r) for r in range(world)]
        flat = sorted(s for p in parts for s in p)
        assert flat == list(range(n))
        assert all(s % world == r for r, p in enumerate(parts) for s in p)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard_streams(4, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'uav-airvision_b200'))
    import torch.distributed as dist
    from collections import namedtuple
    from multi_stream import gather_stats, max_over_ranks, shard_streams, stream_stats
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    fm = namedtuple('feature_msg', ['timestamp', 'features'])
    mine = shard_streams(7, world, rank)
    local = [stream_stats(s, [fm(10.0 + k, [None] * (s + k)) for k in range(3)], seconds=0.1 * (rank + 1)) for s in mine]
    dist.barrier()
    slowest = max_over_ranks(0.1 * (rank + 1), dist)
    allstats = gather_stats(local, dist)
    q.put((rank, slowest, allstats))
    dist.destroy_process_group()


def test_gloo_world2_gather_and_max_over_ranks():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, slowest, stats = q.get(timeout=120)
        got[rank] = (slowest, stats)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][0] == pytest.approx(0.2) and got[1][0] == pytest.approx(0.2)
    assert got[1][1] is None
    stats = got[0][1]
    assert [d['stream'] for d in stats] == list(range(7))
    assert all(d['frames'] == 3 and d['features'] == 3 * d['stream'] + 3 for d in stats)
