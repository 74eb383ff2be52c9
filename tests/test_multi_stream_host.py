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


def _fake_sysfs(root, gpus, node_cpus):
    """gpus: {pci id: numa node}; node_cpus: {node: cpulist text}."""
    for pci, node in gpus.items():
        d = root / 'bus' / 'pci' / 'devices' / pci
        d.mkdir(parents=True)
        (d / 'numa_node').write_text(f'{node}\n')
        (d / 'vendor').write_text('0x10de\n')
        (d / 'class').write_text('0x030200\n')
    for node, cpus in node_cpus.items():
        d = root / 'devices' / 'system' / 'node' / f'node{node}'
        d.mkdir(parents=True)
        (d / 'cpulist').write_text(cpus + '\n')


def test_bind_to_gpu_numa_deals_the_node_cores_to_its_ranks(tmp_path):
    """One process per GPU stays on the host cores of its GPU's NUMA node; unknown topology leaves the affinity alone."""
    from multi_stream import _parse_cpulist, bind_to_gpu_numa
    assert _parse_cpulist('0-3,8,10-11') == {0, 1, 2, 3, 8, 10, 11}
    full = os.sched_getaffinity(0)
    cpus = sorted(full)
    try:
        if len(cpus) >= 4:
            half = len(cpus) // 2
            lo, hi = cpus[:half], cpus[half:]
            gpus = {f'0000:{0x10 + i:02x}:00.0': (0 if i < 2 else 1) for i in range(4)}
            _fake_sysfs(tmp_path, gpus, {0: ','.join(map(str, lo)), 1: ','.join(map(str, hi))})
            got = []
            for r, pci in enumerate(sorted(gpus)):
                os.sched_setaffinity(0, full)
                got.append(bind_to_gpu_numa(pci, r, 4, sysfs=str(tmp_path)))
                assert os.sched_getaffinity(0) == got[-1]
            assert got[0] | got[1] <= set(lo) and got[2] | got[3] <= set(hi)
            assert not (got[0] & got[1]) and not (got[2] & got[3])
            os.sched_setaffinity(0, full)
            # upper-case bus id as cudaDeviceGetPCIBusId prints it; single rank gets the whole node
            assert bind_to_gpu_numa('0000:1A:00.0'.replace('1A', '10'), 0, 1, sysfs=str(tmp_path)) == set(lo)
        os.sched_setaffinity(0, full)
        assert bind_to_gpu_numa('0000:ff:00.0', 0, 1, sysfs=str(tmp_path)) is None      # no such device
        (tmp_path / 'bus' / 'pci' / 'devices' / '0000:fe:00.0').mkdir(parents=True)
        (tmp_path / 'bus' / 'pci' / 'devices' / '0000:fe:00.0' / 'numa_node').write_text('-1\n')
        assert bind_to_gpu_numa('0000:fe:00.0', 0, 1, sysfs=str(tmp_path)) is None      # single-node box
        assert os.sched_getaffinity(0) == full
        if len(cpus) >= 4:              # unknown topology, several ranks on the box: contiguous, disjoint core slices
            a = bind_to_gpu_numa('0000:fe:00.0', 0, 2, sysfs=str(tmp_path))
            os.sched_setaffinity(0, full)
            b = bind_to_gpu_numa('0000:fe:00.0', 1, 2, sysfs=str(tmp_path))
            assert a and b and not (a & b) and (a | b) <= full and max(a) < min(b)
    finally:
        os.sched_setaffinity(0, full)
