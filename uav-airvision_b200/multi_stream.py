"""Multi-stream front end and sweep driver (SURVEY.md section 8 rows e and f2).

The reference runs one stream per process (`python main.py --path <seq> --offset <s>`, src/main.py:10-34) and sweeps
sequences x time offsets with a batch file (run.bat:4-12).  Streams share nothing, so here

  * `MultiStreamFrontEnd` drives S independent streams in lock-step through ONE libavb context: every kernel launch
    of the frame chain covers all S streams (blockIdx = stream), one H2D / D2H per step;
  * `shard_streams` assigns stream s to GPU s mod G -- one process per GPU, no collective on the data path;
  * `replay` is the deterministic synchronous driver that replaces the reference's wall-clock paced DataPublisher +
    VIO threads (src/streaming/publisher.py:32-53, src/modules/vio.py:26-53): before each stereo frame of a stream,
    every IMU message with timestamp <= the frame's is delivered in order;
  * `gather_stats` collects per-stream summaries on rank 0 with torch.distributed AFTER the timed work (gloo or nccl).
"""
from __future__ import annotations

import os
from collections import defaultdict, namedtuple

import numpy as np

from image_processing import FeatureMeasurement, IMUProcessor, _avbhost, _native

feature_msg = namedtuple('feature_msg', ['timestamp', 'features'])


def shard_streams(n_streams: int, world: int, rank: int) -> list:
    """Stream indices owned by `rank`: s mod world == rank."""
    if not (0 <= rank < world):
        raise ValueError('rank out of range')
    return list(range(rank, n_streams, world))


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa(pci_bus_id: str, local_rank: int = 0, local_world: int = 1, sysfs: str = '/sys'):
    """One process per GPU: keep the process (and with it the first-touch placement of its page-locked staging and
    result blocks) on the host cores of the NUMA node the GPU hangs off, so the per-frame H2D copy and the result
    block the last kernel writes into mapped host memory do not cross the socket interconnect.  Call it BEFORE the
    context is created.  `pci_bus_id` is 'dddd:bb:dd.f' (cudaDeviceGetPCIBusId).  The node's allowed cores are dealt
    to the ranks that share the node in contiguous slices.  When the topology is unknown (no sysfs entry, node -1 as on
    single-node or virtualised boxes, no allowed core on the node) and several ranks share the box, the allowed cores
    are dealt to the local ranks in contiguous slices instead, so that the ranks at least do not migrate over each
    other; a single rank is left untouched.  Returns the cpu set bound to, or None when the affinity was not changed."""
    def core_slice():
        allowed = sorted(os.sched_getaffinity(0))
        if local_world <= 1 or len(allowed) < local_world:
            return None
        share = len(allowed) // local_world
        mine = allowed[local_rank * share:(local_rank + 1) * share]
        os.sched_setaffinity(0, mine)
        return set(mine)
    try:
        dev = os.path.join(sysfs, 'bus', 'pci', 'devices', pci_bus_id.lower())
        node = int(open(os.path.join(dev, 'numa_node')).read())
        if node < 0:
            return core_slice()
        cpus = _parse_cpulist(open(os.path.join(sysfs, 'devices', 'system', 'node', f'node{node}', 'cpulist')).read())
        allowed = sorted(cpus & os.sched_getaffinity(0))
        if not allowed:
            return core_slice()
        # ranks sharing this node: assume GPUs are spread evenly over the nodes that have any
        nodes = set()
        for d in os.listdir(os.path.join(sysfs, 'bus', 'pci', 'devices')):
            try:
                base = os.path.join(sysfs, 'bus', 'pci', 'devices', d)
                if open(os.path.join(base, 'vendor')).read().strip() == '0x10de' and \
                        open(os.path.join(base, 'class')).read().strip().startswith('0x0302'):
                    nodes.add(int(open(os.path.join(base, 'numa_node')).read()))
            except (OSError, ValueError):
                continue
        per_node = max(1, -(-local_world // max(1, len(nodes))))
        slot = local_rank % per_node
        share = max(1, len(allowed) // per_node)
        mine = allowed[slot * share:(slot + 1) * share] or allowed
        os.sched_setaffinity(0, mine)
        return set(mine)
    except (OSError, ValueError):
        return core_slice()


class MultiStreamFrontEnd:
    """S lock-stepped streams; per stream the same state an ImageProcessingPipeline keeps (IMU buffer, previous
    message, feature-id counter, per-stage counters)."""

    def __init__(self, config, width, height, n_streams, device=0, use_graph=True):
        self.config, self.S = config, int(n_streams)
        self.ctx = _native.Context(config, width, height, num_streams=self.S, device=device, use_graph=use_graph)
        self.imu = [IMUProcessor(config.T_imu_cam0, config.T_imu_cam1) for _ in range(self.S)]
        self.prev_msg = [None] * self.S
        self.first_frame = True
        self.next_feature_id = [0] * self.S
        self.num_features = [defaultdict(int) for _ in range(self.S)]
        self._R = np.empty((2, self.S, 3, 3))          # [cam][stream]; cam 1 is passed on only with RANSAC on
        self._steps_prepared, self._prepared, self._in_flight, self._done = 0, None, None, None   # store-fed steps (sweep.py)

    def close(self):
        self.ctx.close()

    def imu_callback(self, s, imu_msg):
        self.imu[s].imu_callback(imu_msg)

    def stereo_callback(self, stereo_msgs):
        """One stereo frame per stream (list of S stereo_msg) -> list of S feature_msg."""
        if len(stereo_msgs) != self.S:
            raise ValueError(f'expected {self.S} stereo messages')
        R = None
        if not self.first_frame:
            for s, m in enumerate(stereo_msgs):
                imu = self.imu[s]
                imu.cam0_prev_img_msg, imu.cam0_curr_img_msg = self.prev_msg[s], m.cam0_msg
                self._R[0, s], self._R[1, s] = imu.integrate_imu_data()
            R = self._R if self.ctx.ransac else self._R[0]
        res = _avbhost.process_frames(self.ctx._h.value, [m.cam0_msg.image for m in stereo_msgs],
                                      [m.cam1_msg.image for m in stereo_msgs], R, FeatureMeasurement)
        out = []
        for s, ((feats, hdr), m) in enumerate(zip(res, stereo_msgs)):
            self.next_feature_id[s] = hdr[1]
            if not self.first_frame:
                nf = self.num_features[s]
                nf['before_tracking'] = hdr[2]
                if hdr[2]:
                    nf['after_tracking'], nf['after_matching'], nf['after_ransac'] = hdr[3], hdr[4], hdr[5]
            self.prev_msg[s] = m.cam0_msg
            out.append(feature_msg(m.cam0_msg.timestamp, feats))
        self.first_frame = False
        return out

    def step_from_store(self, image_addrs, cam0_msgs):
        """One frame per stream from HBM-resident images (FrameStore.addr rows: uint64[S, 2]); cam0_msgs[s] carries the
        frame's timestamp (the IMU window reads nothing else).  Returns per stream (timestamp, ids int64[n], meas
        float64[n, 4]) -- copies of the result block, no per-feature Python objects: the form a sweep hands to its
        estimator processes.  = prepare_step + begin_step_from_store + end_step_from_store, which a driver may
        interleave so that the host work of step k+1 runs while the GPU works on step k."""
        self.prepare_step(cam0_msgs)
        self.begin_step_from_store(image_addrs)
        return self.end_step_from_store()

    def prepare_step(self, cam0_msgs):
        """Host half of a step: the IMU window of every stream (imu_processor.py:28-67) -> the camera rotations of the
        step.  Touches host state only, so it may run while the previous step is still on the GPU."""
        if len(cam0_msgs) != self.S:
            raise ValueError(f'expected {self.S} frames')
        R = None
        if self._steps_prepared > 0:
            R = np.empty((2, self.S, 3, 3))
            for s, m in enumerate(cam0_msgs):
                imu = self.imu[s]
                imu.cam0_prev_img_msg, imu.cam0_curr_img_msg = self.prev_msg[s], m
                R[0, s], R[1, s] = imu.integrate_imu_data()
        for s, m in enumerate(cam0_msgs):
            self.prev_msg[s] = m
        self._steps_prepared += 1
        self._prepared = (list(cam0_msgs), R)

    def begin_step_from_store(self, image_addrs):
        """Enqueues gather + frame chain of the prepared step; returns at once."""
        msgs, R = self._prepared
        self._prepared = None
        R0, R1 = (None, None) if R is None else (R[0], R[1] if self.ctx.ransac else None)
        self.ctx.process_gather(image_addrs, R0, R1, wait=False)
        self._in_flight = msgs

    def end_step_from_store(self):
        """Waits for the step begun last and returns its results (see step_from_store)."""
        self.wait_step()
        return self.take_results(prev=False)

    def wait_step(self):
        """Blocks until the step begun last has completed (its result block is in host memory)."""
        self._done, self._in_flight = self._in_flight, None
        self.ctx.sync()

    def take_results(self, prev):
        """Results of the step completed last.  `prev=True`: the NEXT step has already been begun (its chain runs while
        this reads): libavb keeps the result blocks of two consecutive frames (avb_get_result_prev).  The S result
        blocks leave the pinned memory in ONE copy; the per-stream arrays are views into that copy."""
        msgs, self._done = self._done, None
        blk, ids_off, meas_off = self.ctx.result_block(prev=prev)
        cap = self.ctx.capacity
        blk = blk.copy()
        hdr = blk[:, :_native.HEADER_DTYPE.itemsize].view(_native.HEADER_DTYPE)[:, 0]
        ids = blk[:, ids_off:ids_off + 8 * cap].view(np.int64)
        meas = blk[:, meas_off:meas_off + 32 * cap].view(np.float64).reshape(self.S, cap, 4)
        n_all = hdr['n_features']
        next_id, before = hdr['next_feature_id'].tolist(), hdr['before_tracking'].tolist()
        counts = (hdr['after_tracking'].tolist(), hdr['after_matching'].tolist(), hdr['after_ransac'].tolist())
        out = []
        for s, m in enumerate(msgs):
            self.next_feature_id[s] = next_id[s]
            if not self.first_frame:
                nf = self.num_features[s]
                nf['before_tracking'] = before[s]
                if before[s]:
                    nf['after_tracking'], nf['after_matching'], nf['after_ransac'] = counts[0][s], counts[1][s], counts[2][s]
            n = int(n_all[s])
            out.append((m.timestamp, ids[s, :n], meas[s, :n]))
        self.first_frame = False
        return out


def replay(front_end: MultiStreamFrontEnd, streams, on_frame=None, consumers=None):
    """Deterministic lock-step replay of S streams (objects with .events() yielding ('imu'|'stereo', msg)).
    consumers[s] (optional) is an estimator with imu_callback / feature_callback (the reference's MSCKF, unchanged).
    Stops when the shortest stream ends.  Returns per-stream lists of feature_msg."""
    its = [iter(st.events()) for st in streams]
    out = [[] for _ in streams]
    k = 0
    while True:
        frame = []
        for s, it in enumerate(its):
            msg = None
            for kind, m in it:
                if kind == 'imu':
                    front_end.imu_callback(s, m)
                    if consumers is not None:
                        consumers[s].imu_callback(m)
                else:
                    msg = m
                    break
            if msg is None:
                return out
            frame.append(msg)
        fms = front_end.stereo_callback(frame)
        for s, fm in enumerate(fms):
            out[s].append(fm)
            if consumers is not None:
                consumers[s].feature_callback(fm)
        if on_frame is not None:
            on_frame(k, frame, fms)
        k += 1


def stream_stats(stream_index, msgs, seconds=None):
    n = len(msgs)
    feats = int(sum(len(m.features) for m in msgs))
    d = {'stream': int(stream_index), 'frames': n, 'features': feats,
         'last_timestamp': float(msgs[-1].timestamp) if n else None}
    if seconds is not None:
        d['seconds'] = float(seconds)
    return d


def gather_stats(local_stats, dist=None):
    """Per-stream summaries of all ranks on rank 0 (None elsewhere), ordered by stream index.  Host-side, after the
    timed region: the data path itself has no collective."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return sorted(local_stats, key=lambda d: d['stream'])
    rank, world = dist.get_rank(), dist.get_world_size()
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local_stats, bucket, dst=0)
    if rank != 0:
        return None
    return sorted((d for part in bucket for d in part), key=lambda d: d['stream'])


def max_over_ranks(value, dist=None, device='cpu'):
    """Timing rule: a multi-rank step takes as long as its slowest rank."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
