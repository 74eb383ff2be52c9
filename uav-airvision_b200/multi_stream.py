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

from collections import defaultdict, namedtuple

import numpy as np

from image_processing import FeatureMeasurement, IMUProcessor, _avbhost, _native

feature_msg = namedtuple('feature_msg', ['timestamp', 'features'])


def shard_streams(n_streams: int, world: int, rank: int) -> list:
    """Stream indices owned by `rank`: s mod world == rank."""
    if not (0 <= rank < world):
        raise ValueError('rank out of range')
    return list(range(rank, n_streams, world))


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
