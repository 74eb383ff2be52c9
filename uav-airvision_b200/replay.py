"""Deterministic synchronous replay of ONE stream.

Replaces the reference's wall-clock paced DataPublisher + VIO threads (src/streaming/publisher.py:32-53,
src/modules/vio.py:26-53): before each stereo frame every IMU message with timestamp <= the frame's is delivered, in
order, to `imu_callback`; then `stereo_callback` runs inline, and an optional estimator (`msckf.MSCKF`, ours or the
reference's) consumes the feature message.  Imports nothing of the front end, so the same driver feeds the unmodified
reference (tools/make_golden.py), the oracle port and the CUDA front end.  `multi_stream.replay` is the S-stream form.
"""
from __future__ import annotations


def run_stream(front_end, stream, on_frame=None, msckf=None):
    """Returns the list of feature_msg (one per stereo frame)."""
    out = []
    for kind, msg in stream.events():
        if kind == 'imu':
            front_end.imu_callback(msg)
            if msckf is not None:
                msckf.imu_callback(msg)
        else:
            fm = front_end.stereo_callback(msg)
            out.append(fm)
            if on_frame is not None:
                on_frame(len(out) - 1, msg, fm)
            if msckf is not None:
                msckf.feature_callback(fm)
    return out
