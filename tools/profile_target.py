"""Minimal ncu target: one libavb context, S lock-stepped streams, F frames of the bench workload from device-resident
blocks.  Kernel launches per steady-state frame: see avb_kernels_per_frame.

    python tools/profile_target.py --streams 1 --frames 8 [--workload c2]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
    sys.path.insert(0, p)
sys.dont_write_bytecode = True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--streams', type=int, default=1)
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--workload', default='c2')
    a = ap.parse_args()
    import torch
    from bench import rotations_for, workload
    from image_processing import _native
    from synth_euroc import SlidingTextureStream
    cfg, skw, _ = workload(a.workload)
    S, F = a.streams, a.frames
    n = F + 2 * (S - 1)
    stream = SlidingTextureStream(n_frames=n, **skw)
    frames = [stream.frame(k) for k in range(n)]
    stream.frames = lambda: iter(frames)
    Rs = rotations_for(cfg, stream)
    ctx = _native.Context(cfg, stream.w, stream.h, num_streams=S, device=0, use_graph=True)
    bb, ib = ctx.block_bytes, stream.w * stream.h
    host = torch.zeros((F, bb), dtype=torch.uint8)
    hb = host.numpy()
    for k in range(F):
        for s in range(S):
            f = frames[k + 2 * s]
            hb[k, (2 * s) * ib:(2 * s + 1) * ib] = f.cam0_image.reshape(-1)
            hb[k, (2 * s + 1) * ib:(2 * s + 2) * ib] = f.cam1_image.reshape(-1)
        import numpy as np
        ctx.fill_rotations(hb[k], np.stack([Rs[k + 2 * s][0] for s in range(S)]),
                           np.stack([Rs[k + 2 * s][1] for s in range(S)]))
    dev = host.cuda()
    for k in range(F):
        ctx.process_device(dev.data_ptr() + k * bb)
    hdr, ids, _ = ctx.result(0)
    print('frames', F, 'streams', S, 'features in the last frame of stream 0:', int(hdr['n_features']))
    ctx.close()


if __name__ == '__main__':
    main()
