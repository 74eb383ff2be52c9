"""Probe (not a benchmark): the S lock-stepped runs of the C4 shape split into G contexts of S/G runs each, every context on
its own CUDA streams, steps enqueued round-robin: the latency tails of one group's step (second candidate round,
k_select, k_finish, small pyramid levels) overlap the bulk of the other groups'.  Inputs resident in HBM.

    python tools/c4_groups.py --streams 64 --groups 1 2 4 --steps 20
"""
from __future__ import annotations

import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
    sys.path.insert(0, p)
sys.dont_write_bytecode = True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--streams', type=int, default=64)
    ap.add_argument('--groups', type=int, nargs='+', default=[1, 2, 4])
    ap.add_argument('--steps', type=int, default=20)
    a = ap.parse_args()
    import numpy as np
    import torch
    from bench import rotations_for, workload
    from image_processing import _native
    from synth_euroc import SlidingTextureStream
    cfg, skw, _ = workload('c2')
    S, K, W = a.streams, a.steps, 5
    F = W + K + 1
    n = F + 2 * (S - 1)
    stream = SlidingTextureStream(n_frames=n, **skw)
    frames = [stream.frame(k) for k in range(n)]
    stream.frames = lambda: iter(frames)
    Rs = rotations_for(cfg, stream)
    for G in a.groups:
        Sg = S // G
        ctxs, devs = [], []
        for gidx in range(G):
            ctx = _native.Context(cfg, stream.w, stream.h, num_streams=Sg, device=0, use_graph=True)
            bb, ib = ctx.block_bytes, stream.w * stream.h
            host = torch.zeros((F, bb), dtype=torch.uint8)
            hb = host.numpy()
            for k in range(F):
                for j in range(Sg):
                    s = gidx * Sg + j
                    f = frames[k + 2 * s]
                    hb[k, (2 * j) * ib:(2 * j + 1) * ib] = f.cam0_image.reshape(-1)
                    hb[k, (2 * j + 1) * ib:(2 * j + 2) * ib] = f.cam1_image.reshape(-1)
                ctx.fill_rotations(hb[k], np.stack([Rs[k + 2 * (gidx * Sg + j)][0] for j in range(Sg)]),
                                   np.stack([Rs[k + 2 * (gidx * Sg + j)][1] for j in range(Sg)]))
            ctxs.append(ctx)
            devs.append((host.cuda(), bb))
        for k in range(W + 1):
            for ctx, (dev, bb) in zip(ctxs, devs):
                ctx.enqueue_device(dev.data_ptr() + k * bb)
            for ctx in ctxs:
                ctx.sync()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(W + 1, W + 1 + K):
            for ctx, (dev, bb) in zip(ctxs, devs):
                ctx.enqueue_device(dev.data_ptr() + k * bb)
        for ctx in ctxs:
            ctx.sync()
        dt = time.perf_counter() - t0
        nf = sum(int(ctx.result(j)[0]['n_features']) for ctx in ctxs for j in range(Sg))
        print(f'S={S} in {G} group(s) of {Sg}: {1e3 * dt / K:.4f} ms/step = {S * K / dt:9.0f} frames/s, features {nf / S:.1f}')
        for ctx in ctxs:
            ctx.close()
        del devs


if __name__ == '__main__':
    main()
