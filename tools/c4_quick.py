"""Quick A/B probe (not a benchmark): S lock-stepped time-offset runs of the bench sequence through one context, inputs
resident in HBM: ms per step (CUDA events over K enqueued steps) and the serialised stage times.  Also the single-stream
chain when --streams 1.

    python tools/c4_quick.py --streams 64 --steps 20 [--lib path/to/variant/libavb.so]
"""
from __future__ import annotations

import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
    sys.path.insert(0, p)
sys.dont_write_bytecode = True


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--streams', type=int, default=64)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--workload', default='c2')
    ap.add_argument('--lib', default=None, help='variant library copied over lib/libavb.so before loading')
    a = ap.parse_args()
    if a.lib:
        shutil.copyfile(a.lib, os.path.join(ROOT, 'uav-airvision_b200', 'lib', 'libavb.so'))
    import numpy as np
    import torch
    from bench import rotations_for, workload
    from image_processing import _native
    from synth_euroc import SlidingTextureStream
    cfg, skw, _ = workload(a.workload)
    S, K, W = a.streams, a.steps, 5
    F = W + K + 3
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
        ctx.fill_rotations(hb[k], np.stack([Rs[k + 2 * s][0] for s in range(S)]), np.stack([Rs[k + 2 * s][1] for s in range(S)]))
    dev = host.cuda()
    ptr = dev.data_ptr()
    ext = torch.cuda.ExternalStream(ctx.cuda_stream(), device=0)
    for k in range(W + 1):
        ctx.process_device(ptr + k * bb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for k in range(W + 1, W + 1 + K):
        ctx.enqueue_device(ptr + k * bb)
    e1.record(ext)
    ctx.sync()
    ms = e0.elapsed_time(e1) / K
    nf = sum(int(ctx.result(s)[0]['n_features']) for s in range(S))
    st = None
    for k in range(W + 1 + K, F):
        st = ctx.profile_frame_device(ptr + k * bb)
    ctx.close()
    print(f'{os.path.basename(a.lib) if a.lib else "default":28s} S={S}: {ms:.4f} ms/step = {S / ms * 1e3:9.0f} frames/s, features {nf / S:.1f}; '
          + ' '.join(f'{k_}={1e3 * v:.1f}' for k_, v in st.items() if k_ not in ('reserved',)))


if __name__ == '__main__':
    main()
