"""Latency probe for the LK kernels: serialised stage times of the C2 single-stream frame under varied LK parameters
(max_iteration, pyramid levels, AVB_WPF).  Not a benchmark; explains where k_track's time goes."""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
    sys.path.insert(0, p)
sys.dont_write_bytecode = True


def run(max_iter, levels, wpf, n=14):
    import torch
    from bench import rotations_for, workload
    from synth_euroc import SlidingTextureStream
    from image_processing import _native
    cfg, skw, _ = workload('c2')
    cfg.max_iteration = max_iter
    cfg.pyramid_levels = levels
    cfg.lk_params = dict(cfg.lk_params, maxLevel=levels, criteria=(3, max_iter, 0.01))
    os.environ['AVB_WPF'] = str(wpf)
    stream = SlidingTextureStream(n_frames=n, **skw)
    frames = [stream.frame(k) for k in range(n)]
    stream.frames = lambda: iter(frames)
    Rs = rotations_for(cfg, stream)
    ctx = _native.Context(cfg, stream.w, stream.h, num_streams=1, device=0, use_graph=True)
    bb, ib = ctx.block_bytes, stream.w * stream.h
    host = torch.zeros((n, bb), dtype=torch.uint8)
    hb = host.numpy()
    for k, f in enumerate(frames):
        hb[k, :ib] = f.cam0_image.reshape(-1)
        hb[k, ib:2 * ib] = f.cam1_image.reshape(-1)
        ctx.fill_rotations(hb[k], Rs[k][0], Rs[k][1])
    dev = host.cuda()
    for k in range(4):
        ctx.process_device(dev.data_ptr() + k * bb)
    acc = {}
    for k in range(4, n):
        st = ctx.profile_frame_device(dev.data_ptr() + k * bb)
        for a, b in st.items():
            acc.setdefault(a, []).append(b)
    nfeat = int(ctx.result(0)[0]['n_features'])      # read before close: the views point into pinned memory
    ctx.close()
    med = {a: float(np.median(b)) * 1e3 for a, b in acc.items()}
    print(f'max_iter={max_iter:2d} levels={levels + 1} wpf={wpf}: track {med["track"]:6.1f} us  stereo_new {med["stereo_new"]:6.1f} us  '
          f'select {med["select"]:5.1f}  pyramid {med["pyramid"]:5.1f}  fast {med["fast"]:5.1f}  features {nfeat}')


if __name__ == '__main__':
    for wpf in (4, 1):
        for mi, lv in ((30, 3), (1, 3), (0, 3), (30, 0), (1, 0), (0, 0)):
            run(mi, lv, wpf)
