"""Debug probe (needs a library built with AVB_EXTRA_NVCC=-DAVB_DEBUG_CLOCKS): phase durations of k_finish in SM
cycles, reported through the counter fields of the frame header."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
    sys.path.insert(0, p)
from bench import make_sequence, workload
from image_processing import ImageProcessor
from replay import run_stream
cfg, skw, _ = workload(sys.argv[1] if len(sys.argv) > 1 else 'c2')
ip = ImageProcessor(cfg)
rows = []
def on_frame(k, msg, fm):
    h = ip.context.result(0)[0]
    a, b = int(h['n_fast']), int(h['n_candidates'])
    rows.append([int(h[k_]) for k_ in ('before_tracking', 'after_tracking', 'after_matching', 'after_ransac')] +
                ['cell0: head', a & 0xffff, 'place+route', a >> 16, 'gather', b >> 16, 'stores', b & 0xffff])
run_stream(ip, make_sequence(skw, 12), on_frame=on_frame)
print('k_finish cycles [load+hist, scans, per-cell, tail-sync] per frame:')
for r in rows: print(r)
