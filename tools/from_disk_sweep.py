"""From-disk sweep (BASELINE config C4/C5 shape, the way the reference's run.bat consumes data): rendered sequences written
as EuRoC directories (PNG) on tmpfs, then `run_sweep.py` over them.  Reports where the time goes: PNG decode + upload into
the HBM frame store (once per sequence) against the sweep itself, and the resulting frames/s over the whole job.

    python tools/from_disk_sweep.py --sequences 2 --frames 160 --offsets 0 0.5 1 1.5 2 2.5 3 3.5 --out gpurun_out/from_disk.json
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
    sys.path.insert(0, p)
sys.dont_write_bytecode = True


def _write(job):
    q, n, root = job
    for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200')):
        if p not in sys.path:
            sys.path.insert(0, p)
    from euroc import write_euroc
    from frontend_config import FrontEndConfig
    from synth_euroc import RoomSceneStream
    st = RoomSceneStream(FrontEndConfig(), n_frames=n, seed=300 + q, tex_size=1024, amp=0.8 + 0.1 * q)
    path = os.path.join(root, f'SEQ_{q:02d}')
    write_euroc(path, st)
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sequences', type=int, default=2)
    ap.add_argument('--frames', type=int, default=160)
    ap.add_argument('--offsets', nargs='+', type=float, default=[0, 0.5, 1, 1.5, 2, 2.5, 3, 3.5])
    ap.add_argument('--workers', type=int, default=0)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    import multiprocessing as mp
    import numpy as np
    root = tempfile.mkdtemp(prefix='avb_euroc_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None)
    try:
        t0 = time.perf_counter()
        with mp.get_context('spawn').Pool(min(a.sequences, os.cpu_count() or 1)) as pool:
            paths = pool.map(_write, [(q, a.frames, root) for q in range(a.sequences)])
        render_s = time.perf_counter() - t0
        png_bytes = sum(os.path.getsize(os.path.join(dp, f)) for p_ in paths for dp, _, fs in os.walk(p_) for f in fs if f.endswith('.png'))
        # decode alone: the reader's threaded prefetch, no GPU
        from euroc import EuRoCDataset
        t0 = time.perf_counter()
        n_dec = 0
        for p_ in paths:
            ds = EuRoCDataset(p_)
            ds.set_starttime(0)
            for _ in ds.stereo.prefetch():
                n_dec += 1
        decode_s = time.perf_counter() - t0
        # the sweep job itself (decode + upload + lock-stepped runs + estimators + output files)
        from run_sweep import main as sweep_main
        out_dir = os.path.join(root, 'results')
        argv = ['--path', *paths, '--offsets', *[str(o) for o in a.offsets], '--out', out_dir]
        if a.workers:
            argv += ['--workers', str(a.workers)]
        import contextlib
        import io
        buf = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(buf):
            rc = sweep_main(argv)
        job_s = time.perf_counter() - t0
        line = buf.getvalue().strip().splitlines()[-1] if buf.getvalue().strip() else ''
        import csv
        rows = list(csv.DictReader(open(os.path.join(out_dir, 'metrics_summary.csv'))))
        frames_run = sum(int(r['frames']) for r in rows)
        res = {'sequences': a.sequences, 'frames_per_sequence': a.frames, 'offsets_s': a.offsets, 'runs': len(rows),
               'png_bytes': png_bytes, 'render_and_write_s': round(render_s, 2),
               'decode_only': {'stereo_frames': n_dec, 'seconds': round(decode_s, 3), 'stereo_frames_per_s': n_dec / decode_s,
                               'note': 'EuRoCDataset.stereo.prefetch (8 decode threads, zlib inflate + C unfilter), files on tmpfs'},
               'job': {'rc': rc, 'seconds': round(job_s, 3), 'frames_processed': frames_run, 'frames_per_s': frames_run / job_s,
                       'summary_line': line,
                       'note': 'run_sweep.py end to end: decode + upload once per sequence, all offset runs in lock-step, host MSCKF '
                               'workers, trajectory files + metrics_summary.csv'},
               'host_cores': os.cpu_count()}
        print(json.dumps(res, indent=1))
        if a.out:
            with open(a.out, 'w') as f:
                json.dump(res, f, indent=1)
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == '__main__':
    main()
