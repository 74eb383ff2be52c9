"""Generate tests/golden/*.npz by running the UNMODIFIED reference front end
(/root/reference/src/image_processing, imported in place) on seeded synthetic streams with
the deterministic driver.  Runs only in the build container (the reference cannot travel to
the GPU box); the fixtures it writes are committed.

    python tools/make_golden.py            # rewrites every fixture listed in CASES
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.dont_write_bytecode = True


def _reference_first():
    """Only when run as a script (tests import CASES from this module and must keep their own sys.path): the
    reference's src/ FIRST, so that `image_processing` and `config` resolve to the UNMODIFIED reference; the package
    directory stays on the path for synth_euroc / replay / frontend_config (it also holds the drop-in `image_processing`)."""
    for p in (os.path.join(ROOT, 'uav-airvision_b200'), ROOT, '/root/reference/src'):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


CASES = {
    # name: (grid_row, grid_col, grid_min, grid_max, stream kwargs)
    'ref_default_s0': (4, 5, 3, 5, dict(n_frames=10, seed=0, sigma=2.5)),
    'ref_c2_s1': (6, 10, 3, 5, dict(n_frames=8, seed=1, sigma=2.0, drift=(1.7, -0.8))),
    'ref_c1_gyro_s2': (5, 6, 3, 5, dict(n_frames=8, seed=2, sigma=3.0, drift=(-1.1, 0.9),
                                         gyro=(0.02, -0.03, 0.05), noise=1.5)),
    'ref_sparse_s3': (4, 5, 3, 5, dict(n_frames=6, seed=3, sigma=5.0, drift=(0.4, 0.2))),
    # frames 4 and 5 are flat grey (every feature lost, empty feature messages), frame 6 is flat in its left half: the
    # stream refills an EMPTY grid through the adder (not the first-frame initialiser)
    'ref_c2_dark_s4': (6, 10, 3, 5, dict(n_frames=10, seed=4, gyro=(0.01, 0.0, -0.02), blackout={4: 1.0, 5: 1.0, 6: 0.53})),
}


def dump_case(name, spec):
    from config import ConfigEuRoC                      # reference config, unmodified
    from image_processing import ImageProcessor        # reference front end, unmodified
    from synth_euroc import SlidingTextureStream
    from replay import run_stream

    gr, gc, gmin, gmax, skw = spec
    cfg = ConfigEuRoC()
    cfg.grid_row, cfg.grid_col, cfg.grid_num = gr, gc, gr * gc
    cfg.grid_min_feature_num, cfg.grid_max_feature_num = gmin, gmax
    ip = ImageProcessor(cfg)
    stream = SlidingTextureStream(**skw)
    rec = {}

    def on_frame(k, msg, fm):
        ids, life, cell, p0, p1 = [], [], [], [], []
        for c, feats in enumerate(ip.prev_features):       # rolled: this frame's grid
            for f in feats:
                ids.append(f.id); life.append(f.lifetime); cell.append(c)
                p0.append(np.asarray(f.cam0_point, dtype=np.float64))
                p1.append(np.asarray(f.cam1_point, dtype=np.float64))
        n = len(ids)
        rec[f'f{k}_ids'] = np.asarray(ids, np.int64)
        rec[f'f{k}_life'] = np.asarray(life, np.int64)
        rec[f'f{k}_cell'] = np.asarray(cell, np.int64)
        rec[f'f{k}_p0'] = np.asarray(p0, np.float64).reshape(n, 2)
        rec[f'f{k}_p1'] = np.asarray(p1, np.float64).reshape(n, 2)
        pub = np.array([[f.u0, f.v0, f.u1, f.v1] for f in fm.features], np.float64).reshape(-1, 4)
        rec[f'f{k}_pub'] = pub
        rec[f'f{k}_pub_ids'] = np.asarray([f.id for f in fm.features], np.int64)
        rec[f'f{k}_u0_is_f64'] = np.asarray(
            [len(fm.features) > 0 and np.asarray(fm.features[0].u0).dtype == np.float64])
        rec[f'f{k}_ts'] = np.asarray([fm.timestamp])
        nf = ip.num_features
        rec[f'f{k}_counters'] = np.asarray([nf.get('before_tracking', -1), nf.get('after_tracking', -1),
                                            nf.get('after_matching', -1), nf.get('after_ransac', -1)])

    run_stream(ip, stream, on_frame=on_frame)
    rec['n_frames'] = np.asarray([stream.n])
    rec['next_feature_id'] = np.asarray([ip.next_feature_id])
    rec['spec'] = np.asarray([gr, gc, gmin, gmax])
    import image_processing
    assert image_processing.__file__.startswith('/root/reference/'), image_processing.__file__
    path = os.path.join(os.environ.get('GOLDEN_OUT', os.path.join(ROOT, 'tests', 'golden')), name + '.npz')
    np.savez_compressed(path, **rec)
    print(name, 'frames', stream.n, 'features/frame',
          [len(rec[f'f{k}_ids']) for k in range(stream.n)], '->', os.path.getsize(path), 'B')


if __name__ == '__main__':
    _reference_first()
    import cv2
    print('reference run with cv2', cv2.__version__, 'numpy', np.__version__)
    only = sys.argv[1:]                                 # optional: names of the cases to (re)write
    for name, spec in CASES.items():
        if not only or name in only:
            dump_case(name, spec)
