"""Pins oracle/pipeline_port.py against golden dumps of the UNMODIFIED reference front end
(tests/golden/*.npz, written by tools/make_golden.py from /root/reference/src)."""
import os

import numpy as np
import pytest

pytest.importorskip('cv2')

from frontend_config import FrontEndConfig
from replay import run_stream
from oracle.pipeline_port import FrontEndPort
from synth_euroc import SlidingTextureStream
from tools.make_golden import CASES


def _run_port(name, backend='cv2', max_frames=None):
    gr, gc, gmin, gmax, skw = CASES[name]
    cfg = FrontEndConfig(grid_row=gr, grid_col=gc, grid_min=gmin, grid_max=gmax)
    if max_frames:
        skw = dict(skw, n_frames=max_frames)
    fe = FrontEndPort(cfg, backend=backend)
    grids = []

    def on_frame(k, msg, fm):
        grids.append(dict(ids=fe.ids.copy(), life=fe.life.copy(), cell=fe.cell.copy(),
                          p0=fe.p0.copy(), p1=fe.p1.copy(), fresh=fe.fresh.copy(),
                          counters=dict(fe.num_features)))

    msgs = run_stream(fe, SlidingTextureStream(**skw), on_frame=on_frame)
    return fe, msgs, grids


@pytest.mark.parametrize('name', list(CASES))
def test_port_reproduces_reference_exactly(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    fe, msgs, grids = _run_port(name)
    assert len(msgs) == int(g['n_frames'][0])
    for k, (fm, gr) in enumerate(zip(msgs, grids)):
        assert np.array_equal(gr['ids'], g[f'f{k}_ids']), f'frame {k} ids'
        assert np.array_equal(gr['life'], g[f'f{k}_life'])
        assert np.array_equal(gr['cell'], g[f'f{k}_cell'])
        assert np.array_equal(gr['p0'].astype(np.float64), g[f'f{k}_p0'])      # bit-exact
        assert np.array_equal(gr['p1'].astype(np.float64), g[f'f{k}_p1'])
        pub = np.array([[f.u0, f.v0, f.u1, f.v1] for f in fm.features], np.float64).reshape(-1, 4)
        assert np.array_equal([f.id for f in fm.features], g[f'f{k}_pub_ids'])
        assert np.array_equal(pub, g[f'f{k}_pub'])
        if len(fm.features):
            assert (np.asarray(fm.features[0].u0).dtype == np.float64) == bool(g[f'f{k}_u0_is_f64'][0])
        assert fm.timestamp == g[f'f{k}_ts'][0]
        if k > 0:
            c = gr['counters']
            assert [c.get('before_tracking', -1), c.get('after_tracking', -1),
                    c.get('after_matching', -1), c.get('after_ransac', -1)] == list(g[f'f{k}_counters'])
    assert fe.next_feature_id == int(g['next_feature_id'][0])


def test_port_numpy_backend_closes_the_loop(golden_dir):
    """The numpy restatement of the cv2 arithmetic reproduces the reference end to end."""
    name = 'ref_sparse_s3'
    g = np.load(os.path.join(golden_dir, name + '.npz'))
    fe, msgs, grids = _run_port(name, backend='numpy')
    for k, gr in enumerate(grids):
        assert np.array_equal(gr['ids'], g[f'f{k}_ids'])
        assert np.abs(gr['p0'].astype(np.float64) - g[f'f{k}_p0']).max() < 1e-3
        assert np.abs(gr['p1'].astype(np.float64) - g[f'f{k}_p1']).max() < 1e-3
