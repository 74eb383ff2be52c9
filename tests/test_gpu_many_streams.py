"""GPU parity of the MANY-STREAM launch shapes (BASELINE config C4: time-offset runs of a sequence, `run.bat:4-12`, 64 in
total = 64 per GPU or 8 per GPU on eight).  A context with many streams takes different launch paths from the
single-stream one the other parity tests pin: one warp per feature in the LK kernels, `k_pyr_down` for every pyramid
level instead of `k_pyr_pair`, new-feature stereo matching in two dense rounds, a device mirror of the result block and
one bulk D2H copy instead of zero-copy stores.  Here those shapes are compared, stream by stream and frame by frame,
with single-stream contexts AND with the oracle port (reference: `image_processing/pipeline.py:46-150` per stream)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from frontend_config import FrontEndConfig, config_c2
from oracle import cv_semantics as cs
from oracle.pipeline_port import FrontEndPort
from synth_euroc import SlidingTextureStream

LOSSY = dict(seed=5, sigma=2.0, drift=(2.2, -1.4), gyro=(0.02, 0.01, -0.03), noise=2.5,
             movers=[(200, 150, 40, -3.0, 4.0), (520, 300, 50, 5.0, -2.5), (380, 240, 30, -6.0, -5.0)])


def _rotations(cfg, stream):
    """cam0_R_p_c / cam1_R_p_c between consecutive frames of the sequence, as the pipeline's IMUProcessor forms them."""
    from image_processing import IMUProcessor
    imu = IMUProcessor(cfg.T_imu_cam0, cfg.T_imu_cam1)
    out, prev = [], None
    for kind, msg in stream.events():
        if kind == 'imu':
            imu.imu_callback(msg)
            continue
        imu.cam0_prev_img_msg, imu.cam0_curr_img_msg = prev, msg.cam0_msg
        out.append((np.eye(3), np.eye(3)) if prev is None else tuple(imu.integrate_imu_data()))
        prev = msg.cam0_msg
    return out


def _snapshot(ctx, s):
    hdr, ids, meas = ctx.result(s)
    cell, life, p0, p1 = ctx.features(s)
    return dict(ids=ids.copy(), meas=meas.copy(), cell=cell, life=life, p0=p0, p1=p1,
                hdr=tuple(int(hdr[k]) for k in ('n_features', 'next_feature_id', 'before_tracking', 'after_tracking',
                                                'after_matching', 'after_ransac', 'has_new', 'n_fast', 'n_candidates',
                                                'frame_index')))


def _offset_runs(cfg, S, n_steps, stride, skw, width=752, height=480):
    """S time-offset runs of ONE sequence (run s starts `stride * s` frames in), lock-stepped in one S-stream context
    and, for comparison, each alone in a single-stream context.  Returns (multi[s][k], single[s][k], frames, Rs)."""
    from image_processing import _native
    n = stride * (S - 1) + n_steps
    st = SlidingTextureStream(width=width, height=height, n_frames=n, **skw)
    frames = [st.frame(k) for k in range(n)]
    st.frames = lambda: iter(frames)
    Rs = _rotations(cfg, st)
    multi = [[] for _ in range(S)]
    ctx = _native.Context(cfg, width, height, num_streams=S)
    try:
        for k in range(n_steps):
            idx = [stride * s + k for s in range(S)]
            R0 = None if k == 0 else np.stack([Rs[i][0] for i in idx])
            R1 = None if k == 0 else np.stack([Rs[i][1] for i in idx])
            ctx.process([frames[i].cam0_image for i in idx], [frames[i].cam1_image for i in idx], R0, R1)
            for s in range(S):
                multi[s].append(_snapshot(ctx, s))
        kernels = ctx.kernels_per_frame()
    finally:
        ctx.close()
    single = [[] for _ in range(S)]
    one = _native.Context(cfg, width, height, num_streams=1)
    try:
        for s in range(S):
            one.reset()
            for k in range(n_steps):
                i = stride * s + k
                one.process([frames[i].cam0_image], [frames[i].cam1_image], None if k == 0 else Rs[i][0],
                            None if k == 0 else Rs[i][1])
                single[s].append(_snapshot(one, 0))
    finally:
        one.close()
    return multi, single, frames, Rs, kernels


def _port_run(cfg, st, first, n_steps, backend='cv2'):
    """The oracle port on frames first .. first + n_steps - 1 of the sequence (IMU samples older than the run's first
    frame never reach a window: imu_processor.py:28-48 starts at t_prev - 0.01)."""
    fe = FrontEndPort(cfg, backend=backend)
    out, k = [], 0
    for kind, msg in st.events():
        if kind == 'imu':
            fe.imu_callback(msg)
            continue
        if k >= first:
            fm = fe.stereo_callback(msg)
            pub = np.array([[f.u0, f.v0, f.u1, f.v1] for f in fm.features], np.float64).reshape(-1, 4)
            out.append(dict(ids=fe.ids.copy(), cell=fe.cell.copy(), life=fe.life.copy(), p0=fe.p0.astype(np.float64),
                            p1=fe.p1.astype(np.float64), pub=pub))
            if len(out) == n_steps:
                break
        k += 1
    return out


def _assert_equal_runs(a, b, what):
    for k, (x, y) in enumerate(zip(a, b)):
        assert x['hdr'] == y['hdr'], f'{what} frame {k}: header / counters {x["hdr"]} vs {y["hdr"]}'
        for key in ('ids', 'cell', 'life', 'p0', 'p1', 'meas'):
            assert np.array_equal(x[key], y[key]), f'{what} frame {k}: {key} differ'


def _assert_equals_port(run, ref, what):
    worst = 0.0
    for k, (f, r) in enumerate(zip(run, ref)):
        assert np.array_equal(f['ids'], r['ids']), f'{what} frame {k}: ids differ from the port'
        assert np.array_equal(f['cell'], r['cell']) and np.array_equal(f['life'], r['life']), f'{what} frame {k}'
        if len(f['ids']):
            d = max(np.abs(f['p0'] - r['p0']).max(), np.abs(f['p1'] - r['p1']).max())
            worst = max(worst, float(d))
            assert d <= 0.01, f'{what} frame {k}: position off by {d} px'
            assert np.abs(f['meas'] - r['pub']).max() <= 1e-4, f'{what} frame {k}: published coordinates'
    return worst


@pytest.mark.parametrize('S', [8, 64])
def test_offset_runs_in_one_context_equal_single_contexts_and_the_port(S):
    """The C4 shape at C2 geometry (752x480, 6x10 cells x 5): S offset runs of a LOSSY sequence (moving patches, noise,
    gyro) over 7 frames incl. frame 0.  S = 64 is the one-GPU sweep (per-level pyramid kernel, two candidate rounds,
    bulk result copy), S = 8 the per-GPU shard of the eight-GPU sweep."""
    cfg = config_c2()
    n_steps, stride = 7, 2
    multi, single, frames, Rs, kernels = _offset_runs(cfg, S, n_steps, stride, LOSSY)
    for s in range(S):
        _assert_equal_runs(multi[s], single[s], f'S={S} stream {s}')
    st = SlidingTextureStream(n_frames=len(frames), **LOSSY)
    st.frames = lambda: iter(frames)
    worst = 0.0
    for s in sorted({0, S // 2, S - 1}):
        worst = max(worst, _assert_equals_port(multi[s], _port_run(cfg, st, stride * s, n_steps), f'S={S} stream {s}'))
    lost = sum(f['hdr'][3] - f['hdr'][4] for s in range(S) for f in multi[s][1:])
    new_ids = sum(multi[s][-1]['hdr'][1] for s in range(S))
    print(f'S={S}: {S} x {n_steps} frames equal {S} single-stream contexts bit for bit ({kernels} kernels per frame); streams '
          f'0, {S // 2}, {S - 1} equal the port (worst {worst:.3g} px); {lost} features lost in stereo matching, {new_ids} ids handed out')
    assert lost > 0
    if S == 64:
        assert kernels == 9             # fast, 3 x k_pyr_down, track, select, 2 candidate rounds, finish


def test_forced_two_round_candidates_and_bulk_result_copy_on_a_small_context(monkeypatch):
    """The many-stream choices forced onto ONE stream (AVB_WPF=1, AVB_CAND_ROUNDS=2, AVB_PYR_PAIR=0, AVB_ZC_OUT=0): 30
    lossy frames against the port.  Round 1 of the candidate matching leaves stale inlier flags behind unmatched tail
    positions (avb_points.cu); a stream whose cells keep losing and refilling is what would expose them."""
    for k, v in (('AVB_WPF', '1'), ('AVB_CAND_ROUNDS', '2'), ('AVB_PYR_PAIR', '0'), ('AVB_ZC_OUT', '0')):
        monkeypatch.setenv(k, v)
    from image_processing import _native
    cfg = FrontEndConfig(grid_row=6, grid_col=10, grid_min=2, grid_max=5)
    n = 30
    st = SlidingTextureStream(n_frames=n, **LOSSY)
    frames = [st.frame(k) for k in range(n)]
    st.frames = lambda: iter(frames)
    Rs = _rotations(cfg, st)
    ctx = _native.Context(cfg, 752, 480, num_streams=1)
    got = []
    try:
        assert ctx.kernels_per_frame() == 1 + 3 + 1 + 1 + 2 + 1
        for k in range(n):
            ctx.process([frames[k].cam0_image], [frames[k].cam1_image], None if k == 0 else Rs[k][0], None if k == 0 else Rs[k][1])
            got.append(_snapshot(ctx, 0))
    finally:
        ctx.close()
    worst = _assert_equals_port(got, _port_run(cfg, st, 0, n), 'forced many-stream paths')
    print(f'forced two-round candidates + per-level pyramid + bulk result copy: 30 lossy frames equal the port (worst {worst:.3g} px), '
          f'{got[-1]["hdr"][1]} ids handed out')


@pytest.mark.parametrize('w,h,levels', [(752, 480, 3), (1280, 1024, 4), (1040, 1030, 5)])
def test_per_level_pyramid_kernel_on_every_level(w, h, levels, monkeypatch):
    """AVB_PYR_PAIR=0: `k_pyr_down` builds EVERY level (what a many-stream context does), including the two smallest
    ones that `k_pyr_pair` builds in a small context; AVB_PYR_PAIR=1: the pair kernel.  Both bit-exact against pyrDown."""
    from image_processing import _native
    cfg = FrontEndConfig(pyramid_levels=levels, width=w, height=h)
    g = np.random.default_rng(w * 3 + h)
    img0 = g.integers(0, 256, (h, w)).astype(np.uint8)
    img1 = g.integers(0, 256, (h, w)).astype(np.uint8)
    for pair in ('0', '1'):
        monkeypatch.setenv('AVB_PYR_PAIR', pair)
        c = _native.Context(cfg, w, h, use_graph=False)
        try:
            c.upload(img0, img1)
            c.build_pyramids()
            for slot, img in ((0, img0), (1, img1)):
                ref = cs.build_pyramid(img, levels)
                for lvl in range(levels + 1):
                    assert np.array_equal(c.download_level(slot, lvl), ref[lvl]), (pair, slot, lvl)
        finally:
            c.close()


@pytest.mark.parametrize('spec_k,spec_wpf', [('0', '1'), ('2', '1'), ('16', '1'), ('16', '4')])
def test_speculative_candidate_matching_equals_the_port(spec_k, spec_wpf, monkeypatch):
    """Few streams: the candidates' stereo matches run speculatively beside k_track (k_select mode 2 + k_spec_match), and
    k_select only looks its candidates up.  AVB_SPEC_K=2 keeps the list shorter than grid_max, so that most candidates
    MISS it and are matched by k_select itself (one warp each); 0 switches speculation off; 16 is the default.
    AVB_SPEC_WPF=4 gives every speculative candidate four warps (measured slower at 16 per cell: 56 us against 46 us, the
    launch no longer fits the GPU at once).  All must publish what the port publishes on 30 lossy frames."""
    monkeypatch.setenv('AVB_SPEC_K', spec_k)
    monkeypatch.setenv('AVB_SPEC_WPF', spec_wpf)
    from image_processing import _native
    cfg = config_c2()
    n = 30
    st = SlidingTextureStream(n_frames=n, **LOSSY)
    frames = [st.frame(k) for k in range(n)]
    st.frames = lambda: iter(frames)
    Rs = _rotations(cfg, st)
    ctx = _native.Context(cfg, 752, 480, num_streams=1)
    got = []
    try:
        assert ctx.kernels_per_frame() == (7 if spec_k == '0' else 8)
        for k in range(n):
            ctx.process([frames[k].cam0_image], [frames[k].cam1_image], None if k == 0 else Rs[k][0], None if k == 0 else Rs[k][1])
            got.append(_snapshot(ctx, 0))
    finally:
        ctx.close()
    worst = _assert_equals_port(got, _port_run(cfg, st, 0, n), f'speculation k={spec_k}')
    print(f'AVB_SPEC_K={spec_k} AVB_SPEC_WPF={spec_wpf}: 30 lossy frames equal the port (worst {worst:.3g} px), {got[-1]["hdr"][1]} ids handed out')


_PORT_CACHE = {}


@pytest.mark.parametrize('wpf', ['4', '1'])
def test_black_and_white_sequence_takes_the_wide_sum_path_and_equals_the_port(wpf, monkeypatch):
    """Black/white blobs with sharp edges: the structure-tensor sums of nearly every window exceed LK_NARROW_MAX, so the
    LK kernels of the frame chain (four warps per feature and one warp per feature) reduce their sums in two exact
    halves instead of one 32-bit REDUX (avb_lk.cuh).  6 frames against the port on its numpy backend (oracle/cv_semantics:
    the same exact-sum arithmetic -> identical features, positions bit for bit).  cv2 itself sums these large products
    in float32 in SIMD-lane order and loses ONE of the 298 features of frame 2 that exact sums keep (status agreement
    99.7 %, inside BASELINE's 99.5 % bar); ids are sequential, so the feature sets differ from there on and the cv2
    backend can only be compared up to that frame."""
    monkeypatch.setenv('AVB_WPF', wpf)
    from image_processing import _native
    cfg = config_c2()
    n = 6
    st = SlidingTextureStream(n_frames=n, seed=11, sigma=1.5, drift=(2.0, 1.0), gyro=(0.01, -0.02, 0.02))
    st._tex = np.where(st._tex > 127.5, 255.0, 0.0)
    frames = [st.frame(k) for k in range(n)]
    st.frames = lambda: iter(frames)
    ix, _ = cs.scharr(frames[0].cam0_image)
    sq = np.pad(ix.astype(np.int64) ** 2, ((8, 7), (8, 7))).cumsum(0).cumsum(1)
    win = sq[15:, 15:] - sq[:-15, 15:] - sq[15:, :-15] + sq[:-15, :-15]
    assert np.mean(win > 3.0e8) > 0.9
    Rs = _rotations(cfg, st)
    ctx = _native.Context(cfg, 752, 480, num_streams=1)
    got = []
    try:
        for k in range(n):
            ctx.process([frames[k].cam0_image], [frames[k].cam1_image], None if k == 0 else Rs[k][0], None if k == 0 else Rs[k][1])
            got.append(_snapshot(ctx, 0))
    finally:
        ctx.close()
    if 'bw' not in _PORT_CACHE:                               # the same reference for both lane mappings
        _PORT_CACHE['bw'] = _port_run(cfg, st, 0, n, backend='numpy')
    worst = _assert_equals_port(got, _PORT_CACHE['bw'], f'black/white wpf={wpf}')
    assert worst == 0.0, worst
    cv = _port_run(cfg, st, 0, 2, backend='cv2')
    worst_cv = _assert_equals_port(got[:2], cv, f'black/white wpf={wpf} vs cv2')
    print(f'AVB_WPF={wpf}: {n} black/white frames identical to the exact-sum port, first 2 within {worst_cv:.3g} px of the cv2 '
          f'port; {got[-1]["hdr"][0]} features, {got[-1]["hdr"][1]} ids handed out')
    assert got[-1]['hdr'][0] > 100


def test_ragged_streams_some_dark_some_full_in_one_launch():
    """Eight offset runs of a sequence whose frames 6 and 7 are flat grey and whose frame 8 is half flat: in the same
    launch some streams hold 300 features, some none at all (empty tables, empty FAST buckets, empty result blocks),
    some are refilling an empty grid.  Every stream equals its single-stream context bit for bit and the port."""
    cfg = config_c2()
    S, n_steps, stride = 8, 8, 1
    skw = dict(seed=4, gyro=(0.01, 0.0, -0.02), blackout={6: 1.0, 7: 1.0, 8: 0.53})
    multi, single, frames, Rs, kernels = _offset_runs(cfg, S, n_steps, stride, skw)
    for s in range(S):
        _assert_equal_runs(multi[s], single[s], f'ragged stream {s}')
    st = SlidingTextureStream(n_frames=len(frames), **skw)
    st.frames = lambda: iter(frames)
    worst = 0.0
    for s in (0, 3, 6, 7):                          # stream 6 STARTS on a dark frame, stream 7 too (frame 7)
        worst = max(worst, _assert_equals_port(multi[s], _port_run(cfg, st, stride * s, n_steps), f'ragged stream {s}'))
    counts = [[f['hdr'][0] for f in multi[s]] for s in range(S)]
    print(f'ragged: features per step, stream 0 {counts[0]}, stream 3 {counts[3]}, stream 6 {counts[6]}; worst {worst:.3g} px')
    assert any(0 in c and 300 in c for c in counts)
    assert any(counts[a][k] == 0 and counts[b][k] >= 299 for k in range(n_steps) for a in range(S) for b in range(S))


@pytest.mark.parametrize('wpf', ['4', '1'])
def test_four_huge_cells_take_the_reduction_path_of_the_selection(wpf, monkeypatch):
    """2 x 2 cells of 376 x 240 px, up to 32 features each: every FAST bucket holds ~1300 keys, far beyond the 256 that
    k_select ranks by counting, so the per-cell top-k (speculative list of 32, masked final list) runs its reduction
    rounds; the mask sieve sees 32 live features per cell.  6 frames against the port."""
    monkeypatch.setenv('AVB_WPF', wpf)
    from image_processing import _native
    cfg = FrontEndConfig(grid_row=2, grid_col=2, grid_min=20, grid_max=32)
    n = 6
    st = SlidingTextureStream(n_frames=n, **LOSSY)
    frames = [st.frame(k) for k in range(n)]
    st.frames = lambda: iter(frames)
    Rs = _rotations(cfg, st)
    ctx = _native.Context(cfg, 752, 480, num_streams=1)
    got = []
    try:
        for k in range(n):
            ctx.process([frames[k].cam0_image], [frames[k].cam1_image], None if k == 0 else Rs[k][0], None if k == 0 else Rs[k][1])
            got.append(_snapshot(ctx, 0))
    finally:
        ctx.close()
    worst = _assert_equals_port(got, _port_run(cfg, st, 0, n), f'2x2 cells wpf={wpf}')
    print(f'AVB_WPF={wpf}: 2 x 2 cells, n_fast {got[-1]["hdr"][7]}, features {[f["hdr"][0] for f in got]}, worst {worst:.3g} px')
    assert got[-1]['hdr'][7] > 4 * 256 and got[-1]['hdr'][0] > 100


def test_soak_3000_frames_speculative_split_graph_path_equals_the_plain_chain(monkeypatch):
    """A timing-dependent fault (a race between the FAST branch's speculative table and the main branch, a stale
    double-buffered result block, a bucket count not left at zero) would not show in a 30-frame parity test every time.
    3000 frames (120 lossy frames walked forth and back) through two contexts fed the same host images: the default
    single-stream shape (speculative matches on the side stream, cam0 part of the chain as a graph of its own) and the
    plain chain (AVB_SPEC_K=0: candidates matched behind k_select).  Every frame's ids, cells, lifetimes, positions and
    published coordinates must be identical."""
    from image_processing import _native
    cfg = config_c2()
    n = 120
    st = SlidingTextureStream(n_frames=n, **LOSSY)
    frames = [st.frame(k) for k in range(n)]
    st.frames = lambda: iter(frames)
    Rs = _rotations(cfg, st)
    order = list(range(n)) + list(range(n - 2, 0, -1))
    a = _native.Context(cfg, 752, 480, num_streams=1)
    monkeypatch.setenv('AVB_SPEC_K', '0')
    b = _native.Context(cfg, 752, 480, num_streams=1)
    try:
        assert a.kernels_per_frame() == 8 and b.kernels_per_frame() == 7
        total = 0
        for step in range(3000):
            k = order[step % len(order)]
            R0, R1 = (None, None) if step == 0 else Rs[k]
            for ctx in (a, b):
                ctx.process([frames[k].cam0_image], [frames[k].cam1_image], R0, R1)
            x, y = _snapshot(a, 0), _snapshot(b, 0)
            assert x['hdr'] == y['hdr'], f'step {step}: {x["hdr"]} vs {y["hdr"]}'
            for key in ('ids', 'cell', 'life', 'p0', 'p1', 'meas'):
                assert np.array_equal(x[key], y[key]), f'step {step}: {key} differ'
            total += x['hdr'][0]
    finally:
        a.close()
        b.close()
    print(f'soak: 3000 frames, {total} published features, both launch shapes identical on every frame')
    assert total > 3000 * 250


def test_soak_pipelined_enqueue_with_previous_result_blocks_equals_synchronous_steps():
    """The sweep driver's overlap (enqueue step k+1, THEN take step k's results from the other parity's result block,
    avb_get_result_prev) against plain synchronous steps: 8 streams x 400 steps from device-resident input blocks (bulk D2H
    result path), every result block identical byte for byte."""
    import torch
    from image_processing import _native
    cfg = config_c2()
    S, n_base, steps = 8, 40, 400
    n = n_base + 2 * (S - 1)
    st = SlidingTextureStream(n_frames=n, **LOSSY)
    frames = [st.frame(k) for k in range(n)]
    st.frames = lambda: iter(frames)
    Rs = _rotations(cfg, st)
    a = _native.Context(cfg, 752, 480, num_streams=S)
    b = _native.Context(cfg, 752, 480, num_streams=S)
    try:
        bb, ib = a.block_bytes, 752 * 480
        host = np.zeros((n_base, bb), np.uint8)
        for k in range(n_base):
            for s in range(S):
                f = frames[k + 2 * s]
                host[k, (2 * s) * ib:(2 * s + 1) * ib] = f.cam0_image.reshape(-1)
                host[k, (2 * s + 1) * ib:(2 * s + 2) * ib] = f.cam1_image.reshape(-1)
            a.fill_rotations(host[k], np.stack([Rs[k + 2 * s][0] for s in range(S)]), np.stack([Rs[k + 2 * s][1] for s in range(S)]))
        dev = torch.from_numpy(host).cuda()
        order = list(range(n_base)) + list(range(n_base - 2, 0, -1))
        ptr = lambda step: dev.data_ptr() + order[step % len(order)] * bb
        want = []
        for step in range(steps):                                   # synchronous: process, read
            b.process_device(ptr(step))
            want.append(b.result_block()[0].copy())
        got = []
        a.enqueue_device(ptr(0))
        for step in range(steps):                                   # pipelined: sync k, enqueue k+1, read k
            a.sync()
            if step + 1 < steps:
                a.enqueue_device(ptr(step + 1))
                got.append(a.result_block(prev=True)[0].copy())
            else:
                got.append(a.result_block()[0].copy())
        feats = 0
        for step in range(steps):
            assert np.array_equal(got[step], want[step]), f'step {step}: result blocks differ'
            feats += int(want[step][:, :48].view(_native.HEADER_DTYPE)['n_features'].sum())
    finally:
        a.close()
        b.close()
    print(f'soak: {S} streams x {steps} steps, {feats} published features, pipelined == synchronous')
    assert feats > S * steps * 250
