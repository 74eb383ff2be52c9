"""oracle/ransac.py (two-point RANSAC; parity unpinned by the reference, which stubs the stage out): known-answer and
property checks of the restatement itself.  The CUDA kernel is compared with it in tests/test_gpu_ransac.py."""
import numpy as np

from oracle import ransac as rs
from ransac_cases import K_EUROC, make_case


def test_iteration_count_and_draws():
    assert rs.num_iterations(0.99) == 7                      # ceil(log(0.01) / log(0.51))
    for n in (2, 3, 7, 300, 2000):
        for h in range(16):
            i, j = rs.draw_pair(5, 17, 1, h, n)
            assert 0 <= i < n and 0 <= j < n and i != j
    assert rs.draw_pair(1, 2, 0, 3, 100) == rs.draw_pair(1, 2, 0, 3, 100)
    assert len({rs.draw_pair(1, 2, 0, h, 1000) for h in range(7)}) == 7
    assert rs.draw_pair(1, 2, 0, 3, 1000) != rs.draw_pair(1, 2, 1, 3, 1000)      # keyed by camera
    assert rs.draw_pair(1, 2, 0, 3, 1000) != rs.draw_pair(1, 3, 0, 3, 1000)      # ... and by frame


def test_tree_sum_is_a_sum():
    rng = np.random.default_rng(0)
    for n in (1, 31, 256, 257, 2000):
        v = rng.uniform(0, 2, n)
        assert abs(rs.tree_sum(v) - v.sum()) < 1e-9
    assert rs.tree_sum(np.zeros(0)) == 0.0


def test_translation_with_outliers_rejected():
    u1, u2, truth = make_case(n=300, outliers=40, seed=3)
    info = {}
    m = rs.two_point_ransac(u1, u2, K_EUROC, 3.0, seed=0, frame_index=1, cam=0, info=info)
    assert m.dtype == bool and m.shape == (300,)
    assert m[truth].mean() > 0.97, 'true inliers must survive'
    # the epipolar constraint only sees the displacement component across the epipolar line
    assert m[~truth].mean() < 0.35, 'most gross outliers must go'
    assert info['best_count'] == int(m.sum())


def test_pure_rotation_shortcut():
    """Mean displacement below one pixel: inlier iff displacement < threshold (no model is fitted)."""
    u1, u2, truth = make_case(n=200, outliers=4, seed=4, translation=(0.0, 0.0, 0.0), jitter_px=0.2, outlier_px=(8, 20))
    m = rs.two_point_ransac(u1, u2, K_EUROC, 3.0)
    assert np.array_equal(m, truth)


def test_small_and_empty_inputs():
    assert rs.two_point_ransac(np.zeros((0, 2)), np.zeros((0, 2)), K_EUROC, 3.0).shape == (0,)
    for n in (1, 2):
        u1, u2, _ = make_case(n=n, outliers=0, seed=5)
        assert not rs.two_point_ransac(u1, u2, K_EUROC, 3.0).any()          # fewer than 3 points: all outliers
    u1, u2, _ = make_case(n=3, outliers=0, seed=6)
    assert rs.two_point_ransac(u1, u2, K_EUROC, 3.0).shape == (3,)


def test_far_points_are_pre_rejected():
    u1, u2, truth = make_case(n=100, outliers=0, seed=7)
    u2 = u2.copy()
    u2[:5] += 60.0 * 2.0 / (K_EUROC[0] + K_EUROC[1])         # > 50 px
    m = rs.two_point_ransac(u1, u2, K_EUROC, 3.0)
    assert not m[:5].any() and m[5:].mean() > 0.9


def test_deterministic_and_seed_dependent():
    u1, u2, _ = make_case(n=500, outliers=200, seed=8)
    a = rs.two_point_ransac(u1, u2, K_EUROC, 3.0, seed=1, frame_index=9, cam=1)
    b = rs.two_point_ransac(u1, u2, K_EUROC, 3.0, seed=1, frame_index=9, cam=1)
    assert np.array_equal(a, b)


def test_conjugate_rotation_matches_imu_processor_relation():
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
    R01 = q * np.sign(np.linalg.det(q))
    w = np.array([0.01, -0.02, 0.03])
    th = np.linalg.norm(w)
    kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) / th
    R0 = np.eye(3) + np.sin(th) * kx + (1 - np.cos(th)) * kx @ kx
    R1 = rs.conjugate_rotation(R01, R0)
    w1 = R01 @ w
    k1 = np.array([[0, -w1[2], w1[1]], [w1[2], 0, -w1[0]], [-w1[1], w1[0], 0]]) / th
    assert np.allclose(R1, np.eye(3) + np.sin(th) * k1 + (1 - np.cos(th)) * k1 @ k1, atol=1e-14)
