"""Host MSCKF timing on the committed 400-frame feature dump (tests/golden): best of N replays, mean ms per frame over the
frames after gravity initialisation.  Host-only (no GPU): `python tools/bench_msckf.py [N]`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'uav-airvision_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import numpy as np                      # noqa: E402
import test_msckf_host as T             # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
best, med = 1e9, 1e9
for _ in range(n):
    got, want, secs, ref_ms, est = T._replay(os.path.join(ROOT, 'tests', 'golden'))
    best, med = min(best, float(np.sum(secs[20:]))), min(med, float(np.median(secs[20:])))
print(f'host MSCKF: best of {n} replays: {1e3 * best / 380:.2f} ms/frame mean, {1e3 * med:.2f} ms/frame median '
      f'(reference filter when the fixture was made: {ref_ms:.2f} ms/frame median)')
