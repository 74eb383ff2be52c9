import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'uav-airvision_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
sys.dont_write_bytecode = True


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    # a fresh checkout has no built artefacts (they are git-ignored): build them in-tree once (nvcc cross-compiles for
    # sm_100a without a GPU); a box without nvcc uses the libraries that travelled with the tree
    import importlib.util
    spec = importlib.util.spec_from_file_location('avb_build', os.path.join(PKG, 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if mod.needs_build() and os.path.exists(os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')):
        mod.build()


@pytest.fixture(scope='session')
def golden_dir():
    return os.path.join(ROOT, 'tests', 'golden')
