"""Builds libavb.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python uav-airvision_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'lib', 'libavb.so')
SOURCES = ['avb_api.cu', 'avb_pyramid.cu', 'avb_fast.cu', 'avb_points.cu', 'avb_grid.cu']
HEADERS = ['avb_common.cuh', 'avb_lk.cuh', os.path.join('..', '..', 'include', 'avb.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-fmad=false',            # cv2's float32/float64 arithmetic is not FMA-contracted; neither is ours
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '--shared', '-lcudart']


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
          [os.path.join(CSRC, f) for f in SOURCES] + ['-o', LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed building libavb.so')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
