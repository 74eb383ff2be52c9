"""Builds the native code in-tree (cross-compiles without a GPU):

  lib/libavb.so                         nvcc, sm_100a: CUDA kernels + the C-ABI of include/avb.h
  image_processing/_avbhost.*.so        gcc: CPython extension, the per-frame host driver (csrc/avb_host.c)
  _msckfhost.*.so                       gcc: CPython extension, scalar inner loops of the host MSCKF (csrc/msckf_host.c)

    python uav-airvision_b200/build.py [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.environ.get('AVB_LIB_OUT') or os.path.join(HERE, 'lib', 'libavb.so')      # AVB_LIB_OUT: variant builds (tools/c4_quick.py --lib)
HOST_EXT = os.path.join(HERE, 'image_processing', '_avbhost' + sysconfig.get_config_var('EXT_SUFFIX'))
MSCKF_EXT = os.path.join(HERE, '_msckfhost' + sysconfig.get_config_var('EXT_SUFFIX'))
SOURCES = ['avb_api.cu', 'avb_pyramid.cu', 'avb_fast.cu', 'avb_points.cu', 'avb_grid.cu', 'avb_ransac.cu', 'avb_store.cu']
HEADERS = ['avb_common.cuh', 'avb_lk.cuh', os.path.join('..', '..', 'include', 'avb.h')]
HOST_DEPS = [os.path.join(CSRC, 'avb_host.c'), os.path.join(HERE, '..', 'include', 'avb.h'), os.path.abspath(__file__)]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-fmad=false',            # cv2's float32/float64 arithmetic is not FMA-contracted; neither is ours
              '-Xcompiler', '-fPIC', '-Xcompiler', '-O2', '--shared', '-lcudart']


def _digest(deps) -> str:
    h = hashlib.sha256()
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, 'rb') as f:
            h.update(f.read())
    h.update(os.environ.get('AVB_EXTRA_NVCC', '').encode())
    return h.hexdigest()


def _stale(target, deps) -> bool:
    """A target is up to date when the digest of its sources (contents, not time stamps: a tree copied to the GPU box
    keeps its built libraries but not necessarily its mtimes) equals the one recorded beside it at build time."""
    if not os.path.exists(target):
        return True
    deps = [d for d in deps if not d.endswith('.so')]
    try:
        with open(target + '.srchash') as f:
            return f.read().strip() != _digest(deps)
    except OSError:
        return True


def _stamp(target, deps) -> None:
    with open(target + '.srchash', 'w') as f:
        f.write(_digest([d for d in deps if not d.endswith('.so')]) + '\n')


def needs_build() -> bool:
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return (_stale(LIB, deps) or _stale(HOST_EXT, HOST_DEPS) or
            _stale(MSCKF_EXT, [os.path.join(CSRC, 'msckf_host.c'), os.path.abspath(__file__)]))


def _run(cmd, verbose, what):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f'{what} failed')


def build(force: bool = False, verbose: bool = False) -> str:
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    if force or _stale(LIB, deps):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
        _run([nvcc] + NVCC_FLAGS + os.environ.get('AVB_EXTRA_NVCC', '').split() + (['-Xptxas', '-v'] if verbose else []) +
             [os.path.join(CSRC, f) for f in SOURCES] + ['-o', LIB], verbose, 'nvcc (libavb.so)')
        _stamp(LIB, deps)
    if force or _stale(HOST_EXT, HOST_DEPS):
        inc = sysconfig.get_paths()['include']
        _run([os.environ.get('CC', 'gcc'), '-O2', '-fPIC', '-shared', '-ffp-contract=off', '-Wall', '-I', inc,
              os.path.join(CSRC, 'avb_host.c'), '-o', HOST_EXT, '-L', os.path.dirname(LIB), '-lavb', '-lm',
              '-Wl,-rpath,$ORIGIN/../lib'], verbose, 'gcc (_avbhost)')
        _stamp(HOST_EXT, HOST_DEPS)
    if force or _stale(MSCKF_EXT, [os.path.join(CSRC, 'msckf_host.c'), os.path.abspath(__file__)]):
        inc = sysconfig.get_paths()['include']
        _run([os.environ.get('CC', 'gcc'), '-O3', '-fPIC', '-shared', '-ffp-contract=off', '-Wall', '-I', inc,
              os.path.join(CSRC, 'msckf_host.c'), '-o', MSCKF_EXT, '-lm'], verbose, 'gcc (_msckfhost)')
        _stamp(MSCKF_EXT, [os.path.join(CSRC, 'msckf_host.c'), os.path.abspath(__file__)])
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
