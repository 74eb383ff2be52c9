"""ORACLE (test infrastructure).  The deterministic synchronous replay driver lives in the package
(uav-airvision_b200/replay.py, plain host logic); re-exported here for the oracle's callers."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'uav-airvision_b200')
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from replay import run_stream  # noqa: E402,F401
