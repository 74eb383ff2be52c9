"""ORACLE (test infrastructure).  The BASELINE configurations live in the package
(uav-airvision_b200/frontend_config.py, plain data); re-exported here for the oracle's callers."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'uav-airvision_b200')
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

from frontend_config import (FrontEndConfig, config_c1, config_c2, config_c3, config_default,  # noqa: E402,F401
                             OPTFLOW_USE_INITIAL_FLOW, TERM_COUNT, TERM_EPS)
