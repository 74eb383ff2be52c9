"""FeatureMeasurement: stereo measurement of one feature in normalized coordinates; the MSCKF reads exactly
id, u0, v0, u1, v1 (reference image_processing/feature_measurment.py:1-9, msckf.py:430-438).

Same name, same attributes, assignable like the reference's plain class -- but implemented as a C struct with member
descriptors (csrc/avb_host.c) so a frame's list of measurements is built in a few microseconds.  Unset attributes read
0 / 0.0 instead of None."""
from ._avbhost import FeatureMeasurement

__all__ = ['FeatureMeasurement']
