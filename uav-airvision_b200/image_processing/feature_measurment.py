class FeatureMeasurement(object):
    """Stereo measurement of one feature in normalized coordinates; the MSCKF reads exactly
    id, u0, v0, u1, v1 (reference image_processing/feature_measurment.py:1-9, msckf.py:430-438)."""
    __slots__ = ('id', 'u0', 'v0', 'u1', 'v1')

    def __init__(self, id=None, u0=None, v0=None, u1=None, v1=None):
        self.id = id
        self.u0 = u0
        self.v0 = v0
        self.u1 = u1
        self.v1 = v1
