class FeatureMetaData(object):
    """Per-feature bookkeeping record (same attribute set as the reference's
    image_processing/feature_meta_data.py:1-10)."""
    __slots__ = ('id', 'response', 'lifetime', 'cam0_point', 'cam1_point')

    def __init__(self):
        self.id = None
        self.response = None
        self.lifetime = None
        self.cam0_point = None
        self.cam1_point = None
