"""B200-native drop-in for the reference package `image_processing`
(/root/reference/src/image_processing/__init__.py:1-27): same export list, same ImageProcessor facade with
the `stareo_callback` alias.  Put this directory's parent on sys.path ahead of the reference's src/ and
`from image_processing import ImageProcessor` (vio.py:3) resolves here."""
from .pipeline import ImageProcessingPipeline, create_context, current_context
from .camera_model import CameraModel
from .imu_processor import IMUProcessor
from .pyramid_builder import PyramidBuilder
from .feature_meta_data import FeatureMetaData
from .feature_measurment import FeatureMeasurement
from .feature_initializer import FeatureInitializer
from .feature_adder import FeatureAdder
from .feature_tracker import FeatureTracker
from .feature_pruner import FeaturePruner
from .stereo_matcher import StereoMatcher
from .feature_publisher import FeaturePublisher
from .fast_detector import FastDetector


class ImageProcessor(ImageProcessingPipeline):
    """Facade kept for the reference's legacy API (image_processing/__init__.py:14-27)."""

    def __init__(self, config, **kw):
        super().__init__(config, **kw)

    stareo_callback = ImageProcessingPipeline.stereo_callback
