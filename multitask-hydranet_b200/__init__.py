"""hydranet_b200: B200-native implementation of HydraNet's multitask forward path + post-processing.

Importable as ``hydranet_b200`` (see ``hydranet_b200.py`` at the repo root; the package directory
itself is named ``multitask-hydranet_b200`` and is loaded through importlib).
"""
from . import _native  # noqa: F401  (fails loudly if libhydranet_b200.so is missing)
from .heads import DetectionHeader, LaneHeader, SegmentHeader, make_anchors  # noqa: F401
from .lane_codec import Lane, LaneCodec, Point, convert_lane_to_dict, order_lane_x_axis  # noqa: F401
from .model import HydraNet  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .parallel import GradAllReduce  # noqa: F401
from .preprocess import preprocess  # noqa: F401
from .train import TrainStep  # noqa: F401

__all__ = ["HydraNet", "SegmentHeader", "DetectionHeader", "LaneHeader", "LaneCodec", "Lane", "Point",
           "make_anchors", "order_lane_x_axis", "convert_lane_to_dict", "preprocess", "FusedAdam", "GradAllReduce", "TrainStep"]
