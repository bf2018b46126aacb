# Executed as the body of the `masklab_b200` package (see masklab_b200/__init__.py).
# B200-native drop-in for the post-backbone hot path of MaskLab / RetinaMask
# (craftsangjae/instance-segmentation-road-project): host-side mirror of the reference's
# Keras layer interface over the C ABI in include/masklab_b200.h.
from .prior import PriorBoxes                                    # noqa: F401
from .layers import (PriorLayer, RestoreBoxes, NormalizeBoxes, DetectionProposal,   # noqa: F401
                     DownSampleInput, MoldBatch, ResizeLike, MaskDistribute, PyramidRoiAlign, TrimInstances,
                     UpSampleOutput, CropAndPadMask, CrackToInstance, SummaryOutput, IncludeMyRoad,
                     CalculateInstanceSize, DrawBoxes, DrawSegmentation, DrawInstance, SemanticSmoothing,
                     CalculateIOU, AssignBoxes, AssignMasks, DetectionIOUMetric, EncodeImageContent, get_custom_objects)
from .pipeline import PostProcessPipeline, DetectionConfig, expand_clipped      # noqa: F401
from .serving import PostProcessConfig, serving_outputs        # noqa: F401
from .runtime import Context, MaskLabError, InvalidArgumentError, load_library   # noqa: F401

__version__ = "0.1.0"
