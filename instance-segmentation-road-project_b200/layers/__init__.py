"""Drop-in mirror of /root/reference/engine/layers/ for the post-backbone hot path."""
from .base import Layer, get_custom_objects                                    # noqa: F401
from .detection import PriorLayer, RestoreBoxes, NormalizeBoxes, DetectionProposal   # noqa: F401
from .instance import MaskDistribute, PyramidRoiAlign, TrimInstances          # noqa: F401
from .misc import DownSampleInput, MoldBatch, ResizeLike, UpSampleOutput, CropAndPadMask, EncodeImageContent                   # noqa: F401
from .summary import CrackToInstance, SummaryOutput, IncludeMyRoad, CalculateInstanceSize   # noqa: F401
from .draw import DrawBoxes, DrawSegmentation, DrawInstance                              # noqa: F401
from .semantic import SemanticSmoothing                                       # noqa: F401
from .training import CalculateIOU, AssignBoxes, AssignMasks, DetectionIOUMetric  # noqa: F401
