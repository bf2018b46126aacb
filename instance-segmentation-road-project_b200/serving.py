"""The post-model part of the reference's serving graph, composed from the drop-in layers in the
reference's order: /root/reference/engine/retinamasklab.py:613-636
(load_masklab_inference_model_from_h5: TrimInstances, SemanticSmoothing per class, ResizeLike,
UpSampleOutput) and /root/reference/road_project/setup/serving.py:29-48 (CropAndPadMask, DrawBoxes,
DrawInstance, DrawSegmentation, SummaryOutput).  The dense model in the middle (backbone, heads, mask
head, semantic decoder) and the JPEG decode at the front are not part of this library; the JPEG encode at the
end (serving.py:41, EncodeImageContent) is the optional last step.
"""
from dataclasses import dataclass, field

from .layers import (TrimInstances, SemanticSmoothing, ResizeLike, UpSampleOutput, CropAndPadMask, DrawBoxes,
                     DrawInstance, DrawSegmentation, SummaryOutput, EncodeImageContent)


@dataclass
class PostProcessConfig:
    """ModelConfiguration.postprocess of the reference (engine/config.py:11-45)."""
    resolution: tuple = (540, 960)
    smoothing_kernel_sizes: tuple = (0, 0, 0)
    smoothing_weights: tuple = (1., 1., 1.)
    instance_colors: list = field(default_factory=lambda: [[192, 32, 128], [160, 96, 0], [96, 0, 128],
                                                           [32, 96, 192], [96, 32, 128]])
    instance_alpha: float = 0.3
    semantic_colors: list = field(default_factory=lambda: [[64, 0, 128], [128, 96, 0], [128, 192, 0]])
    semantic_alpha: float = 0.3
    default_road_size: float = 3.25


def serving_outputs(frames, downsampled, box_pred, mask_pred, seg_pred, config=None, mask_output="float32",
                    encode=False):
    """frames uint8 [B,PH,PW,3]; downsampled: the model input ([B,h,w,3] tensor or (h, w)); box_pred
    [B,R,6] and mask_pred [B,R,mh,mw,C]: the model's RoI boxes and mask-head output; seg_pred
    [B,hs,ws,S]: semantic probabilities.  Returns (visualize uint8 [B,PH,PW,3], summarize float32
    [B,M',11], det_outs, ins_outs, seg_outs) like the serving model before the JPEG encode.
    mask_output: dtype of the intermediate pasted masks ('float32' as in the reference, or 'uint8').
    encode=True appends the serving model's first output, EncodeImageContent()([visualize]) (serving.py:41): the
    JPEG file of frame 0 as a one-element list of bytes."""
    cfg = config or PostProcessConfig()
    detection_pred, instance_pred = TrimInstances(mold=True)([box_pred, mask_pred])             # :613-614
    S = int(seg_pred.shape[-1])
    per = S // len(cfg.smoothing_kernel_sizes)
    posts = []
    for i, (k, w) in enumerate(zip(cfg.smoothing_kernel_sizes, cfg.smoothing_weights)):          # :617-626
        posts.append(SemanticSmoothing(kernel_size=k, weight=w)(seg_pred[..., i * per:(i + 1) * per].contiguous()))
    import torch
    post_semantics = torch.cat(posts, dim=-1)
    semantic_pred = ResizeLike()(post_semantics, target=downsampled)                             # :628
    det_outs, ins_outs, seg_outs = UpSampleOutput()([detection_pred, instance_pred, semantic_pred],
                                                    target=frames)                               # :635-636
    masks = CropAndPadMask(output=mask_output)([frames, det_outs, ins_outs])                     # serving.py:30
    vis = DrawBoxes()([frames, det_outs])                                                        # :34
    vis = DrawInstance(cfg.instance_colors, cfg.instance_alpha)([vis, det_outs, masks])          # :35-37
    vis = DrawSegmentation(cfg.semantic_colors, cfg.semantic_alpha)([vis, seg_outs])             # :38-40
    summary = SummaryOutput(default_road_size=cfg.default_road_size)([det_outs, seg_outs, masks])   # :47-48
    if encode:
        return vis, summary, det_outs, ins_outs, seg_outs, EncodeImageContent()([vis])                 # :41
    return vis, summary, det_outs, ins_outs, seg_outs
