"""dino_b200 — B200-native (sm_100a) implementation of the DINOSeg inference hot path of
sachaMorin/dino, behind the reference's own Python surface.

    from dino_b200 import DINOSeg          # or: from dt_segmentation import DINOSeg
    m = DINOSeg.load_from_checkpoint(path).to('cuda:0')
    m.set_resolution(480)
    labels = m.predict(pil_image)          # int64 [480, 480]
"""
from .model import DINOSeg  # noqa: F401
from .transforms import get_transforms  # noqa: F401

__all__ = ["DINOSeg", "get_transforms"]
