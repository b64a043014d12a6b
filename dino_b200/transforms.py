"""Inference-time preprocessing of DINOSeg.predict (reference pl_torch_modules.py:33-41,291).

The reference composes albumentations 1.1.0 `Resize(r, r)` -> `Normalize(ImageNet mean/std)` ->
`ToTensorV2()`.  albumentations is not available offline, so this is a restatement of its
documented behaviour (parity unpinned): Resize = cv2.resize(INTER_LINEAR);
Normalize = (img - mean*255) * (1 / (std*255)) in fp32; ToTensorV2 = HWC -> CHW torch tensor.
It is host-side and outside the timed hot path.
"""
from __future__ import annotations

import numpy as np
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class Compose:
    """Callable with albumentations' calling convention: t(image=ndarray) -> {'image': Tensor}."""

    def __init__(self, resolution: int):
        self.resolution = int(resolution)
        self.mean = np.array(IMAGENET_MEAN, dtype=np.float32) * 255.0
        self.denom = np.reciprocal(np.array(IMAGENET_STD, dtype=np.float32) * 255.0)

    def __call__(self, *, image, **kwargs):
        import cv2
        img = np.asarray(image)
        if img.shape[0] != self.resolution or img.shape[1] != self.resolution:
            img = cv2.resize(img, dsize=(self.resolution, self.resolution), interpolation=cv2.INTER_LINEAR)
        img = img.astype(np.float32)
        img -= self.mean
        img *= self.denom
        out = {"image": torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))}
        out.update(kwargs)
        return out


def get_transforms(resolution: int = 480) -> Compose:
    """Same name and role as the reference's get_transforms (pl_torch_modules.py:33-41)."""
    return Compose(resolution)
