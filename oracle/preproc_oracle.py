"""CPU ORACLE for the inference preprocessing of DINOSeg.predict — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Reference: dt_segmentation/src/pl_torch_modules.py:33-41 (`get_transforms`: albumentations Resize(r, r) ->
Normalize(ImageNet mean/std) -> ToTensorV2) applied at :291.  albumentations 1.1.0 (requirements.txt:1) is not in
/root/reference and not installed here; its two arithmetic steps are restated from their documented behaviour:
  * Resize      = cv2.resize(img, (r, r), interpolation=cv2.INTER_LINEAR)        (cv2 IS available: used as the pin)
  * Normalize   = (img.astype(float32) - mean*255) * (1 / (std*255)), float32
`resize_linear_u8` below restates OpenCV's 8-bit bilinear path in integer arithmetic (11-bit fixed-point weights);
tests/test_oracle_golden.py pins it bit-exactly against cv2.resize itself.
"""
from __future__ import annotations

import numpy as np

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def _coef(dst: int, src: int, horizontal: bool):
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * (src / dst) - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if horizontal:                                  # OpenCV zeroes the fraction at the left/right borders only
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    a0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int64)
    a1 = np.rint(f * np.float32(2048)).astype(np.int64)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), a0, a1


def resize_linear_u8(img: np.ndarray, r: int) -> np.ndarray:
    """cv2.resize(img, (r, r), interpolation=cv2.INTER_LINEAR) for uint8 HWC images, in integer arithmetic."""
    h, w, _ = img.shape
    if h == r and w == r:
        return img.copy()
    x0, x1, a0, a1 = _coef(r, w, True)
    y0, y1, b0, b1 = _coef(r, h, False)
    im = img.astype(np.int64)
    rows0 = im[y0][:, x0] * a0[None, :, None] + im[y0][:, x1] * a1[None, :, None]
    rows1 = im[y1][:, x0] * a0[None, :, None] + im[y1][:, x1] * a1[None, :, None]
    out = (((b0[:, None, None] * (rows0 >> 4)) >> 16) + ((b1[:, None, None] * (rows1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def preprocess(img_u8: np.ndarray, r: int) -> np.ndarray:
    """uint8 HWC -> float32 CHW [3, r, r] exactly as the reference's transforms produce it."""
    x = resize_linear_u8(img_u8, r).astype(np.float32)
    x -= np.array(IMAGENET_MEAN, dtype=np.float32) * np.float32(255.0)
    x *= np.reciprocal(np.array(IMAGENET_STD, dtype=np.float32) * np.float32(255.0))
    return np.ascontiguousarray(x.transpose(2, 0, 1))
