"""Batched, streaming inference over a folder of images: the caller side of the hot path that the reference's
`visualize.py` implements one image at a time (dt_segmentation/visualize.py:21-54: glob *.jpg then *.png,
`Image.open(...).convert('RGB')`, `predict`, overlay, save).

Here the images are grouped into batches of equal frame size and go through the pipelined host entry point
(`DINOSeg.predict_batch_async` on raw uint8 frames: resize + normalise + forward + argmax + replication on the GPU), with
two batches in flight so that decoding / copying batch k+1 overlaps the GPU work of batch k.  Results are yielded in
the reference's order and are the maps `predict()` returns for each image.
"""
from __future__ import annotations

import glob
import os
from typing import Iterable, Iterator, List, Sequence, Tuple

import numpy as np
import torch


def list_images(image_dir: str) -> List[str]:
    """The reference's traversal order (visualize.py:36-37): every *.jpg, then every *.png, in glob order."""
    files: List[str] = []
    for ext in ("jpg", "png"):
        files.extend(glob.glob(os.path.join(image_dir, f"*.{ext}")))
    return files


def plan_batches(shapes: Sequence[Tuple[int, int]], batch_size: int) -> List[List[int]]:
    """Indices of consecutive images that share a frame size, cut into batches of at most `batch_size`.
    Keeping the batches consecutive keeps the output in input order without a reorder buffer."""
    if batch_size < 1:
        raise ValueError("batch_size must be >= 1")
    out: List[List[int]] = []
    for i, hw in enumerate(shapes):
        if out and len(out[-1]) < batch_size and shapes[out[-1][0]] == hw:
            out[-1].append(i)
        else:
            out.append([i])
    return out


def load_rgb(path: str) -> np.ndarray:
    """visualize.py:38-40: PIL decode, convert('RGB') -> uint8 [H, W, 3]."""
    from PIL import Image
    with open(path, "rb") as f:
        return np.asarray(Image.open(f).convert("RGB"))


def predict_images(model, images: Iterable[np.ndarray], batch_size: int = 32, resolution: int | None = None
                   ) -> Iterator[np.ndarray]:
    """Label maps (int64 [g*p, g*p] each, what `model.predict(img)` returns) for a stream of uint8 RGB images, in order.
    Two batches are kept in flight on the GPU."""
    res = int(model.resolution if resolution is None else resolution)
    pending = []                                   # [(ticket, n_images)]

    def drain(limit):
        while len(pending) > limit:
            ticket, n = pending.pop(0)
            labels = model.predict_wait(ticket)
            for k in range(n):
                yield labels[k]

    batch: List[np.ndarray] = []

    def submit():
        frames = torch.from_numpy(np.stack(batch))
        if torch.cuda.is_available():
            frames = frames.pin_memory()
        pending.append((model.predict_batch_async(frames, resolution=res, output="labels"), len(batch)))
        batch.clear()

    for img in images:
        img = np.asarray(img)
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError(f"expected uint8 RGB images [H, W, 3], got {img.dtype} {img.shape}")
        if batch and (len(batch) == batch_size or batch[0].shape != img.shape):
            submit()
            yield from drain(1)
        batch.append(img)
    if batch:
        submit()
    yield from drain(0)


def predict_folder(model, image_dir: str, batch_size: int = 32, resolution: int | None = None
                   ) -> Iterator[Tuple[str, np.ndarray, np.ndarray]]:
    """(path, rgb image, label map) for every image of `image_dir` in the reference's order."""
    files = list_images(image_dir)
    images: List[np.ndarray] = []

    def gen():
        for f in files:
            img = load_rgb(f)
            images.append(img)
            yield img

    for k, pred in enumerate(predict_images(model, gen(), batch_size, resolution)):
        yield files[k], images[k], pred
        images[k] = None                           # decoded frames are dropped as soon as they have been handed out


# ---------------------------------------------------------------------------------------------
# overlay (visualize.py:46-54 uses imgviz.label2rgb over the grey image; imgviz is an optional dependency)
# ---------------------------------------------------------------------------------------------
def label_colormap(n: int = 256) -> np.ndarray:
    """The PASCAL-VOC label colour map (bit-interleaved class index), the table imgviz.label_colormap() generates."""
    cmap = np.zeros((n, 3), dtype=np.uint8)
    for i in range(n):
        c, r, g, b = i, 0, 0, 0
        for j in range(8):
            r |= ((c >> 0) & 1) << (7 - j)
            g |= ((c >> 1) & 1) << (7 - j)
            b |= ((c >> 2) & 1) << (7 - j)
            c >>= 3
        cmap[i] = (r, g, b)
    return cmap


def overlay(pred: np.ndarray, rgb: np.ndarray, alpha: float = 0.5) -> np.ndarray:
    """Label colours blended over the grey-scale image resized to the label map (class 0 keeps the grey image, as
    imgviz.label2rgb does for the background).  uint8 [H, W, 3]."""
    from PIL import Image
    try:
        import imgviz                                # the reference's renderer, when it is installed
        grey = imgviz.rgb2gray(np.array(Image.fromarray(rgb).resize(pred.shape[::-1])))
        return imgviz.label2rgb(pred, grey, font_size=15, loc="rb")
    except ImportError:
        pass
    h, w = pred.shape
    g = np.asarray(Image.fromarray(rgb).resize((w, h)).convert("L"), dtype=np.float32)[..., None].repeat(3, axis=2)
    colour = label_colormap()[np.clip(pred, 0, 255)].astype(np.float32)
    out = np.where((pred > 0)[..., None], (1.0 - alpha) * g + alpha * colour, g)
    return out.round().clip(0, 255).astype(np.uint8)
