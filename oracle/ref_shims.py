"""Import the UNMODIFIED reference `dt_segmentation` package from /root/reference.

TEST INFRASTRUCTURE (used by oracle/make_golden.py in the build container only; the GPU box has
no /root/reference and nothing at run time imports this).

Two packages the reference imports are absent offline:
  * pytorch_lightning -> stub whose LightningModule is an nn.Module with a no-op
    save_hyperparameters() and a `device` property;
  * albumentations    -> stub exposing Compose/Resize/Normalize + albumentations.pytorch.ToTensorV2
    backed by dino_b200.transforms (a restatement; preprocessing parity is unpinned).
`get_dino` (dt_utils.py:19-29) downloads weights; it is replaced by the same constructor call
without the download.  No reference file is copied or modified.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DINO_REFERENCE_ROOT", "/root/reference")


def _install_stubs():
    import torch
    from torch import nn

    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class LightningModule(nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            @property
            def device(self):
                try:
                    return next(self.parameters()).device
                except StopIteration:
                    return torch.device("cpu")

            def log(self, *a, **k):
                pass

        class Trainer:
            def __init__(self, *a, **k):
                raise RuntimeError("training is out of scope")

        pl.LightningModule = LightningModule
        pl.Trainer = Trainer
        cb = types.ModuleType("pytorch_lightning.callbacks")
        es = types.ModuleType("pytorch_lightning.callbacks.early_stopping")
        mc = types.ModuleType("pytorch_lightning.callbacks.model_checkpoint")
        es.EarlyStopping = type("EarlyStopping", (), {"__init__": lambda self, *a, **k: None})
        mc.ModelCheckpoint = type("ModelCheckpoint", (), {"__init__": lambda self, *a, **k: None})
        pl.callbacks = cb
        cb.early_stopping, cb.model_checkpoint = es, mc
        sys.modules.update({"pytorch_lightning": pl, "pytorch_lightning.callbacks": cb,
                            "pytorch_lightning.callbacks.early_stopping": es,
                            "pytorch_lightning.callbacks.model_checkpoint": mc})

    if "albumentations" not in sys.modules:
        here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        if here not in sys.path:
            sys.path.insert(0, here)
        from dino_b200 import transforms as T

        A = types.ModuleType("albumentations")

        class _Op:
            def __init__(self, *a, **k):
                self.args, self.kwargs = a, k

        class Resize(_Op):
            pass

        class Normalize(_Op):
            pass

        class ToTensorV2(_Op):
            pass

        class Compose:
            def __init__(self, ops):
                res = [o for o in ops if isinstance(o, Resize)]
                self._t = T.Compose(res[0].args[0]) if res else None

            def __call__(self, **kw):
                return self._t(**kw)

        A.Compose, A.Resize, A.Normalize = Compose, Resize, Normalize
        for name in ("RandomResizedCrop", "ShiftScaleRotate", "HorizontalFlip", "ColorJitter", "GaussianBlur"):
            setattr(A, name, type(name, (_Op,), {}))
        ap = types.ModuleType("albumentations.pytorch")
        ap.ToTensorV2 = ToTensorV2
        A.pytorch = ap
        sys.modules.update({"albumentations": A, "albumentations.pytorch": ap})


def import_reference():
    """Returns (pl_torch_modules, vision_transformer) of the reference."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"{REFERENCE_ROOT} not present (the reference only exists in the build container)")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # our own drop-in package is also called dt_segmentation: make sure the reference's wins here
    for k in [k for k in sys.modules if k == "dt_segmentation" or k.startswith("dt_segmentation.")]:
        del sys.modules[k]
    import importlib
    vt = importlib.import_module("dt_segmentation.src.vision_transformer")
    plm = importlib.import_module("dt_segmentation.src.pl_torch_modules")
    plm.get_dino = lambda patch_size=8, device="cpu": vt.vit_small(patch_size=patch_size, num_classes=0)
    return plm, vt
