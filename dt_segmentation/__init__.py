"""Drop-in import path of the reference package (reference dt_segmentation/__init__.py:1-2,
README.md:26): `from dt_segmentation import DINOSeg` resolves to the B200 implementation."""
from dino_b200 import DINOSeg  # noqa: F401


def parse_class_names(path):
    """Reference dt_utils.py:117-130 (labels.txt -> (class_names, class_name_to_id)); pure
    host-side bookkeeping used by callers of predict()."""
    class_names = []
    class_name_to_id = {}
    with open(path) as f:
        for i, line in enumerate(f.readlines()):
            class_id = i - 1  # the first line is __ignore__ (-1)
            name = line.strip()
            class_name_to_id[name] = class_id
            if class_id == -1:
                assert name == "__ignore__"
                continue
            if class_id == 0:
                assert name == "_background_"
            class_names.append(name)
    return tuple(class_names), class_name_to_id


__all__ = ["DINOSeg", "parse_class_names"]
