#!/usr/bin/env python
"""Run inference on a folder of images and save the visualisations in target_dir: the reference's
dt_segmentation/visualize.py (same command line, same traversal order, same `inference` signature plus `batch_size`),
batched and pipelined on the B200 path instead of one `predict` call per image."""
import argparse
import os

from dino_b200 import DINOSeg
from dino_b200 import folder


def inference(checkpoint_path, image_dir, target_dir, labels_path=None, resolution=480, cpu=False, batch_size=32):
    """Use a trained PL checkpoint to run inference on all images in image_dir (visualize.py:21-54)."""
    if cpu:
        raise RuntimeError("the B200 build has no CPU path (DINOSeg runs on CUDA only)")
    model = DINOSeg.load_from_checkpoint(checkpoint_path).to("cuda:0")
    model.set_resolution(resolution)        # only the inference resolution; the output is still 480 x 480
    os.makedirs(target_dir, exist_ok=True)
    from PIL import Image
    n = 0
    for path, rgb, pred in folder.predict_folder(model, image_dir, batch_size=batch_size, resolution=resolution):
        Image.fromarray(folder.overlay(pred, rgb)).save(os.path.join(target_dir, path.split(os.sep)[-1]))
        n += 1
    return n


if __name__ == "__main__":
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("checkpoint_path", help="Trained PL checkpoint")
    parser.add_argument("image_dir", help="Images to run inference on")
    parser.add_argument("target_dir", help="Where to save predictions")
    parser.add_argument("--labels_path", help="Txt file with class labels.", required=False,
                        default=os.path.join("data", "labels.txt"))
    parser.add_argument("--resolution", help="Prediction resolutions.", required=False, default=480, type=int)
    parser.add_argument("--cpu", help="Force usage of cpu.", required=False, action="store_true")
    parser.add_argument("--batch_size", help="Images per GPU batch.", required=False, default=32, type=int)
    args = parser.parse_args()
    print(inference(**vars(args)), "images")
