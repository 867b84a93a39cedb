"""Image -> matrix adapter and class-directory loader (rcn/src/lib.rs:27-41, rcn/src/rcn.rs:367-404).

PNG decoding itself is outside the accelerated path: Pillow decodes on the host and the
u8 pixel buffers go to the GPU feature kernel.
"""
from __future__ import annotations

import os

import numpy as np


class InvalidGrayscaleImageError(ValueError):
    """errors.rs:1-13"""

    def __str__(self):
        return "InvalidGrayscaleImageError: Image provided was not Luma8 (grayscaled image)"


def get_pixel_matrix(image) -> np.ndarray:
    """lib.rs:27-41: Luma8 / LumaA8 -> H x W float64 matrix of 0..255 values; anything else is an error."""
    mode = getattr(image, "mode", None)
    if mode == "L":
        return np.asarray(image, dtype=np.uint8).astype(np.float64)
    if mode == "LA":
        return np.asarray(image, dtype=np.uint8)[:, :, 0].astype(np.float64)
    raise InvalidGrayscaleImageError()


def load_grayscale(path: str) -> np.ndarray:
    """``ImageReader::open(path)?.decode()?.grayscale()`` (rcn.rs:83,394-398) -> (H, W) uint8."""
    from PIL import Image
    with Image.open(path) as im:
        return np.asarray(im.convert("L"), dtype=np.uint8)


def load_data(path: str, class_size_limit: int, rng: np.random.Generator):
    """rcn.rs:367-404 without the feature/standardise part: sorted class directories, ``class_size_limit`` files
    drawn without replacement per class (seeded generator instead of thread_rng), decoded to grayscale.
    Returns (images (N, H, W) uint8, labels (N,) int64)."""
    classes = sorted(os.path.join(path, d) for d in os.listdir(path))
    images, labels = [], []
    for i, cdir in enumerate(classes):
        paths = [os.path.join(cdir, f) for f in os.listdir(cdir)]
        if class_size_limit > len(paths):  # rcn.rs:383-390
            raise ValueError("provided class_size_limit for {} too large! expected {} <= {}".format(
                path, class_size_limit, len(paths)))
        for _ in range(class_size_limit):
            idx = int(rng.integers(0, len(paths)))
            images.append(load_grayscale(paths.pop(idx)))
            labels.append(i)
    return np.stack(images), np.asarray(labels, dtype=np.int64)
