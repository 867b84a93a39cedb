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


def grayscale_u8(im) -> np.ndarray:
    """``DynamicImage::grayscale()`` followed by ``get_pixel_matrix`` (rcn.rs:83,398 + lib.rs:27-41) for a decoded Pillow
    image -> (H, W) uint8.

    The ``image`` crate (0.24, rcn/Cargo.toml; not vendored -- formula from its published source, unverifiable here) turns
    Rgb8 / Rgba8 into Luma8 / LumaA8 with the integer sRGB luminance ``(2126 R + 7152 G + 722 B) / 10000`` (truncating
    division), NOT the Rec.601 weights of Pillow's ``convert("L")``; palette PNGs decode to RGB(A) first; 1-bit grey decodes
    to 0 / 255. 16-bit (and any other) colour types stay 16-bit through ``grayscale()``, which ``get_pixel_matrix``
    rejects with ``InvalidGrayscaleImageError`` (lib.rs:39) -- same here instead of a silent clip."""
    mode = im.mode
    if mode == "P":
        im = im.convert("RGBA" if "transparency" in im.info else "RGB")
        mode = im.mode
    if mode == "1":
        im = im.convert("L")
        mode = "L"
    if mode == "L":
        return np.asarray(im, dtype=np.uint8)
    if mode == "LA":
        return np.ascontiguousarray(np.asarray(im, dtype=np.uint8)[:, :, 0])
    if mode in ("RGB", "RGBA"):
        a = np.asarray(im, dtype=np.uint8).astype(np.uint32)
        return ((2126 * a[..., 0] + 7152 * a[..., 1] + 722 * a[..., 2]) // 10000).astype(np.uint8)
    raise InvalidGrayscaleImageError()


def load_grayscale(path: str) -> np.ndarray:
    """``ImageReader::open(path)?.decode()?.grayscale()`` + ``get_pixel_matrix`` (rcn.rs:83,394-400) -> (H, W) uint8."""
    from PIL import Image
    with Image.open(path) as im:
        return grayscale_u8(im)


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
