"""TEST INFRASTRUCTURE (oracle): numpy restatement of the image preprocessing the reference feeds CLIP with --
``preprocess(Image.open(...))`` at CLIP/train.py:56 and CLIP/predict.py:31, i.e. upstream ``clip._transform``:
``Resize(n_px, BICUBIC)`` -> ``CenterCrop(n_px)`` -> ``convert("RGB")`` (-> ToTensor -> Normalize, which the B200
path fuses into the patch-embedding im2col).  The arithmetic lives in two third-party packages that ARE installed
here, so this restatement is pinned against them (tests/test_cpu.py::test_resize_oracle_matches_pil):

* torchvision 0.26 ``transforms.functional``: output size of ``Resize(int)`` (``_compute_resized_output_size``:
  short edge -> n_px, long edge -> int(n_px * long / short)) and the crop offsets of ``CenterCrop``
  (``int(round((h - th) / 2.0))``, Python's round-half-to-even).
* Pillow 12.2 ``src/libImaging/Resample.c`` (8 bits per channel): ``precompute_coeffs`` (support = 2 * max(scale, 1)
  for the bicubic filter with a = -0.5, weights normalised to sum 1), ``normalize_coeffs_8bpc`` (fixed point,
  22 fractional bits, round half away from zero), then a HORIZONTAL pass and a VERTICAL pass, each accumulating in
  int32 from ``1 << 21``, shifting right by 22 and clipping to 0..255; a pass whose output size equals its input size
  is skipped.

Only ``tests/`` may import this module; the product's host logic (construction_clip_b200/data.py) has its own
implementation of the coefficient tables, which the tests compare with this one.
"""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole axis (box = the full input).
    -> (bounds int32 [out, 2] = (xmin, count), coeffs int32 [out, ksize])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = max(xmin, 0)
        xmax = int(center + support + 0.5)
        xmax = min(xmax, in_size)
        xmax -= xmin
        ww = 0.0
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, np.trunc(-0.5 + kk * (1 << PRECISION_BITS)), np.trunc(0.5 + kk * (1 << PRECISION_BITS)))
    return bounds, fixed.astype(np.int32)


def _pass(img: np.ndarray, out_size: int, axis: int) -> np.ndarray:
    """One resample pass of a uint8 [H, W, C] image along `axis` (0 = vertical, 1 = horizontal)."""
    in_size = img.shape[axis]
    if in_size == out_size:
        return img
    bounds, coeffs = precompute_coeffs(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        xmin, n = bounds[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(coeffs[xx, :n].astype(np.int64), src[xmin:xmin + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resized_size(h: int, w: int, n_px: int):
    """torchvision Resize(int): (new_h, new_w)."""
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = n_px, int(n_px * long / short)
    return (new_long, new_short) if w <= h else (new_short, new_long)


def crop_offsets(h: int, w: int, n_px: int):
    """torchvision CenterCrop: (top, left)."""
    return int(round((h - n_px) / 2.0)), int(round((w - n_px) / 2.0))


def preprocess_uint8(rgb: np.ndarray, n_px: int) -> np.ndarray:
    """uint8 RGB [H, W, 3] -> uint8 [3, n_px, n_px], what upstream's transform yields before ToTensor."""
    h, w, _ = rgb.shape
    oh, ow = resized_size(h, w, n_px)
    res = _pass(_pass(rgb, ow, 1), oh, 0)   # horizontal, then vertical (ImagingResample)
    top, left = crop_offsets(oh, ow, n_px)
    return np.ascontiguousarray(res[top:top + n_px, left:left + n_px].transpose(2, 0, 1))
