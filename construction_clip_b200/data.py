"""The step before the hot path: batch composition and image wire format.

``CombinationBatches`` reproduces the batch composition of the reference's ``ClipPairDataset``
(CLIP/train.py:36-91): one "item" IS one training batch -- one (image, text) pair per class of a
class combination, so that the contrastive labels ``arange(B)`` are unambiguous -- and the index
arithmetic (50 items per combination, ``item % len(class list)``) is kept exactly.  It returns
annotation records; decoding is left to ``load`` so that the composition can be tested without
image files.

``GpuPreprocess`` moves the bicubic resize + centre crop of that transform onto the GPU as well
(``b200clip_resize_crop_u8``, bit-exact with Pillow's resampler): the host only decodes the file and ships the
pixels once, as uint8, at their original size.

``preprocess_uint8`` is upstream's ``_transform`` (bicubic resize, centre crop, RGB) WITHOUT
ToTensor + Normalize: it yields uint8 ``[3, R, R]`` pixels, a quarter of the host->device bytes.
The B200 path normalises inside the patch-embedding im2col (``b200clip_im2col_patch`` with
``B200CLIP_DT_U8``), so ``model(image_uint8, text)`` / ``ClipTrainer.step_from_host`` give the
same features as the fp32 route (tests/test_model_gpu.py::test_uint8_pixels_with_fused_normalize).
"""
from __future__ import annotations

import collections
import itertools
import os

import torch

ITEMS_PER_COMBINATION = 50   # CLIP/train.py:87 (`self.cumulative_sizes = [50 for p in self.pair_list]`)


def preprocess_uint8(n_px: int):
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, PILToTensor, Resize
    return Compose([
        Resize(n_px, interpolation=InterpolationMode.BICUBIC),
        CenterCrop(n_px),
        lambda image: image.convert("RGB"),
        PILToTensor(),   # uint8 [3, n_px, n_px]; ToTensor + Normalize happen on the GPU
    ])


# ------------------------------------------------------------------------------------------------
# The same transform with the resize + crop on the GPU (b200clip_resize_crop_u8): the host only decodes.
_PRECISION_BITS = 22   # Pillow: 32 - 8 - 2


def _bicubic(x):
    """Pillow's bicubic filter (a = -0.5), elementwise on a float64 array, same operation order."""
    import numpy as np
    a = -0.5
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def axis_taps(in_size: int, out_size: int, first: int, count: int):
    """Fixed-point taps of Pillow's 8-bit resampler (Resample.c precompute_coeffs + normalize_coeffs_8bpc) for output
    positions first .. first + count - 1 of an axis resized in_size -> out_size.
    -> (bounds int32 [count, 2] = (first source index, taps), coeffs int32 [count, ksize])."""
    import math
    import numpy as np
    pos = np.arange(first, first + count, dtype=np.int64)
    if in_size == out_size:   # Pillow skips the pass: identity taps reproduce that exactly
        return (np.stack([pos, np.ones_like(pos)], 1).astype(np.int32),
                np.full((count, 1), 1 << _PRECISION_BITS, np.int32))
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    center = (pos + 0.5) * scale
    lo = np.maximum(np.trunc(center - support + 0.5).astype(np.int64), 0)          # C's (int) cast truncates
    n = np.minimum(np.trunc(center + support + 0.5).astype(np.int64), in_size) - lo
    j = np.arange(ksize, dtype=np.int64)[None, :]
    w = _bicubic(((j + lo[:, None]) - center[:, None] + 0.5) * (1.0 / filterscale))
    w = np.where(j < n[:, None], w, 0.0)
    total = np.cumsum(w, axis=1)[:, -1:]     # sequential accumulation, like the C loop (trailing zeros add nothing)
    w = np.where(total != 0.0, w / np.where(total != 0.0, total, 1.0), w)
    one = float(1 << _PRECISION_BITS)
    fixed = np.where(w < 0, np.trunc(-0.5 + w * one), np.trunc(0.5 + w * one)).astype(np.int32)
    return np.stack([lo, n], 1).astype(np.int32), fixed


class ResizePlan:
    """Everything `b200clip_resize_crop_u8` needs for one input size: torchvision's Resize(int) output size and
    CenterCrop offsets, and the taps of the surviving n_px output columns / rows."""

    def __init__(self, h: int, w: int, n_px: int):
        short, long = (w, h) if w <= h else (h, w)
        new_long = int(n_px * long / short)
        self.oh, self.ow = (new_long, n_px) if w <= h else (n_px, new_long)
        self.top, self.left = int(round((self.oh - n_px) / 2.0)), int(round((self.ow - n_px) / 2.0))
        self.n_px = n_px
        self.xbounds, self.xcoef = axis_taps(w, self.ow, self.left, n_px)
        self.ybounds, self.ycoef = axis_taps(h, self.oh, self.top, n_px)
        self.row0 = int(self.ybounds[:, 0].min())
        self.rows = int((self.ybounds[:, 0] + self.ybounds[:, 1]).max()) - self.row0


class GpuPreprocess:
    """``preprocess_uint8`` with the bicubic resize + centre crop on the GPU.  Call it with a decoded image -- a PIL
    image, or uint8 RGB ``[H, W, 3]`` pixels as a numpy array / torch tensor (host or device) -- and get the uint8
    ``[3, n_px, n_px]`` CUDA tensor of ``preprocess_uint8`` back, bit for bit (tests/test_model_gpu.py::
    test_gpu_preprocess_matches_pil).  One difference to upstream: a PIL image is converted to RGB BEFORE the resize
    (upstream converts after it); identical for RGB inputs such as the reference's JPEGs."""

    def __init__(self, n_px: int, device="cuda"):
        self.n_px = n_px
        self.device = torch.device(device)
        self._plans: dict = {}

    def _plan(self, h: int, w: int):
        key = (h, w)
        if key not in self._plans:
            p = ResizePlan(h, w, self.n_px)
            dev = [torch.from_numpy(a).to(self.device) for a in (p.xbounds, p.xcoef, p.ybounds, p.ycoef)]
            self._plans[key] = (p, dev)
        return self._plans[key]

    def __call__(self, image) -> torch.Tensor:
        import numpy as np
        from . import ops as O
        if hasattr(image, "convert"):   # PIL
            image = np.array(image.convert("RGB"))   # a writable copy (torch.from_numpy refuses read-only buffers)
        if isinstance(image, np.ndarray):
            image = torch.from_numpy(np.ascontiguousarray(image))
        if image.dtype != torch.uint8 or image.dim() != 3 or image.shape[2] != 3:
            raise RuntimeError("GpuPreprocess: expected uint8 RGB pixels [H, W, 3]")
        src = image.to(self.device, non_blocking=True).contiguous()
        p, (xb, xc, yb, yc) = self._plan(int(src.shape[0]), int(src.shape[1]))
        return O.resize_crop_u8(src, xb, xc, yb, yc, p.row0, p.rows, self.n_px)

    def batch(self, images) -> torch.Tensor:
        """-> uint8 [B, 3, n_px, n_px]."""
        return torch.stack([self(im) for im in images])


class CombinationBatches:
    """Batch composition of CLIP/train.py:36-91.

    annotations: list of dicts (``data["annotations"]`` of the reference's json); ``key``: the caption field
    (``'violation_type'`` / ``'caption_type'``, CLIP/train.py:121); classes = distinct non-empty values of
    that field in order of first appearance; every ``combination_num``-subset of the classes (itertools
    order) contributes ``ITEMS_PER_COMBINATION`` batches; a class's records are split train / test at
    ``int(count * train_ratio)``."""

    def __init__(self, annotations, key: str, combination_num: int, train_ratio: float = 0.8, split: str = "train"):
        if split not in ("train", "test"):
            raise ValueError("split must be 'train' or 'test'")
        records = [a for a in annotations if a[key] != ""]
        counts = collections.Counter(a[key] for a in records)
        self.key = key
        self.classes = list(counts.keys())
        self.combinations = list(itertools.combinations(self.classes, combination_num))
        cut = {k: int(n * train_ratio) for k, n in counts.items()}
        by_class = {k: [a for a in records if a[key] == k] for k in self.classes}
        self.lists = {k: (v[:cut[k]] if split == "train" else v[cut[k]:]) for k, v in by_class.items()}

    def __len__(self) -> int:
        return ITEMS_PER_COMBINATION * len(self.combinations)

    def __getitem__(self, item: int):
        """-> the batch's annotation records, one per class of the combination."""
        if not 0 <= item < len(self):
            raise IndexError(item)
        combo = self.combinations[item // ITEMS_PER_COMBINATION]
        offset = item % ITEMS_PER_COMBINATION
        batch = []
        for k in combo:
            rows = self.lists[k]
            if not rows:
                raise RuntimeError(f"class {k!r} has no records in this split (the reference divides by zero here)")
            batch.append(rows[offset % len(rows)])
        return batch

    def load(self, item: int, preprocess, tokenize, image_root: str = ""):
        """Decodes one batch like ClipPairDataset.__getitem__: (images [B,3,R,R], tokens [B,77])."""
        from PIL import Image
        batch = self[item]
        images = torch.stack([torch.as_tensor(preprocess(Image.open(os.path.join(image_root, a["file_name"]))))
                              for a in batch])
        return images, tokenize([a[self.key] for a in batch])
