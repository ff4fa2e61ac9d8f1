"""The step before the hot path: batch composition and image wire format.

``CombinationBatches`` reproduces the batch composition of the reference's ``ClipPairDataset``
(CLIP/train.py:36-91): one "item" IS one training batch -- one (image, text) pair per class of a
class combination, so that the contrastive labels ``arange(B)`` are unambiguous -- and the index
arithmetic (50 items per combination, ``item % len(class list)``) is kept exactly.  It returns
annotation records; decoding is left to ``load`` so that the composition can be tested without
image files.

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
