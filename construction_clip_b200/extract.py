"""Batched prefix-feature extraction: the work of CLIP_prefix_caption/parse_coco.py:37-65 with every
image encoded ONCE and the prompt embeddings encoded ONCE (the reference runs the vision tower three
times per image and re-encodes the 2 + 9 prompts for every image).

Output is the pickle the ClipCap trainer consumes (CLIP_prefix_caption/train.py:72-76):
    {"clip_embedding": Tensor[N, E] (model dtype, un-normalised encode_image output),
     "captions": [annotation dicts + "clip_embedding": row index + "attribute": "<caption type> <violation type> "]}
with the zero-shot decision rule of parse_coco.py:45-53: argmax over softmax(logit_scale * I.T^T).
"""
from __future__ import annotations

import pickle
from typing import Iterable, Sequence

import torch

from . import ops as O

# parse_coco.py:24-28
CAPTION_TYPES = {"status": "現況", "violation": "缺失"}
VIOLATION_TYPES = ["墜落", "防護具", "感電", "工作場所", "物料", "爆炸", "穿刺", "機械", "搬運"]


@torch.no_grad()
def _normalised(feat_f32):
    y, _ = O.l2norm_fwd(feat_f32.contiguous())
    return y


@torch.no_grad()
def extract_prefix_features(model, images: Iterable[torch.Tensor], annotations: Sequence[dict],
                            caption_type_tokens: torch.Tensor, violation_type_tokens: torch.Tensor,
                            caption_type_labels: Sequence[str] = tuple(CAPTION_TYPES.values()),
                            violation_type_labels: Sequence[str] = tuple(VIOLATION_TYPES), batch_size: int = 512,
                            out_path: str | None = None) -> dict:
    """``images`` yields preprocessed [3,R,R] (or [n,3,R,R]) tensors in annotation order; the token
    tensors are ``clip.tokenize(...)`` of the two prompt lists (parse_coco.py:29-30)."""
    dev = next(model.parameters()).device
    ls = model.logit_scale.detach().float().reshape(1).contiguous()
    txt_c = _normalised(model._features("text", caption_type_tokens.to(dev)))      # encoded once
    txt_v = _normalised(model._features("text", violation_type_tokens.to(dev)))
    annotations = [dict(a) for a in annotations]
    embeddings, row = [], 0

    def flush(batch):
        nonlocal row
        x = torch.cat(batch, 0).to(dev, non_blocking=True)
        feat = model._features("visual", x)                 # fp32 [n, E]: one vision-tower pass per image
        img_n = _normalised(feat)
        idx_c = O.logits(img_n, txt_c, ls).softmax(dim=-1).argmax(dim=1).tolist()
        idx_v = O.logits(img_n, txt_v, ls).softmax(dim=-1).argmax(dim=1).tolist()
        embeddings.append(feat.to(model.dtype))
        for a, b in zip(idx_c, idx_v):
            annotations[row]["clip_embedding"] = row
            annotations[row]["attribute"] = f"{caption_type_labels[a]} {violation_type_labels[b]} "
            row += 1

    pending, n_pending = [], 0
    for img in images:
        img = img if img.dim() == 4 else img.unsqueeze(0)
        pending.append(img)
        n_pending += img.shape[0]
        if n_pending >= batch_size:
            flush(pending)
            pending, n_pending = [], 0
    if pending:
        flush(pending)
    if row != len(annotations):
        raise RuntimeError(f"{row} images for {len(annotations)} annotations")
    result = {"clip_embedding": torch.cat(embeddings, 0), "captions": annotations}
    if out_path is not None:
        with open(out_path, "wb") as fh:
            pickle.dump(result, fh)
    return result
