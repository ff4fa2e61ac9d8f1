"""`clip.load`, `clip.tokenize`, `clip.available_models`, `_transform` -- the entry points the
reference calls (CLIP/predict.py:12,31,40; CLIP/train.py:56,60,105; parse_coco.py:20,29-30).

Differences from upstream, all forced by the environment and documented in DESIGN.md:
  * no network: weights are looked up in `download_root` (default ~/.cache/clip) under upstream's
    file names; if absent the model is RANDOM-INITIALISED with upstream's scheme and a warning is
    printed (set CLIP_B200_REQUIRE_WEIGHTS=1 to make that an error).  The reference scripts
    load their own fine-tuned state dict right after `clip.load` anyway.
  * on CUDA the kernels compute in bf16 with fp32 accumulation (upstream: fp16), reading a bf16 shadow of
    the weights, while the nn.Parameters returned by `clip.load` stay **fp32**: the reference fine-tunes
    with `AdamW(model.parameters(), lr=1e-5)` (CLIP/train.py:143), and an update of 1e-5 is below half a
    bf16 ulp for almost every weight -- bf16 parameters would silently not train.  The shadow is
    refreshed (one multi-tensor cast) whenever a parameter's version counter moves.
    `clip.model.convert_weights(model)` still gives bf16 parameters that alias the shadow (zero copy,
    inference / ClipTrainer with its own fp32 master weights).  There is no CPU execution path --
    `device="cpu"` builds the module (state_dict round trips work) but calling it raises.
  * ViT models only (BASELINE config 4: "RN-free"); `jit=True` is not supported.


Attribution: the algorithm, constants and public names here follow openai/CLIP's `clip/clip.py (`tokenize`, `_transform`, `load`)` (MIT License,
Copyright (c) 2021 OpenAI) -- byte-exact behaviour is the contract of this boundary (token ids, pixel
normalisation constants); the file is a re-implementation kept under the same MIT terms.
"""
from __future__ import annotations

import os
import warnings
from typing import List, Union

import torch

from construction_clip_b200.model import CLIP, CONFIGS, build_model

from .simple_tokenizer import SimpleTokenizer as _Tokenizer

__all__ = ["available_models", "load", "tokenize"]

_MODEL_FILES = {
    "ViT-B/32": "ViT-B-32.pt",
    "ViT-B/16": "ViT-B-16.pt",
    "ViT-L/14": "ViT-L-14.pt",
    "ViT-L/14@336px": "ViT-L-14-336px.pt",
}

_tokenizer = None


def available_models() -> List[str]:
    return list(_MODEL_FILES.keys())


def _convert_image_to_rgb(image):
    return image.convert("RGB")


def _transform(n_px: int):
    from torchvision.transforms import CenterCrop, Compose, InterpolationMode, Normalize, Resize, ToTensor
    return Compose([
        Resize(n_px, interpolation=InterpolationMode.BICUBIC),
        CenterCrop(n_px),
        _convert_image_to_rgb,
        ToTensor(),
        Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711)),
    ])


def _read_checkpoint(path: str) -> dict:
    try:  # upstream ships TorchScript archives
        return torch.jit.load(path, map_location="cpu").state_dict()
    except RuntimeError:
        sd = torch.load(path, map_location="cpu")
        return sd.get("state_dict", sd) if isinstance(sd, dict) else sd.state_dict()


def load(name: str, device: Union[str, torch.device] = "cuda" if torch.cuda.is_available() else "cpu",
         jit: bool = False, download_root: str | None = None, pretrained: bool | None = None):
    """Returns ``(model, preprocess)`` like upstream.  ``pretrained=False`` forces random init."""
    if jit:
        raise RuntimeError("jit=True is not supported by the B200 CLIP path; use jit=False")
    if name in _MODEL_FILES:
        path = os.path.join(download_root or os.path.expanduser("~/.cache/clip"), _MODEL_FILES[name])
    elif os.path.isfile(name):
        path = name
    elif name in ("RN50", "RN101", "RN50x4", "RN50x16", "RN50x64"):
        raise RuntimeError(f"Model {name} is a ResNet CLIP; only ViT models are supported: {available_models()}")
    else:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")

    if pretrained is not False and os.path.isfile(path):
        model = build_model(_read_checkpoint(path))
    else:
        if pretrained or os.environ.get("CLIP_B200_REQUIRE_WEIGHTS") == "1":
            raise RuntimeError(f"no checkpoint at {path} and no network to download it")
        if name not in CONFIGS:
            raise RuntimeError(f"no checkpoint at {path}")
        if pretrained is None:
            warnings.warn(f"clip.load: no checkpoint at {path} (offline); {name} is RANDOM-INITIALISED with upstream's "
                          "initialize_parameters scheme -- load a state dict before use")
        model = CLIP(CONFIGS[name])

    device = torch.device(device)
    model = model.float().to(device)   # fp32 master parameters on either device (see the module docstring)
    model.eval()
    return model, _transform(model.visual.input_resolution)


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False) -> torch.Tensor:
    """[SOT] + BPE(text) + [EOT], zero padded to ``context_length``; raises if too long."""
    global _tokenizer
    if isinstance(texts, str):
        texts = [texts]
    if _tokenizer is None:
        _tokenizer = _Tokenizer()
    sot, eot = _tokenizer.encoder["<|startoftext|>"], _tokenizer.encoder["<|endoftext|>"]
    all_tokens = [[sot] + _tokenizer.encode(t) + [eot] for t in texts]
    result = torch.zeros(len(all_tokens), context_length, dtype=torch.int)
    for i, tokens in enumerate(all_tokens):
        if len(tokens) > context_length:
            if truncate:
                tokens = tokens[:context_length]
                tokens[-1] = eot
            else:
                raise RuntimeError(f"Input {texts[i]} is too long for context length {context_length}")
        result[i, :len(tokens)] = torch.tensor(tokens)
    return result
