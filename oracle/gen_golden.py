"""Generates the golden fixtures under tests/golden/ from the CPU oracle (oracle/clip_oracle.py).

Run from the repository root:  python -m oracle.gen_golden
The reference holds no golden vectors of its own, so these pin the oracle's
behaviour (guarding it against drift) and give the GPU tests fixed targets that do not require
the oracle to run at full size.  Weights are NOT stored: they are regenerated from the seed by
``clip_oracle.build`` (deterministic CPU RNG) and rounded to bf16, the precision both sides
share.  Inputs are regenerated from the seed as well.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import clip_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 567  # CLIP/train.py:28


def bf16_round_(model):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n != "logit_scale":
                p.copy_(p.to(torch.bfloat16).float())
    return model


def make_case(name: str, n_img: int, n_txt: int, jitter: float, grads: bool):
    torch.manual_seed(SEED)
    cfg = O.CONFIGS[name]
    model = bf16_round_(O.build(name, seed=SEED, jitter=jitter))
    img = O.synth_images(n_img, cfg.image_resolution, seed=SEED)
    tok = O.synth_tokens(n_txt, seed=SEED, min_len=3, max_len=12 if n_txt <= 16 else 76)
    out = {}
    with torch.no_grad():
        fi = model.encode_image(img)
        ft = model.encode_text(tok)
        lpi, lpt = model(img, tok)
    out["image_features"] = fi.numpy()
    out["text_features"] = ft.numpy()
    out["logits_per_image"] = lpi.numpy()
    out["tokens"] = tok.numpy()
    out["image_checksum"] = np.array([img.double().sum().item(), img.double().abs().sum().item()])
    if grads:
        assert n_img == n_txt
        model.zero_grad()
        lpi, lpt = model(img, tok)
        loss = O.clip_loss(lpi, lpt)
        loss.backward()
        out["loss"] = np.array(loss.item())
        out["grad_logit_scale"] = np.array(model.logit_scale.grad.item())
        names, norms = [], []
        for n, p in model.named_parameters():
            names.append(n)
            norms.append(p.grad.double().norm().item())
        out["grad_names"] = np.array(names)
        out["grad_norms"] = np.array(norms)
        # a few full gradients (small tensors) for direction checks
        for n in ("visual.ln_post.weight", "ln_final.bias", "visual.class_embedding",
                  "transformer.resblocks.0.attn.in_proj_bias", "visual.transformer.resblocks.0.mlp.c_proj.bias"):
            out["grad::" + n] = dict(model.named_parameters())[n].grad.numpy()
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    cases = {
        "tiny_fwd_6x4": ("tiny", 6, 4, 0.05, False),
        "tiny_train_8": ("tiny", 8, 8, 0.05, True),
        "vitb32_fwd_4x3": ("ViT-B/32", 4, 3, 0.05, False),
        "vitb32_train_8": ("ViT-B/32", 8, 8, 0.05, True),
    }
    for fname, (name, ni, nt, jit, grads) in cases.items():
        data = make_case(name, ni, nt, jit, grads)
        path = os.path.join(OUT, fname + ".npz")
        np.savez_compressed(path, **data)
        print(fname, {k: v.shape for k, v in data.items() if hasattr(v, "shape")}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
