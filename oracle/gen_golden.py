"""Generates the golden fixtures under tests/golden/ from HuggingFace ``transformers.CLIPModel``
-- an implementation of the published model that shares no code with ``oracle/clip_oracle.py``.

Run from the repository root:  python -m oracle.gen_golden

TEST INFRASTRUCTURE.  The reference holds no golden vectors of its own and its
arithmetic lives in the absent third-party ``clip`` package, so the oracle is pinned against the
one independent implementation that can be imported offline: every OUTPUT stored in a fixture
(features, logits, loss, gradients) is computed by ``CLIPModel`` (transformers, modeling_clip.py);
``clip_oracle`` only supplies the INPUTS -- the seeded random weights (upstream's
``initialize_parameters`` scheme, seed 567 = CLIP/train.py:28, rounded to bf16, mapped onto HF's
parameter names by ``to_hf_state_dict``) and the seeded synthetic images / tokens.  Weights and
inputs are not stored; tests regenerate them from the seed.  ``tests/test_cpu.py`` then checks the
oracle against these vectors, and the GPU tests check the CUDA path against both.
The fixtures record the transformers version they were generated with.
"""
from __future__ import annotations

import os
import warnings

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


def hf_model(name: str, jitter: float):
    """HF CLIPModel carrying the seeded, bf16-rounded weights of ``O.build(name)``."""
    cfg = O.CONFIGS[name]
    src = bf16_round_(O.build(name, seed=SEED, jitter=jitter))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hf = O.build_hf(cfg)
    res = hf.load_state_dict(O.to_hf_state_dict(src.state_dict(), cfg), strict=False)
    assert not res.unexpected_keys and all("position_ids" in k for k in res.missing_keys), res
    return hf, cfg


def hf_grads_upstream_names(hf, cfg) -> dict:
    """Gradients of the HF parameters under upstream's state-dict names (inverse of to_hf_state_dict)."""
    g = {n: p.grad for n, p in hf.named_parameters()}
    out = {
        "logit_scale": g["logit_scale"],
        "visual.conv1.weight": g["vision_model.embeddings.patch_embedding.weight"],
        "visual.class_embedding": g["vision_model.embeddings.class_embedding"],
        "visual.positional_embedding": g["vision_model.embeddings.position_embedding.weight"],
        "visual.ln_pre.weight": g["vision_model.pre_layrnorm.weight"],
        "visual.ln_pre.bias": g["vision_model.pre_layrnorm.bias"],
        "visual.ln_post.weight": g["vision_model.post_layernorm.weight"],
        "visual.ln_post.bias": g["vision_model.post_layernorm.bias"],
        "visual.proj": g["visual_projection.weight"].t(),
        "token_embedding.weight": g["text_model.embeddings.token_embedding.weight"],
        "positional_embedding": g["text_model.embeddings.position_embedding.weight"],
        "ln_final.weight": g["text_model.final_layer_norm.weight"],
        "ln_final.bias": g["text_model.final_layer_norm.bias"],
        "text_projection": g["text_projection.weight"].t(),
    }
    for dst, src, layers in (("visual.transformer", "vision_model", cfg.vision_layers),
                             ("transformer", "text_model", cfg.transformer_layers)):
        for i in range(layers):
            d, s = f"{dst}.resblocks.{i}.", f"{src}.encoder.layers.{i}."
            out[d + "attn.in_proj_weight"] = torch.cat([g[s + f"self_attn.{n}.weight"] for n in ("q_proj", "k_proj", "v_proj")])
            out[d + "attn.in_proj_bias"] = torch.cat([g[s + f"self_attn.{n}.bias"] for n in ("q_proj", "k_proj", "v_proj")])
            for a, b in (("attn.out_proj", "self_attn.out_proj"), ("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"),
                         ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
                out[d + a + ".weight"] = g[s + b + ".weight"]
                out[d + a + ".bias"] = g[s + b + ".bias"]
    return out


def make_case(name: str, n_img: int, n_txt: int, jitter: float, grads: bool, max_len: int = 12):
    import transformers
    hf, cfg = hf_model(name, jitter)
    img = O.synth_images(n_img, cfg.image_resolution, seed=SEED)
    tok = O.synth_tokens(n_txt, seed=SEED, min_len=3, max_len=max_len)
    out = {"generator": np.array(f"transformers.CLIPModel {transformers.__version__}")}
    with torch.no_grad():
        fi = hf.get_image_features(pixel_values=img).pooler_output   # un-normalised, like upstream encode_image
        ft = hf.get_text_features(input_ids=tok).pooler_output
        res = hf(input_ids=tok, pixel_values=img)
    out["image_features"] = fi.numpy()
    out["text_features"] = ft.numpy()
    out["logits_per_image"] = res.logits_per_image.numpy()
    out["tokens"] = tok.numpy()
    out["image_checksum"] = np.array([img.double().sum().item(), img.double().abs().sum().item()])
    if grads:
        assert n_img == n_txt
        hf.train()  # no dropout in CLIP; for symmetry with the fine-tune script
        hf.zero_grad()
        res = hf(input_ids=tok, pixel_values=img, return_loss=True)  # transformers' clip_loss == CLIP/train.py:162-166
        res.loss.backward()
        out["loss"] = np.array(res.loss.item())
        g = hf_grads_upstream_names(hf, cfg)
        out["grad_logit_scale"] = np.array(g["logit_scale"].item())
        order = [n for n, _ in O.build(name, seed=SEED).named_parameters()]
        assert set(order) == set(g), set(order) ^ set(g)
        out["grad_names"] = np.array(order)
        out["grad_norms"] = np.array([g[n].double().norm().item() for n in order])
        # a few full gradients (small tensors) for direction checks
        for n in ("visual.ln_post.weight", "ln_final.bias", "visual.class_embedding",
                  "transformer.resblocks.0.attn.in_proj_bias", "visual.transformer.resblocks.0.mlp.c_proj.bias",
                  "visual.proj", "text_projection"):
            if g[n].numel() <= 1 << 14:   # keep the fixtures small
                out["grad::" + n] = g[n].contiguous().numpy()
    return out


CASES = {
    # file name: (model, images, texts, jitter, gradients, longest caption)
    "tiny_fwd_6x4": ("tiny", 6, 4, 0.05, False, 12),
    "tiny_train_8": ("tiny", 8, 8, 0.05, True, 12),
    "vitb32_fwd_4x3": ("ViT-B/32", 4, 3, 0.05, False, 12),
    "vitb32_train_8": ("ViT-B/32", 8, 8, 0.05, True, 12),
    "vitb32_fwd_32x16": ("ViT-B/32", 32, 16, 0.05, False, 12),     # BASELINE config 1 (CLIP/predict.py shapes)
    "tiny_train_64_ragged": ("tiny", 64, 64, 0.05, True, 76),      # caption lengths U{3..76}: the packed text tower
}


def main():
    os.makedirs(OUT, exist_ok=True)
    for fname, (name, ni, nt, jit, grads, max_len) in CASES.items():
        data = make_case(name, ni, nt, jit, grads, max_len)
        path = os.path.join(OUT, fname + ".npz")
        np.savez_compressed(path, **data)
        print(fname, {k: v.shape for k, v in data.items() if hasattr(v, "shape")}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
