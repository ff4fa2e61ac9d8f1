"""CPU oracle for the CLIP dual-encoder hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product path
(``construction_clip_b200`` / ``clip``) never does.

What it restates
----------------
The reference (zhuluntsai/Construction-CLIP) does not contain the arithmetic of its
own hot path: every script does ``import clip`` (``CLIP/predict.py:2``,
``CLIP/train.py:8``, ``CLIP_prefix_caption/parse_coco.py:3``) and calls
``clip.load`` (``CLIP/predict.py:12``), ``model(image, text)`` (``CLIP/predict.py:46``,
``CLIP/train.py:161``) and ``model.encode_image`` (``parse_coco.py:43``).  ``clip`` is
the un-vendored, un-pinned third-party package openai/CLIP (setup.py version "1.0",
installed with ``pip install git+https://github.com/openai/CLIP.git``); it is absent
from ``/root/reference`` and from this image.  This file therefore restates the
published algorithm of upstream ``clip/model.py`` (ViT variants) in plain fp32
PyTorch, with upstream's module / parameter names, and anchors on the reference's
call sites:

* ``model(image, text) -> (logits_per_image, logits_per_text)``  CLIP/train.py:161
* symmetric InfoNCE  ``(CE(lpi, arange) + CE(lpt, arange)) / 2``     CLIP/train.py:162-166
* zero-shot rule ``argmax(softmax(logits_per_image))``               CLIP/predict.py:47,54
* un-normalised ``encode_image`` output used as the prefix           parse_coco.py:43

PARITY PINNING: the reference holds no tests, golden vectors or fixtures for this
path ("parity unpinned" by the reference itself, and upstream ``clip`` cannot be imported
here).  The restatement is pinned instead against the one independent implementation
of the same published model that is importable offline, HuggingFace ``transformers``
``CLIPModel``: ``oracle/gen_golden.py`` computes every stored output (features, logits,
loss, the gradient norm of all parameter tensors and a few full gradients) with
``CLIPModel`` -- this module only supplies the seeded weights and inputs -- and
``tests/test_cpu.py`` checks this oracle against those fixtures and against HF live.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


@dataclass(frozen=True)
class ClipConfig:
    """Hyper-parameters upstream ``build_model`` infers from state-dict shapes."""
    name: str
    embed_dim: int
    image_resolution: int
    vision_layers: int
    vision_width: int
    vision_patch_size: int
    context_length: int = 77
    vocab_size: int = 49408
    transformer_width: int = 512
    transformer_heads: int = 8
    transformer_layers: int = 12

    @property
    def vision_heads(self) -> int:
        return self.vision_width // 64

    @property
    def grid(self) -> int:
        return self.image_resolution // self.vision_patch_size

    @property
    def vision_tokens(self) -> int:
        return self.grid * self.grid + 1


CONFIGS = {
    "ViT-B/32": ClipConfig("ViT-B/32", 512, 224, 12, 768, 32, 77, 49408, 512, 8, 12),
    "ViT-B/16": ClipConfig("ViT-B/16", 512, 224, 12, 768, 16, 77, 49408, 512, 8, 12),
    "ViT-L/14": ClipConfig("ViT-L/14", 768, 224, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "ViT-L/14@336px": ClipConfig("ViT-L/14@336px", 768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12),
    # a tiny configuration for fast CPU tests (same code path, head_dim 64)
    "tiny": ClipConfig("tiny", 64, 64, 2, 128, 32, 77, 49408, 128, 2, 2),
}


class LayerNorm(nn.LayerNorm):
    """Upstream: LayerNorm computed in fp32, cast back to the input dtype."""

    def forward(self, x: torch.Tensor):
        orig_type = x.dtype
        ret = super().forward(x.type(torch.float32))
        return ret.type(orig_type)


class QuickGELU(nn.Module):
    def forward(self, x: torch.Tensor):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor | None = None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def attention(self, x: torch.Tensor):
        mask = self.attn_mask.to(dtype=x.dtype, device=x.device) if self.attn_mask is not None else None
        return self.attn(x, x, x, need_weights=False, attn_mask=mask)[0]

    def forward(self, x: torch.Tensor):
        x = x + self.attention(self.ln_1(x))
        x = x + self.mlp(self.ln_2(x))
        return x


class Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int, attn_mask: torch.Tensor | None = None):
        super().__init__()
        self.width = width
        self.layers = layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x: torch.Tensor):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution: int, patch_size: int, width: int, layers: int, heads: int, output_dim: int):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x: torch.Tensor):
        x = self.conv1(x)                                   # [B, width, g, g]
        x = x.reshape(x.shape[0], x.shape[1], -1)           # [B, width, g*g]
        x = x.permute(0, 2, 1)                              # [B, g*g, width]
        cls = self.class_embedding.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype, device=x.device)
        x = torch.cat([cls, x], dim=1)                      # [B, g*g+1, width]
        x = x + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = x.permute(1, 0, 2)                              # NLD -> LND
        x = self.transformer(x)
        x = x.permute(1, 0, 2)                              # LND -> NLD
        x = self.ln_post(x[:, 0, :])
        if self.proj is not None:
            x = x @ self.proj
        return x


class CLIP(nn.Module):
    """Restated upstream ``clip.model.CLIP`` (ViT image tower only)."""

    def __init__(self, cfg: ClipConfig):
        super().__init__()
        self.cfg = cfg
        self.context_length = cfg.context_length
        self.visual = VisionTransformer(cfg.image_resolution, cfg.vision_patch_size, cfg.vision_width,
                                        cfg.vision_layers, cfg.vision_heads, cfg.embed_dim)
        self.transformer = Transformer(cfg.transformer_width, cfg.transformer_layers, cfg.transformer_heads,
                                       attn_mask=self.build_attention_mask())
        self.vocab_size = cfg.vocab_size
        self.token_embedding = nn.Embedding(cfg.vocab_size, cfg.transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(cfg.context_length, cfg.transformer_width))
        self.ln_final = LayerNorm(cfg.transformer_width)
        self.text_projection = nn.Parameter(torch.empty(cfg.transformer_width, cfg.embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.initialize_parameters()

    def initialize_parameters(self):
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (self.transformer.width ** -0.5) * ((2 * self.transformer.layers) ** -0.5)
        attn_std = self.transformer.width ** -0.5
        fc_std = (2 * self.transformer.width) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=self.transformer.width ** -0.5)

    def build_attention_mask(self):
        # causal; padding positions are NOT masked (upstream behaviour)
        mask = torch.empty(self.context_length, self.context_length)
        mask.fill_(float("-inf"))
        mask.triu_(1)
        return mask

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))

    def encode_text(self, text):
        x = self.token_embedding(text).type(self.dtype)     # [B, n_ctx, d]
        x = x + self.positional_embedding.type(self.dtype)
        x = x.permute(1, 0, 2)
        x = self.transformer(x)
        x = x.permute(1, 0, 2)
        x = self.ln_final(x).type(self.dtype)
        # EOT pooling: the EOT id (49407) is the largest id in each row
        x = x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection
        return x

    def forward(self, image, text):
        image_features = self.encode_image(image)
        text_features = self.encode_text(text)
        image_features = image_features / image_features.norm(dim=1, keepdim=True)
        text_features = text_features / text_features.norm(dim=1, keepdim=True)
        logit_scale = self.logit_scale.exp()
        logits_per_image = logit_scale * image_features @ text_features.t()
        logits_per_text = logits_per_image.t()
        return logits_per_image, logits_per_text


def build(name: str, seed: int = 567, jitter: float = 0.0) -> CLIP:
    """Random-init oracle model (upstream ``initialize_parameters`` scheme, seed per
    CLIP/train.py:28).  ``jitter`` > 0 perturbs LayerNorm affine parameters and the
    biases so that every parameter tensor is exercised by parity tests."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    model = CLIP(CONFIGS[name]).float().eval()
    if jitter > 0:
        with torch.no_grad():
            for n, p in model.named_parameters():
                if n.endswith("ln_1.weight") or n.endswith("ln_2.weight") or "ln_pre.weight" in n \
                        or "ln_post.weight" in n or "ln_final.weight" in n:
                    p.add_(jitter * torch.randn_like(p))
                elif n.endswith("bias"):
                    p.add_(jitter * torch.randn_like(p))
    torch.random.set_rng_state(g)
    return model


# ----------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------
SOT, EOT = 49406, 49407


def synth_images(batch: int, resolution: int, seed: int = 567) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, resolution, resolution, generator=g)


def synth_tokens(batch: int, seed: int = 567, context_length: int = 77,
                 min_len: int = 3, max_len: int = 76) -> torch.Tensor:
    """[B,77] ids: SOT, random body, one EOT (unique arg-max), zero padding."""
    g = torch.Generator().manual_seed(seed + 1)
    ids = torch.zeros(batch, context_length, dtype=torch.int64)
    lens = torch.randint(min_len, max_len + 1, (batch,), generator=g)
    body = torch.randint(1, SOT, (batch, context_length), generator=g)
    for b in range(batch):
        L = int(lens[b])
        ids[b, 0] = SOT
        ids[b, 1:L] = body[b, 1:L]
        ids[b, L] = EOT
    return ids


def clip_loss(logits_per_image: torch.Tensor, logits_per_text: torch.Tensor):
    """Symmetric InfoNCE exactly as CLIP/train.py:162-166 (labels = arange, mean CE, /2)."""
    label = torch.arange(logits_per_image.shape[0], device=logits_per_image.device)
    loss_i = F.cross_entropy(logits_per_image.float(), label)
    loss_t = F.cross_entropy(logits_per_text.float(), label)
    return (loss_i + loss_t) / 2


# ----------------------------------------------------------------------------------------
# algorithmic FLOP model -- used by bench.py for the roofline
# ----------------------------------------------------------------------------------------
def flops_image(cfg: ClipConfig) -> float:
    n, d, L = cfg.vision_tokens, cfg.vision_width, cfg.vision_layers
    per_layer = 24 * n * d * d + 4 * n * n * d
    patch = 2 * cfg.grid ** 2 * 3 * cfg.vision_patch_size ** 2 * d
    return L * per_layer + patch + 2 * d * cfg.embed_dim


def flops_text(cfg: ClipConfig) -> float:
    n, d, L = cfg.context_length, cfg.transformer_width, cfg.transformer_layers
    per_layer = 24 * n * d * d + 4 * (n * (n + 1) // 2) * d
    return L * per_layer + 2 * d * cfg.embed_dim


def flops_pair(cfg: ClipConfig) -> float:
    return flops_image(cfg) + flops_text(cfg)


# ----------------------------------------------------------------------------------------
# HuggingFace cross-check helpers (independent second implementation)
# ----------------------------------------------------------------------------------------
def to_hf_state_dict(sd: dict, cfg: ClipConfig) -> dict:
    """Map an upstream-named state dict onto transformers.CLIPModel names."""
    out = {}
    def put(k, v):
        out[k] = v.clone()
    put("logit_scale", sd["logit_scale"])
    put("vision_model.embeddings.patch_embedding.weight", sd["visual.conv1.weight"])
    put("vision_model.embeddings.class_embedding", sd["visual.class_embedding"])
    put("vision_model.embeddings.position_embedding.weight", sd["visual.positional_embedding"])
    put("vision_model.pre_layrnorm.weight", sd["visual.ln_pre.weight"])
    put("vision_model.pre_layrnorm.bias", sd["visual.ln_pre.bias"])
    put("vision_model.post_layernorm.weight", sd["visual.ln_post.weight"])
    put("vision_model.post_layernorm.bias", sd["visual.ln_post.bias"])
    put("visual_projection.weight", sd["visual.proj"].t())
    put("text_model.embeddings.token_embedding.weight", sd["token_embedding.weight"])
    put("text_model.embeddings.position_embedding.weight", sd["positional_embedding"])
    put("text_model.final_layer_norm.weight", sd["ln_final.weight"])
    put("text_model.final_layer_norm.bias", sd["ln_final.bias"])
    put("text_projection.weight", sd["text_projection"].t())
    for src, dst, layers, w in (("visual.transformer", "vision_model", cfg.vision_layers, cfg.vision_width),
                                ("transformer", "text_model", cfg.transformer_layers, cfg.transformer_width)):
        for i in range(layers):
            s = f"{src}.resblocks.{i}."
            d = f"{dst}.encoder.layers.{i}."
            ipw, ipb = sd[s + "attn.in_proj_weight"], sd[s + "attn.in_proj_bias"]
            for j, nm in enumerate(("q_proj", "k_proj", "v_proj")):
                put(d + f"self_attn.{nm}.weight", ipw[j * w:(j + 1) * w])
                put(d + f"self_attn.{nm}.bias", ipb[j * w:(j + 1) * w])
            put(d + "self_attn.out_proj.weight", sd[s + "attn.out_proj.weight"])
            put(d + "self_attn.out_proj.bias", sd[s + "attn.out_proj.bias"])
            for a, b in (("ln_1", "layer_norm1"), ("ln_2", "layer_norm2"),
                         ("mlp.c_fc", "mlp.fc1"), ("mlp.c_proj", "mlp.fc2")):
                put(d + b + ".weight", sd[s + a + ".weight"])
                put(d + b + ".bias", sd[s + a + ".bias"])
    return out


def build_hf(cfg: ClipConfig):
    from transformers import CLIPConfig, CLIPModel
    hf_cfg = CLIPConfig(
        text_config=dict(hidden_size=cfg.transformer_width, intermediate_size=4 * cfg.transformer_width,
                         num_hidden_layers=cfg.transformer_layers, num_attention_heads=cfg.transformer_heads,
                         max_position_embeddings=cfg.context_length, vocab_size=cfg.vocab_size,
                         eos_token_id=EOT, bos_token_id=SOT, pad_token_id=0, hidden_act="quick_gelu",
                         projection_dim=cfg.embed_dim),
        vision_config=dict(hidden_size=cfg.vision_width, intermediate_size=4 * cfg.vision_width,
                           num_hidden_layers=cfg.vision_layers, num_attention_heads=cfg.vision_heads,
                           image_size=cfg.image_resolution, patch_size=cfg.vision_patch_size,
                           hidden_act="quick_gelu", projection_dim=cfg.embed_dim),
        projection_dim=cfg.embed_dim,
    )
    return CLIPModel(hf_cfg).float().eval()
