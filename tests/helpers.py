"""Shared helpers of the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = 567  # CLIP/train.py:28


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def bf16_round_(model):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n != "logit_scale":
                p.copy_(p.to(torch.bfloat16).float())
    return model


def oracle_model(name, jitter=0.05):
    from oracle import clip_oracle as O
    return bf16_round_(O.build(name, seed=SEED, jitter=jitter))


def device_model(name, oracle, device="cuda", dtype=torch.bfloat16):
    """Our CLIP with exactly the oracle's (bf16-representable) weights."""
    from construction_clip_b200.model import CLIP, CONFIGS
    m = CLIP(CONFIGS[name])
    m.load_state_dict(oracle.state_dict(), strict=True)
    m = m.to(device)
    if dtype == torch.bfloat16:
        ls = m.logit_scale.data.float().clone()
        m = m.to(torch.bfloat16)
        m.logit_scale.data = ls
    return m


def cosine_rows(a, b):
    a, b = a.double(), b.double()
    return (a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-300))
