"""`clip.model` namespace of upstream: CLIP, build_model, convert_weights."""
from construction_clip_b200.model import CLIP, CONFIGS, ClipConfig, build_model  # noqa: F401


def convert_weights(model):
    """Upstream casts applicable parameters to fp16; the B200 path computes in bf16 with fp32
    accumulation, so this casts to bf16 (logit_scale stays fp32).  bf16 parameters alias the
    kernels' weight buffer (zero copy) -- right for inference and for ClipTrainer (which keeps fp32
    master weights); do NOT run a plain optimiser with lr ~ 1e-5 on them, the updates round away."""
    import torch
    ls = model.logit_scale.data.float().clone()
    model.to(torch.bfloat16)
    model.logit_scale.data = ls
    return model
