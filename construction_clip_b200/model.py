"""``CLIP`` nn.Module with upstream openai/CLIP's attribute names, state-dict keys and call
surface, whose arithmetic runs entirely in libb200clip (hand-written sm_100a kernels).

Reference call sites this mirrors:
  clip.load(...)->model, model.load_state_dict(...)            CLIP/predict.py:12-16, CLIP/train.py:105-111
  model(image, text) -> (logits_per_image, logits_per_text)    CLIP/predict.py:46, CLIP/train.py:161
  model.encode_image(image)                                    CLIP_prefix_caption/parse_coco.py:43
  loss.backward() through the model                            CLIP/train.py:168
  AdamW(model.parameters())                                    CLIP/train.py:143
The nn.Module tree (visual.conv1, visual.transformer.resblocks.N.attn.in_proj_weight, mlp.c_fc,
ln_1, ...) only HOLDS parameters so that state_dict()/load_state_dict()/parameters() behave like
upstream; no torch operator of those modules is ever called.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass

import torch
from torch import nn

from . import lib as L
from . import ops as O
from . import towers as T

bf16, f32 = torch.bfloat16, torch.float32


@dataclass(frozen=True)
class ClipConfig:
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
    def vision_heads(self):
        return self.vision_width // 64

    @property
    def grid(self):
        return self.image_resolution // self.vision_patch_size

    @property
    def vision_tokens(self):
        return self.grid ** 2 + 1


CONFIGS = {
    "ViT-B/32": ClipConfig("ViT-B/32", 512, 224, 12, 768, 32, 77, 49408, 512, 8, 12),
    "ViT-B/16": ClipConfig("ViT-B/16", 512, 224, 12, 768, 16, 77, 49408, 512, 8, 12),
    "ViT-L/14": ClipConfig("ViT-L/14", 768, 224, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "ViT-L/14@336px": ClipConfig("ViT-L/14@336px", 768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12),
    "tiny": ClipConfig("tiny", 64, 64, 2, 128, 32, 77, 49408, 128, 2, 2),
}


def config_from_state_dict(sd: dict) -> ClipConfig:
    """Hyper-parameters from tensor shapes, like upstream ``build_model``."""
    if "visual.proj" not in sd:
        raise RuntimeError("only ViT checkpoints are supported (BASELINE config 4: 'RN-free')")
    vision_width = sd["visual.conv1.weight"].shape[0]
    vision_layers = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    patch = sd["visual.conv1.weight"].shape[-1]
    grid = round((sd["visual.positional_embedding"].shape[0] - 1) ** 0.5)
    width = sd["ln_final.weight"].shape[0]
    layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks")})
    return ClipConfig("custom", sd["text_projection"].shape[1], patch * grid, vision_layers, vision_width, patch,
                      sd["positional_embedding"].shape[0], sd["token_embedding.weight"].shape[0], width, width // 64,
                      layers)


# ------------------------------------------------------------------------------------------------
# parameter containers (names == upstream)
class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class QuickGELU(nn.Module):
    def forward(self, x):  # placeholder: the activation is fused into the c_fc GEMM epilogue
        raise RuntimeError("QuickGELU is fused into the mlp.c_fc GEMM epilogue; call the CLIP model instead")


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.attn = _Attn(d)
        self.ln_1 = nn.LayerNorm(d)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d, 4 * d)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(4 * d, d))]))
        self.ln_2 = nn.LayerNorm(d)


class Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers, self.heads = width, layers, heads
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width) for _ in range(layers)])


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.input_resolution = input_resolution
        self.output_dim = output_dim
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))


# ------------------------------------------------------------------------------------------------
class ParamStore:
    """Flat bf16 shadow of one tower's parameters (what the kernels read) + layout of the flat
    fp32 gradient buffer.  bf16 parameters on the right device are re-pointed at views of the
    shadow (zero-copy); anything else (fp32 / fp16 parameters after ``model.float()``) is copied in
    whenever its version counter moves."""

    ALIGN = 64  # elements

    def __init__(self, named_params, device, special_shapes=None):
        self.device = device
        self.entries = []  # (name, param, offset, store_shape)
        off = 0
        special_shapes = special_shapes or {}
        for name, p in named_params:
            shape = tuple(special_shapes.get(name, p.shape))
            n = math.prod(shape)
            self.entries.append((name, p, off, shape))
            off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        # pad the flat length so that it splits evenly (in 64-element units) over 1..8 ranks: the
        # trainer reduce-scatters gradients / all-gathers updated weights in equal shards
        unit = self.ALIGN * 840
        self.total = (off + unit - 1) // unit * unit
        self.w = torch.zeros(self.total, device=device, dtype=bf16)
        self.W = {name: self.w[o:o + math.prod(s)].view(s) for name, _, o, s in self.entries}
        self._seen = {}
        self.link()

    def grad_views(self, flat):
        return {name: flat[o:o + math.prod(s)].view(s) for name, _, o, s in self.entries}

    def _copy_in(self, name, p):
        dst = self.W[name]
        if dst.shape == p.shape:
            dst.copy_(p.detach())
        else:  # zero-padded 2-D view of conv1.weight
            dst.zero_()
            dst[:, :p[0].numel()].copy_(p.detach().reshape(p.shape[0], -1))

    def _copy_in_many(self, items):
        """Refresh many shadow entries with ONE multi-tensor copy (an optimizer.step() on fp32 parameters
        touches all ~150 tensors of a tower; one launch per tensor would cost more than the forward's LayerNorms)."""
        same = [(self.W[n], p.detach()) for n, p in items if self.W[n].shape == p.shape]
        if len(same) > 1 and hasattr(torch, "_foreach_copy_"):
            torch._foreach_copy_([d for d, _ in same], [q for _, q in same])
        else:
            for d, q in same:
                d.copy_(q)
        for n, p in items:
            if self.W[n].shape != p.shape:
                self._copy_in(n, p)

    def link(self):
        """(Re-)point bf16 parameters at the shadow so kernels and optimiser share storage."""
        with torch.no_grad():
            self._copy_in_many([(name, p) for name, p, _, _ in self.entries])
            for name, p, o, s in self.entries:
                if p.dtype == bf16 and p.device == self.w.device and math.prod(s) == p.numel():
                    p.data = self.W[name].view(p.shape)
                self._seen[name] = (p.data_ptr(), p._version)

    def sync(self):
        """Refresh shadow entries whose parameter is not a view of the shadow and has changed."""
        stale = []
        for name, p, o, s in self.entries:
            ptr = p.data_ptr()
            if ptr == self.W[name].data_ptr() and p.dtype == bf16:
                continue
            if self._seen.get(name) != (ptr, p._version):
                stale.append((name, p))
                self._seen[name] = (ptr, p._version)
        if stale:
            with torch.no_grad():
                self._copy_in_many(stale)

    def unlinked(self):
        """Entries whose nn.Parameter is NOT a zero-copy view of the shadow (fp32 / fp16 parameters, the
        zero-padded conv1.weight of ViT-L/14): a trainer that updates the shadow must write them back."""
        return [(name, p, o, s) for name, p, o, s in self.entries
                if not (p.data_ptr() == self.W[name].data_ptr() and p.dtype == bf16)]

    def master_f32(self):
        """Flat fp32 copy of the tower's weights for an optimiser: taken from the nn.Parameters themselves
        when they carry more precision than the bf16 shadow (fp32 parameters, what clip.load returns)."""
        out = self.w.float()
        with torch.no_grad():
            for name, p, o, s in self.entries:
                if p.dtype == f32 and p.device == self.w.device:
                    dst = out[o:o + math.prod(s)].view(s)
                    if tuple(s) == tuple(p.shape):
                        dst.copy_(p.detach())
                    else:
                        dst[:, :p[0].numel()].copy_(p.detach().reshape(p.shape[0], -1))
        return out


# ------------------------------------------------------------------------------------------------
class _TowerFn(torch.autograd.Function):
    """One tower = one autograd node: forward launches the tower's kernels and keeps the
    activations; backward launches the hand-written backward kernels and returns one gradient per
    parameter (fp32 accumulation, cast once to the parameter dtype)."""

    @staticmethod
    def forward(ctx, model, which, need_grad, inp, *params):
        store = model._store(which, inp.device)
        store.sync()
        fwd = T.vision_fwd if which == "visual" else T.text_fwd
        feat, saved = fwd(store.W, model.cfg, inp, need_grad)
        ctx.model, ctx.which, ctx.saved, ctx.store = model, which, saved, store
        ctx.param_meta = [(p.dtype, p.shape) for p in params]
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        store, model = ctx.store, ctx.model
        if ctx.saved is None:
            raise RuntimeError("backward through a CLIP tower that ran without saved activations")
        flat = torch.zeros(store.total, device=store.device, dtype=f32)
        G = store.grad_views(flat)
        bwd = T.vision_bwd if ctx.which == "visual" else T.text_bwd
        bwd(store.W, G, model.cfg, ctx.saved, dfeat.to(f32))
        ctx.saved = None
        flat_bf = None
        grads = []
        for (name, p, o, s), (dt, shape) in zip(store.entries, ctx.param_meta):
            n = math.prod(s)
            if dt == f32:
                g = flat[o:o + n].view(s)
            else:
                if flat_bf is None:
                    flat_bf = O.cast_f32_to_bf16(flat)
                g = flat_bf[o:o + n].view(s)
                if dt != bf16:
                    g = g.to(dt)
            if tuple(s) != tuple(shape):  # padded conv1.weight
                g = g[:, :math.prod(shape[1:])].reshape(shape)
            grads.append(g)
        return (None, None, None, None, *grads)


class _LogitsFn(torch.autograd.Function):
    """logit_scale.exp() * normalize(I) @ normalize(T).t()   (materialised; the drop-in path)."""

    @staticmethod
    def forward(ctx, img_f, txt_f, logit_scale):
        img_n, inv_i = O.l2norm_fwd(img_f.contiguous())
        txt_n, inv_t = O.l2norm_fwd(txt_f.contiguous())
        ls = logit_scale.detach().to(f32).reshape(1).contiguous()
        logits = O.logits(img_n, txt_n, ls)
        ctx.save_for_backward(img_n, inv_i, txt_n, inv_t, ls, logits)
        ctx.ls_dtype = logit_scale.dtype
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        img_n, inv_i, txt_n, inv_t, ls, logits = ctx.saved_tensors
        Bi, Bt = logits.shape
        E = img_n.shape[1]
        dev = logits.device
        dlogits = dlogits.to(f32)
        # operands of the tensor-core GEMMs need 16-byte row pitches: pad the batch extents to 8
        Bip, Btp = (Bi + 7) // 8 * 8, (Bt + 7) // 8 * 8
        dl = torch.zeros((Bip, Btp), device=dev, dtype=bf16)
        dl[:Bi, :Bt] = dlogits
        img_b = torch.zeros((Bip, E), device=dev, dtype=bf16)
        img_b[:Bi] = img_n
        txt_b = torch.zeros((Btp, E), device=dev, dtype=bf16)
        txt_b[:Bt] = txt_n
        s = ls.exp()
        d_img_n = torch.zeros((Bip, E), device=dev, dtype=f32)
        d_txt_n = torch.zeros((Btp, E), device=dev, dtype=f32)
        O.gemm(dl, txt_b, b_major=L.MAJOR_MN, out=d_img_n, scale=s, split_k=0, accumulate=True)
        O.gemm(dl, img_b, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, out=d_txt_n, scale=s, split_k=0, accumulate=True)
        d_img = O.l2norm_bwd(d_img_n[:Bi].contiguous(), img_n, inv_i).float()
        d_txt = O.l2norm_bwd(d_txt_n[:Bt].contiguous(), txt_n, inv_t).float()
        d_ls = (dlogits * logits).sum().to(ctx.ls_dtype)  # d logits / d logit_scale = logits
        return d_img, d_txt, d_ls


# ------------------------------------------------------------------------------------------------
class CLIP(nn.Module):
    def __init__(self, cfg: ClipConfig):
        super().__init__()
        self.cfg = cfg
        self.context_length = cfg.context_length
        self.vocab_size = cfg.vocab_size
        self.visual = VisionTransformer(cfg.image_resolution, cfg.vision_patch_size, cfg.vision_width,
                                        cfg.vision_layers, cfg.vision_heads, cfg.embed_dim)
        self.transformer = Transformer(cfg.transformer_width, cfg.transformer_layers, cfg.transformer_heads)
        self.token_embedding = nn.Embedding(cfg.vocab_size, cfg.transformer_width)
        self.positional_embedding = nn.Parameter(torch.empty(cfg.context_length, cfg.transformer_width))
        self.ln_final = nn.LayerNorm(cfg.transformer_width)
        self.text_projection = nn.Parameter(torch.empty(cfg.transformer_width, cfg.embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self._stores = {}
        self._trainer = None          # weakref to the ClipTrainer that owns the master weights, if any
        self._igraphs = {}            # captured no-grad forwards, one per input shape (see _graphed)
        self._side_stream = None
        self.fp32_check_mode = False  # see set_fp32_check_mode
        self.initialize_parameters()

    # upstream's scheme (clip.model.CLIP.initialize_parameters)
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

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    # -------------------------------------------------------------------------------- parameters
    def _tower_named_params(self, which):
        if which == "visual":
            return [(n, p) for n, p in self.visual.named_parameters()]
        return [(n, p) for n, p in self.named_parameters() if not n.startswith("visual.") and n != "logit_scale"]

    def _store(self, which, device) -> ParamStore:
        st = self._stores.get(which)
        if st is None or st.device != device:
            if device.type != "cuda":
                raise RuntimeError("the B200 CLIP path has no CPU fallback: move the model and inputs to a CUDA "
                                   "(sm_100) device")
            special = {}
            if which == "visual":
                k = 3 * self.cfg.vision_patch_size ** 2
                special["conv1.weight"] = (self.cfg.vision_width, (k + 63) // 64 * 64)
            st = ParamStore(self._tower_named_params(which), device, special)
            self._stores[which] = st
        return st

    def state_dict(self, *a, **k):
        """A ClipTrainer updates flat master weights + the bf16 shadow; parameters that are not views of the
        shadow (fp32 parameters, padded conv1.weight) are refreshed here so that the checkpoints of
        CLIP/train.py:210-216 (`torch.save(model.state_dict())`) always hold the trained weights.  With a
        sharded optimiser this gathers the master shards: call it on every rank."""
        tr = self._trainer() if getattr(self, "_trainer", None) is not None else None
        if tr is not None and tr.dirty:
            tr.write_back()
        return super().state_dict(*a, **k)

    def _apply(self, fn, *a, **k):  # .to() / .float() / .half() invalidate the zero-copy links
        out = super()._apply(fn, *a, **k)
        self._stores = {}
        self._igraphs = {}
        return out

    # ---------------------------------------------------------------------------------- forward
    def set_fp32_check_mode(self, on=True):
        """Route the (inference) forward through the tensor-core-free fp32 kernels: fp32 activations
        end to end, same bf16-representable weights.  Meets the 1e-4 logit tolerance of
        BASELINE.json's "fp32 check mode"; ~50x slower, forward only."""
        self.fp32_check_mode = bool(on)
        return self

    def _features(self, which, inp):
        params = [p for _, p in self._tower_named_params(which)]
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        if self.fp32_check_mode:
            if need_grad:
                raise RuntimeError("fp32 check mode is forward only: call under torch.no_grad() / model.eval()")
            store = self._store(which, inp.device)
            store.sync()
            fwd = T.vision_fwd_f32 if which == "visual" else T.text_fwd_f32
            return fwd(store.W, self.cfg, inp)
        return _TowerFn.apply(self, which, need_grad, inp, *params)

    # ------------------------------------------------------------------- launch-bound inference
    # The reference's inference calls are tiny (CLIP/predict.py:40-54: a few images x 2..16 prompts;
    # parse_coco.py:37-53: ONE image and three tower passes per iteration): ~180 kernel launches whose host
    # cost (Python + ctypes + tensor-map encoding, ~15 us each) is 5-10x their GPU time.  Under torch.no_grad()
    # such calls are captured ONCE per input shape into a CUDA graph and replayed: inputs are copied into the
    # graph's static buffers, outputs are handed back as copies.  Large batches (where the host is not the
    # limiter and the graph's private activation pool would be big) keep running eagerly.
    GRAPH_MAX_ROWS = int(__import__("os").environ.get("B200CLIP_INFER_GRAPH_MAX_ROWS", "16384"))

    def _graph_ok(self, *inputs):
        if torch.is_grad_enabled() or self.fp32_check_mode or self.GRAPH_MAX_ROWS <= 0:
            return False
        if not all(t.is_cuda for t in inputs) or torch.cuda.is_current_stream_capturing():
            return False
        rows = 0
        for t in inputs:
            rows += t.shape[0] * (self.cfg.vision_tokens if t.dim() == 4 else t.shape[1])
        return rows <= self.GRAPH_MAX_ROWS

    def _graphed(self, tag, fn, *inputs):
        key = (tag,) + tuple((tuple(t.shape), t.dtype, t.device) for t in inputs)
        for which in ("visual", "text"):     # refresh the weight shadow OUTSIDE the graph (fp32 parameters that moved)
            st = self._stores.get(which)
            if st is not None:
                st.sync()
        ent = self._igraphs.get(key)
        if ent is None:
            static_in = [t.detach().clone() for t in inputs]
            dev = inputs[0].device
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):        # warm-up: allocator, lazy stores, per-shape index caches
                fn(*static_in)
                fn(*static_in)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = fn(*static_in)
            if len(self._igraphs) >= 8:          # a handful of shapes per process; drop the oldest
                self._igraphs.pop(next(iter(self._igraphs)))
            ent = self._igraphs[key] = (graph, static_in, out)
        graph, static_in, out = ent
        for s_, t in zip(static_in, inputs):
            s_.copy_(t, non_blocking=True)
        graph.replay()
        return tuple(o.clone() for o in out) if isinstance(out, tuple) else out.clone()

    def encode_image(self, image):
        """[B,3,R,R] -> un-normalised [B, embed_dim] in the model dtype (parse_coco.py:43)."""
        if self._graph_ok(image):
            return self._graphed("img", lambda im: self._features("visual", im).to(self.dtype), image)
        return self._features("visual", image).to(self.dtype)

    def encode_text(self, text):
        if self._graph_ok(text):
            return self._graphed("txt", lambda tx: self._features("text", tx).to(self.dtype), text)
        return self._features("text", text).to(self.dtype)

    def _forward_eager(self, image, text):
        if torch.is_grad_enabled() or not image.is_cuda:
            img_f = self._features("visual", image)
            txt_f = self._features("text", text)
        else:
            # inference: the two towers are independent until the logits -> the text tower runs on a second
            # stream (two parallel branches of the captured graph; small calls are bound by the length of the
            # dependent kernel chain, not by SM count)
            cur = torch.cuda.current_stream(image.device)
            if self._side_stream is None or self._side_stream.device != image.device:
                self._side_stream = torch.cuda.Stream(device=image.device)
            side = self._side_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                txt_f = self._features("text", text)
            img_f = self._features("visual", image)
            cur.wait_stream(side)
            txt_f.record_stream(cur)
        return _LogitsFn.apply(img_f, txt_f, self.logit_scale)

    def forward(self, image, text):
        """-> (logits_per_image [Bi,Bt], logits_per_text [Bt,Bi]); fp32 logits (CLIP/train.py:161)."""
        if self._graph_ok(image, text):
            logits_per_image = self._graphed("fwd", self._forward_eager, image, text)
        else:
            logits_per_image = self._forward_eager(image, text)
        return logits_per_image, logits_per_image.t()


def build_model(state_dict: dict | None = None, name: str | None = None) -> CLIP:
    """Upstream ``build_model``: hyper-parameters from a state dict (or a named config)."""
    if state_dict is not None:
        sd = {k: v for k, v in state_dict.items() if k not in ("input_resolution", "context_length", "vocab_size")}
        cfg = config_from_state_dict(sd)
        model = CLIP(cfg)
        model.load_state_dict(sd)
        return model
    return CLIP(CONFIGS[name])
