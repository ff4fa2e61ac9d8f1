"""Forward / backward of the two CLIP towers as sequences of C-ABI kernel launches.

Mirrors upstream ``clip/model.py`` (``VisionTransformer.forward``, ``CLIP.encode_text``,
``ResidualAttentionBlock.forward``) -- the code the reference reaches through
``model(image, text)`` (CLIP/train.py:161, CLIP/predict.py:46) and ``model.encode_image``
(CLIP_prefix_caption/parse_coco.py:43) -- but every operator is one hand-written sm_100a kernel
(see include/b200clip.h).  Activations are token-major ``[B*S, d]`` (upstream permutes to seq-first
LND; the maths is layout independent): the residual stream is fp32 (it accumulates 2 x layers
branch outputs -- rounding it to bf16 each time alone costs more than the 1e-2 logit tolerance),
every GEMM / attention operand is bf16, the backward gradient stream is bf16.

``W`` / ``G`` are dicts ``name -> tensor`` of bf16 weights / fp32 gradient accumulators that use
upstream's state-dict names relative to the tower prefix.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import torch

from . import lib as L
from . import ops as O

bf16, f32, i32 = torch.bfloat16, torch.float32, torch.int32

# Run the last block's out_proj / ln_2 / MLP on the pooled tokens only (see blocks_fwd).  The switch
# exists for A/B timing and for the test that proves both settings give the same features and gradients.
POOL_LAST_BLOCK = os.environ.get("B200CLIP_POOL_LAST_BLOCK", "1") != "0"

# Packed text tower (default; B200CLIP_PACK_TEXT=0 disables it).  Under upstream's causal mask nothing after a
# caption's EOT token can reach the pooled feature x[arange, text.argmax(-1)], and those positions receive
# exactly-zero gradients (tests/test_cpu.py::test_positions_after_eot_are_dead_in_the_oracle), so the text
# tower runs on sum(EOT position + 1) rows instead of B x 77 -- the same dead-code argument as POOL_LAST_BLOCK,
# applied to every block.  Same features, loss and gradients (tests/test_model_gpu.py::test_packed_text_*).
# Shapes stay STATIC: the caller passes the row count of the packed buffers (`rows`, >= the real count --
# ClipTrainer rounds it up to a bucket so that one CUDA graph per bucket can be replayed); the surplus rows
# hold zeros (b200clip_embed_tokens_packed_fwd / the packed attention kernels write them), stay finite through
# every row-wise kernel and keep exactly-zero gradients.  Without a caller-supplied count (drop-in
# `model(image, text)`), text_fwd reads it back from the device (one host sync) and only when the batch is
# large enough for that to pay (PACK_MIN_ROWS).
# What the MLP saves for its backward: the 8-bit code of QuickGELU'(c_fc(x)) instead of the bf16 pre-activation
# (B200CLIP_EPI_QUICKGELU_D8 / _BWD_D8, csrc/gemm.cu).  The backward needs nothing else of the pre-activation; the code is
# computed from the fp32 accumulator on a grid of 1/210 (rms error 1.4e-3 -- what re-evaluating the derivative from a
# bf16-rounded pre-activation costs anyway), the c_fc forward writes 3 instead of 4 bytes per element and the c_proj
# dgrad reads 1 instead of 2 and needs no MUFU.  B200CLIP_QGELU_D8=0 saves the pre-activation like round 1.
QGELU_D8 = os.environ.get("B200CLIP_QGELU_D8", "1") != "0"
PACK_TEXT = os.environ.get("B200CLIP_PACK_TEXT", "1") != "0"
PACK_MIN_ROWS = int(os.environ.get("B200CLIP_PACK_MIN_ROWS", "4096"))


@dataclass
class BlockSaved:
    x: torch.Tensor = None
    mean1: torch.Tensor = None
    rstd1: torch.Tensor = None
    h1: torch.Tensor = None
    qkv: torch.Tensor = None
    a: torch.Tensor = None
    lse: torch.Tensor = None
    x2: torch.Tensor = None
    mean2: torch.Tensor = None
    rstd2: torch.Tensor = None
    h2: torch.Tensor = None
    f: torch.Tensor = None
    g: torch.Tensor = None
    pool: torch.Tensor = None   # last block only: int64 rows of the pooled tokens (see blocks_fwd)
    a_p: torch.Tensor = None    # ... and the attention output gathered at those rows


@dataclass
class TowerSaved:
    blocks: list = field(default_factory=list)
    extra: dict = field(default_factory=dict)
    recompute: bool = False   # keep block inputs only; blocks_bwd re-runs each block's forward (see blocks_fwd)


def _blk(prefix: str, i: int) -> str:
    return f"{prefix}resblocks.{i}."


# ------------------------------------------------------------------------------------------------
# clip.model.ResidualAttentionBlock:  x = x + attn(ln_1(x)) ; x = x + c_proj(QuickGELU(c_fc(ln_2(x))))
def _block_fwd(W, p, x, B, S, H, causal, save, cu=None):
    """One residual block on every token -> (y fp32, BlockSaved or None)."""
    if save:
        h1, mean1, rstd1 = O.layernorm_fwd(x, W[p + "ln_1.weight"], W[p + "ln_1.bias"], want_stats=True)
    else:
        h1 = O.layernorm_fwd(x, W[p + "ln_1.weight"], W[p + "ln_1.bias"])
    qkv = O.linear_fwd(h1, W[p + "attn.in_proj_weight"], W[p + "attn.in_proj_bias"])
    if save:
        a, lse = O.attn_fwd(qkv, B, S, H, causal, want_lse=True, cu=cu)
    else:
        a, lse = O.attn_fwd(qkv, B, S, H, causal, cu=cu), None
    x2 = O.linear_fwd(a, W[p + "attn.out_proj.weight"], W[p + "attn.out_proj.bias"], epilogue=L.EPI_RESIDUAL, aux=x,
                      out_dtype=f32)
    if save:
        h2, mean2, rstd2 = O.layernorm_fwd(x2, W[p + "ln_2.weight"], W[p + "ln_2.bias"], want_stats=True)
        f = torch.empty((x.shape[0], 4 * x.shape[1]), device=x.device, dtype=torch.uint8 if QGELU_D8 else bf16)
    else:
        h2 = O.layernorm_fwd(x2, W[p + "ln_2.weight"], W[p + "ln_2.bias"])
        f = None
    g = O.linear_fwd(h2, W[p + "mlp.c_fc.weight"], W[p + "mlp.c_fc.bias"], preact=f,
                     epilogue=L.EPI_QUICKGELU_D8 if (f is not None and QGELU_D8) else L.EPI_QUICKGELU)
    y = O.linear_fwd(g, W[p + "mlp.c_proj.weight"], W[p + "mlp.c_proj.bias"], epilogue=L.EPI_RESIDUAL, aux=x2,
                     out_dtype=f32)
    return y, (BlockSaved(x, mean1, rstd1, h1, qkv, a, lse, x2, mean2, rstd2, h2, f, g) if save else None)


def blocks_fwd(W, prefix, layers, x, B, S, H, causal, saved: TowerSaved | None, pool_rows=None, cu=None):
    """``pool_rows`` (int32 [B], rows of ``x``): the only tokens of the LAST block's output that the
    tower uses (CLS rows for ``visual``: ``x[:, 0, :]``; EOT rows for text: ``x[arange, argmax]``).
    Every token of the last block still feeds attention through K and V, but its out_proj, ln_2 and
    MLP outputs are dead for all other tokens (upstream computes and discards them), and so are the
    corresponding gradients (exactly zero).  With ``pool_rows`` the last block runs those three
    operators on the B pooled rows only and returns ``[B, d]`` -- same features, same gradients,
    18/24 of the last block's GEMM work less (6 % of a 12-layer tower, forward and backward).

    ``saved.recompute`` (activation recompute, BASELINE config 5: ViT-L/14@336px at 512 pairs / GPU would keep
    255 GB of activations): a block keeps only its INPUT (fp32 residual stream, 4 bytes x d per token) and
    blocks_bwd re-runs its forward right before its backward -- one extra forward of arithmetic, 1/9 of the
    memory.  The algorithmic FLOP count of a step does not change."""
    for i in range(layers):
        p = _blk(prefix, i)
        if pool_rows is not None and i == layers - 1:
            return _last_block_fwd(W, p, x, B, S, H, causal, saved, pool_rows, cu)
        if saved is not None and saved.recompute:
            y, _ = _block_fwd(W, p, x, B, S, H, causal, False, cu)
            saved.blocks.append(BlockSaved(x=x))
        else:
            y, blk = _block_fwd(W, p, x, B, S, H, causal, saved is not None, cu)
            if saved is not None:
                saved.blocks.append(blk)
        x = y
    return x


def _last_block_fwd(W, p, x, B, S, H, causal, saved, pool_rows, cu=None):
    save = saved is not None
    if save:
        h1, mean1, rstd1 = O.layernorm_fwd(x, W[p + "ln_1.weight"], W[p + "ln_1.bias"], want_stats=True)
        qkv = O.linear_fwd(h1, W[p + "attn.in_proj_weight"], W[p + "attn.in_proj_bias"])
        a, lse = O.attn_fwd(qkv, B, S, H, causal, want_lse=True, cu=cu)   # only pooled rows of `a` are read below
    else:
        h1, mean1, rstd1 = O.layernorm_fwd(x, W[p + "ln_1.weight"], W[p + "ln_1.bias"]), None, None
        qkv = O.linear_fwd(h1, W[p + "attn.in_proj_weight"], W[p + "attn.in_proj_bias"])
        a, lse = O.attn_fwd(qkv, B, S, H, causal, cu=cu), None
    rows = pool_rows
    a_p = O.gather_rows(a, rows)   # [B, d] bf16
    x_p = O.gather_rows(x, rows)   # [B, d] fp32 residual stream at the pooled tokens
    x2 = O.linear_fwd(a_p, W[p + "attn.out_proj.weight"], W[p + "attn.out_proj.bias"], epilogue=L.EPI_RESIDUAL, aux=x_p,
                      out_dtype=f32)
    if save:
        h2, mean2, rstd2 = O.layernorm_fwd(x2, W[p + "ln_2.weight"], W[p + "ln_2.bias"], want_stats=True)
        f = torch.empty((x2.shape[0], 4 * x2.shape[1]), device=x.device, dtype=torch.uint8 if QGELU_D8 else bf16)
    else:
        h2, mean2, rstd2 = O.layernorm_fwd(x2, W[p + "ln_2.weight"], W[p + "ln_2.bias"]), None, None
        f = None
    g = O.linear_fwd(h2, W[p + "mlp.c_fc.weight"], W[p + "mlp.c_fc.bias"], preact=f,
                     epilogue=L.EPI_QUICKGELU_D8 if (f is not None and QGELU_D8) else L.EPI_QUICKGELU)
    y = O.linear_fwd(g, W[p + "mlp.c_proj.weight"], W[p + "mlp.c_proj.bias"], epilogue=L.EPI_RESIDUAL, aux=x2,
                     out_dtype=f32)
    if save:
        saved.blocks.append(BlockSaved(x, mean1, rstd1, h1, qkv, a, lse, x2, mean2, rstd2, h2, f, g, rows, a_p))
    return y


def blocks_bwd(W, G, prefix, layers, dy, B, S, H, causal, saved: TowerSaved, on_layer_done=None, cu=None,
               wgrad_stream=None):
    """Bias gradients are column sums of the gradient stream; they are produced by the kernel that
    WRITES each tensor (GEMM epilogue ``colsum`` / LayerNorm-backward ``dx_colsum``) instead of by
    separate reduction passes -- only the Q third of dqkv (written by the attention backward) and the
    incoming dy of the last block use the stand-alone colsum kernel.

    ``wgrad_stream``: the weight-gradient GEMMs (and the Q-bias column sum) have no consumer before the
    optimiser, so they leave the critical path dgrad -> LayerNorm' -> dgrad -> attention' -> dgrad: they are
    issued on this second stream, each behind the kernel that produced its operand, and fill the SMs the
    critical path leaves idle (ramps, tails, partial waves -- at 128 pairs / GPU that is half the machine).
    A layer's activations are released one layer late, after an event says its weight gradients are done."""
    cur = torch.cuda.current_stream(dy.device) if wgrad_stream is not None else None
    held, held_ev = None, None

    def wgrad(dy_, x_, out_):
        if wgrad_stream is None:
            return O.linear_wgrad(dy_, x_, out_)
        wgrad_stream.wait_stream(cur)          # dy_ has just been produced on the tower's stream
        with torch.cuda.stream(wgrad_stream):
            O.linear_wgrad(dy_, x_, out_)

    O.colsum(dy, G[_blk(prefix, layers - 1) + "mlp.c_proj.bias"])
    for i in reversed(range(layers)):
        p = _blk(prefix, i)
        s: BlockSaved = saved.blocks[i]
        if s.qkv is None:   # activation recompute: only the block's input was kept
            _, s = _block_fwd(W, p, s.x, B, S, H, causal, True, cu)
        dy_in = dy
        # ---- MLP branch: y = x2 + c_proj(gelu(c_fc(ln_2(x2))))
        wgrad(dy, s.g, G[p + "mlp.c_proj.weight"])
        df = O.linear_dgrad(dy, W[p + "mlp.c_proj.weight"],
                            epilogue=L.EPI_QUICKGELU_BWD_D8 if s.f.dtype == torch.uint8 else L.EPI_QUICKGELU_BWD, aux=s.f,
                            colsum=G[p + "mlp.c_fc.bias"])
        wgrad(df, s.h2, G[p + "mlp.c_fc.weight"])
        dh2 = O.linear_dgrad(df, W[p + "mlp.c_fc.weight"])
        dx2 = O.layernorm_bwd(dh2, s.x2, W[p + "ln_2.weight"], s.mean2, s.rstd2, G[p + "ln_2.weight"],
                              G[p + "ln_2.bias"], dres=dy, dx_colsum=G[p + "attn.out_proj.bias"])
        # ---- attention branch: x2 = x + out_proj(attn(in_proj(ln_1(x))))
        wgrad(dx2, s.a if s.pool is None else s.a_p, G[p + "attn.out_proj.weight"])
        dx2_in = dx2
        # in_proj_bias gradient = column sums of dqkv = [dQ | dK | dV] without reading all of dqkv again:
        #   V third: sum_kv dV = sum_q (P^T dO) = sum_q dO because softmax rows sum to one -> the column
        #            sums of `da`, produced by this dgrad GEMM's epilogue;
        #   K third: sum_k dK = sum_q (sum_k dS[q,k]) Q[q] = 0 exactly (rows of dS sum to zero; a key
        #            bias shifts every score of a row equally) -> stays at the zero of the gradient buffer;
        #   Q third: a real reduction, the stand-alone colsum over the first d columns of dqkv.
        gb = G[p + "attn.in_proj_bias"]
        d_model = gb.numel() // 3
        da = O.linear_dgrad(dx2, W[p + "attn.out_proj.weight"], colsum=gb[2 * d_model:])
        if s.pool is not None:  # pooled last block: the gradients of every other token are exactly zero
            da = O.scatter_rows(da, s.pool, s.a.shape[0])
            dx2 = O.scatter_rows(dx2, s.pool, s.a.shape[0])
        dqkv = O.attn_bwd(s.qkv, s.a, s.lse, da, B, S, H, causal, cu=cu)
        if wgrad_stream is None:
            O.colsum(dqkv[:, :d_model], gb[:d_model])
        wgrad(dqkv, s.h1, G[p + "attn.in_proj_weight"])
        if wgrad_stream is not None:
            with torch.cuda.stream(wgrad_stream):
                O.colsum(dqkv[:, :d_model], gb[:d_model])
        dh1 = O.linear_dgrad(dqkv, W[p + "attn.in_proj_weight"])
        # dx of this LayerNorm is the dy of block i-1: its column sums are that block's c_proj bias gradient
        prev_bias = G[_blk(prefix, i - 1) + "mlp.c_proj.bias"] if i > 0 else None
        dy = O.layernorm_bwd(dh1, s.x, W[p + "ln_1.weight"], s.mean1, s.rstd1, G[p + "ln_1.weight"],
                             G[p + "ln_1.bias"], dres=dx2, dx_colsum=prev_bias)
        saved.blocks[i] = None  # release this layer's activations ...
        if wgrad_stream is not None:
            # ... one layer late: the PREVIOUS layer's weight gradients must have read them first
            if held_ev is not None:
                cur.wait_event(held_ev)
            held = (s, dy_in, df, dx2_in, dqkv)   # noqa: F841 -- keeps the operands of this layer's wgrads alive
            held_ev = torch.cuda.Event()
            held_ev.record(wgrad_stream)
        del s
        if on_layer_done is not None:  # every parameter gradient of layers >= i is final now (given wgrad_stream's work)
            on_layer_done(i)
    if wgrad_stream is not None:
        cur.wait_stream(wgrad_stream)
        del held
    return dy


# ------------------------------------------------------------------------------------------------
def _pool_project_fwd(W, x, row_index, ln_w, ln_b, proj, save):
    """ln(x[row_index]) @ proj -> fp32 [B, E]   (visual.ln_post + visual.proj / ln_final + text_projection)."""
    if save:
        pooled32, mean, rstd = O.layernorm_fwd(x, W[ln_w], W[ln_b], row_index=row_index, want_stats=True, out_dtype=f32)
    else:
        pooled32, mean, rstd = O.layernorm_fwd(x, W[ln_w], W[ln_b], row_index=row_index, out_dtype=f32), None, None
    # The projection input is only B rows, but rounding it to bf16 is the single largest term of the
    # logit error (100 * 2^-9 / sqrt(E) per feature): feed it as a hi + lo bf16 pair instead.
    pooled, lo = O.split_f32_to_bf16(pooled32)
    feat = O.gemm(pooled, W[proj], b_major=L.MAJOR_MN, out_dtype=f32)  # proj stored [d, E]
    O.gemm(lo, W[proj], b_major=L.MAJOR_MN, out=feat, accumulate=True)
    return feat, (pooled, mean, rstd)


def _pool_project_bwd(W, G, dfeat, x, row_index, ln_w, ln_b, proj, pooled, mean, rstd):
    dfeat_bf = dfeat.contiguous() if dfeat.dtype == bf16 else O.cast_f32_to_bf16(dfeat.contiguous())
    # dproj[d,E] += pooled^T dfeat ; dpooled = dfeat proj^T
    O.gemm(pooled, dfeat_bf, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, out=G[proj], split_k=0, accumulate=True)
    dpooled = O.gemm(dfeat_bf, W[proj])  # B operand = proj [d, E] read K-major (N = d, K = E)
    dx = torch.zeros(x.shape, device=x.device, dtype=bf16)  # the gradient stream is bf16
    O.layernorm_bwd(dpooled, x, W[ln_w], mean, rstd, G[ln_w], G[ln_b], row_index=row_index, dx=dx)
    return dx


# ------------------------------------------------------------------------------------------------
_VISION_INDEX = {}


def _vision_index(B, n, device):
    """(ridx, cls_rows): source row of every token of the assembled sequence (-1 = class embedding, else the
    patch row) and the rows of the CLS tokens.  Depends on (B, n) only: built once per shape."""
    key = (B, n, device)
    hit = _VISION_INDEX.get(key)
    if hit is None:
        g2 = n - 1
        ridx = torch.arange(-1, g2, device=device, dtype=i32).repeat(B, 1)
        ridx[:, 1:] += (torch.arange(B, device=device, dtype=i32) * g2)[:, None]
        cls_rows = torch.arange(B, device=device, dtype=i32) * n
        if len(_VISION_INDEX) >= 16:
            _VISION_INDEX.clear()
        hit = _VISION_INDEX[key] = (ridx.reshape(-1).contiguous(), cls_rows)
    return hit


# clip.model.VisionTransformer.forward
def vision_fwd(W, cfg, image, save: bool, recompute: bool = False):
    B = image.shape[0]
    p, n, d = cfg.vision_patch_size, cfg.vision_tokens, cfg.vision_width
    H = d // 64
    if image.dtype not in (bf16, f32, torch.uint8):   # uint8 = raw pixels: ToTensor + Normalize run inside im2col
        image = image.float()
    image = image.contiguous()
    kpad = W["conv1.weight"].shape[1]
    cols = O.im2col_patch(image, p, ldcols=kpad)                      # [B*g*g, kpad]
    patch = O.gemm(cols, W["conv1.weight"])                           # conv1 as GEMM -> [B*g*g, d]
    ridx, cls_rows = _vision_index(B, n, image.device)
    pre = torch.empty((B * n, d), device=image.device, dtype=bf16) if save else None
    r = O.layernorm_fwd(patch, W["ln_pre.weight"], W["ln_pre.bias"], rows=B * n, row_index=ridx,
                        neg_row=W["class_embedding"], add=W["positional_embedding"], add_period=n, pre_out=pre,
                        want_stats=save, out_dtype=f32)  # the residual stream is fp32
    x, mean0, rstd0 = r if save else (r, None, None)
    saved = TowerSaved(recompute=recompute) if save else None
    x = blocks_fwd(W, "transformer.", cfg.vision_layers, x, B, n, H, False, saved,
                   cls_rows if POOL_LAST_BLOCK else None)
    if POOL_LAST_BLOCK:
        cls_rows = None  # x is already [B, d]: the CLS tokens
    feat, head = _pool_project_fwd(W, x, cls_rows, "ln_post.weight", "ln_post.bias", "proj", save)
    if save:
        saved.extra = dict(B=B, cols=cols, pre=pre, mean0=mean0, rstd0=rstd0, x_last=x, cls_rows=cls_rows, head=head)
    return feat, saved


def vision_bwd(W, G, cfg, saved: TowerSaved, dfeat, on_layer_done=None, wgrad_stream=None):
    e = saved.extra
    B, n, d = e["B"], cfg.vision_tokens, cfg.vision_width
    H = d // 64
    pooled, mean, rstd = e["head"]
    dx = _pool_project_bwd(W, G, dfeat, e["x_last"], e["cls_rows"], "ln_post.weight", "ln_post.bias", "proj", pooled,
                           mean, rstd)
    dx = blocks_bwd(W, G, "transformer.", cfg.vision_layers, dx, B, n, H, False, saved, on_layer_done,
                    wgrad_stream=wgrad_stream)
    dpre = O.layernorm_bwd(dx, e["pre"], W["ln_pre.weight"], e["mean0"], e["rstd0"], G["ln_pre.weight"],
                           G["ln_pre.bias"])
    dpatch = O.vision_assemble_bwd(dpre, B, n, G["positional_embedding"], G["class_embedding"])
    O.linear_wgrad(dpatch, e["cols"], G["conv1.weight"])


# ------------------------------------------------------------------------------------------------
# clip.model.CLIP.encode_text
def text_fwd(W, cfg, text, save: bool, rows=None, recompute: bool = False):
    """``rows``: static row count of the packed layout (>= sum of EOT position + 1; see PACK_TEXT); None lets
    this function decide -- unpacked for small batches, else packed to the exact count (one host sync)."""
    B, S = text.shape
    d = cfg.transformer_width
    H = cfg.transformer_heads
    ids = text.to(i32).contiguous()
    saved = TowerSaved(recompute=recompute) if save else None
    cu = None
    if rows is None and ids.is_cuda and torch.cuda.is_current_stream_capturing():
        pack = False   # the exact row count needs a host sync: not inside a graph capture (small no-grad forwards)
    else:
        pack = PACK_TEXT and S <= 128 and (rows is not None or B * S >= PACK_MIN_ROWS)
    if pack:
        # keep positions 0 .. EOT of every caption only
        cu, eot = O.text_pack_plan(ids, rows if rows is not None else B * S)
        if rows is None:
            rows = max(1, int(cu[B].item()))   # dynamic shape: host sync
        x = O.embed_tokens_packed_fwd(ids, W["token_embedding.weight"], W["positional_embedding"], cu, rows, out_dtype=f32)
    else:
        x, eot = O.embed_tokens_fwd(ids, W["token_embedding.weight"], W["positional_embedding"], out_dtype=f32)
    x = blocks_fwd(W, "transformer.", cfg.transformer_layers, x, B, S, H, True, saved,
                   eot if POOL_LAST_BLOCK else None, cu)
    if POOL_LAST_BLOCK:
        eot = None  # x is already [B, d]: the EOT tokens
    feat, head = _pool_project_fwd(W, x, eot, "ln_final.weight", "ln_final.bias", "text_projection", save)
    if save:
        saved.extra = dict(B=B, S=S, ids=ids, eot=eot, x_last=x, head=head, cu=cu)
    return feat, saved


def text_bwd(W, G, cfg, saved: TowerSaved, dfeat, on_layer_done=None, wgrad_stream=None):
    e = saved.extra
    B, S = e["B"], e["S"]
    H = cfg.transformer_heads
    pooled, mean, rstd = e["head"]
    dx = _pool_project_bwd(W, G, dfeat, e["x_last"], e["eot"], "ln_final.weight", "ln_final.bias", "text_projection",
                           pooled, mean, rstd)
    dx = blocks_bwd(W, G, "transformer.", cfg.transformer_layers, dx, B, S, H, True, saved, on_layer_done, e.get("cu"),
                    wgrad_stream=wgrad_stream)
    if e.get("cu") is not None:  # packed rows; the dropped positions have zero gradient
        O.embed_tokens_packed_bwd(e["ids"], dx, e["cu"], G["token_embedding.weight"], G["positional_embedding"])
    else:
        O.embed_tokens_bwd(e["ids"], dx, G["token_embedding.weight"], G["positional_embedding"])


# ------------------------------------------------------------------------------------------------
# fp32 check mode: the same forward with fp32 activations and tensor-core-free fp32 kernels
# (BASELINE.json: logits within 1e-4 of the reference in an fp32 check mode).  Forward only.
def _blocks_fwd_f32(W, prefix, layers, x, B, S, H, causal):
    for i in range(layers):
        p = _blk(prefix, i)
        h1 = O.layernorm_fwd(x, W[p + "ln_1.weight"], W[p + "ln_1.bias"], out_dtype=f32)
        qkv = O.check_gemm_f32(h1, W[p + "attn.in_proj_weight"], bias=W[p + "attn.in_proj_bias"])
        a = O.check_attn_fwd_f32(qkv, B, S, H, causal)
        x = O.check_gemm_f32(a, W[p + "attn.out_proj.weight"], bias=W[p + "attn.out_proj.bias"], residual=x)
        h2 = O.layernorm_fwd(x, W[p + "ln_2.weight"], W[p + "ln_2.bias"], out_dtype=f32)
        g = O.check_gemm_f32(h2, W[p + "mlp.c_fc.weight"], bias=W[p + "mlp.c_fc.bias"], quickgelu=True)
        x = O.check_gemm_f32(g, W[p + "mlp.c_proj.weight"], bias=W[p + "mlp.c_proj.bias"], residual=x)
    return x


def vision_fwd_f32(W, cfg, image):
    B = image.shape[0]
    p, n, d = cfg.vision_patch_size, cfg.vision_tokens, cfg.vision_width
    kpad = W["conv1.weight"].shape[1]
    if image.dtype == torch.uint8:   # raw pixels: upstream's ToTensor + Normalize (clip._transform)
        mean = torch.tensor((0.48145466, 0.4578275, 0.40821073), device=image.device).view(1, 3, 1, 1)
        std = torch.tensor((0.26862954, 0.26130258, 0.27577711), device=image.device).view(1, 3, 1, 1)
        image = (image.float() / 255.0 - mean) / std
    cols = O.check_im2col_f32(image.float().contiguous(), p, kpad)
    patch = O.check_gemm_f32(cols, W["conv1.weight"])
    g2 = n - 1
    ridx = torch.arange(-1, g2, device=image.device, dtype=i32).repeat(B, 1)
    ridx[:, 1:] += (torch.arange(B, device=image.device, dtype=i32) * g2)[:, None]
    x = O.layernorm_fwd(patch, W["ln_pre.weight"], W["ln_pre.bias"], rows=B * n, row_index=ridx.reshape(-1),
                        neg_row=W["class_embedding"], add=W["positional_embedding"], add_period=n, out_dtype=f32)
    x = _blocks_fwd_f32(W, "transformer.", cfg.vision_layers, x, B, n, d // 64, False)
    cls_rows = torch.arange(B, device=image.device, dtype=i32) * n
    pooled = O.layernorm_fwd(x, W["ln_post.weight"], W["ln_post.bias"], row_index=cls_rows, out_dtype=f32)
    return O.check_gemm_f32(pooled, W["proj"], b_major=L.MAJOR_MN)


def text_fwd_f32(W, cfg, text):
    B, S = text.shape
    ids = text.to(i32).contiguous()
    x, eot = O.embed_tokens_fwd(ids, W["token_embedding.weight"], W["positional_embedding"], out_dtype=f32)
    x = _blocks_fwd_f32(W, "transformer.", cfg.transformer_layers, x, B, S, cfg.transformer_heads, True)
    pooled = O.layernorm_fwd(x, W["ln_final.weight"], W["ln_final.bias"], row_index=eot, out_dtype=f32)
    return O.check_gemm_f32(pooled, W["text_projection"], b_major=L.MAJOR_MN)
