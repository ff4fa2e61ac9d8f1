"""Thin torch-tensor wrappers over the C ABI (one function per entry point of include/b200clip.h).

These only marshal pointers / extents / the current CUDA stream; every byte of arithmetic runs
in ``libb200clip.so``.  Tensors must live on a CUDA device (there is no CPU path).
"""
from __future__ import annotations

import torch

from . import lib as L

bf16, f32, i32 = torch.bfloat16, torch.float32, torch.int32


def _ctx_stream(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("construction_clip_b200 ops need CUDA tensors on an sm_100 device (no CPU fallback)")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return L.ctx(dev), torch.cuda.current_stream(t.device).cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _dt(t):
    if t.dtype == bf16:
        return L.DT_BF16
    if t.dtype == f32:
        return L.DT_F32
    raise RuntimeError(f"unsupported dtype {t.dtype}")


def _row_major(t: torch.Tensor, name: str):
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError(f"{name}: expected a 2-D tensor with unit inner stride, got {tuple(t.shape)} strides {t.stride()}")
    return t.stride(0)


# ------------------------------------------------------------------------------------------ GEMM
# bench.py sets this to a list to time every GEMM launch with CUDA events on the launching stream
# (the roofline of the dominant kernel is measured live, inside the timed region)
GEMM_PROFILE = None


def gemm(a, b, *, a_major=L.MAJOR_K, b_major=L.MAJOR_K, bias=None, aux=None, preact=None, scale=None, colsum=None,
         epilogue=L.EPI_NONE, out=None, out_dtype=bf16, split_k=1, accumulate=False):
    """C[M,N] = sum_k A(m,k) B(n,k) (+bias) -> epilogue.  ``a``/``b`` are the STORED matrices:
    K-major operands are [rows, K]; MN-major operands are [K, rows]."""
    lda, ldb = _row_major(a, "a"), _row_major(b, "b")
    if a_major == L.MAJOR_K:
        M, K = a.shape
    else:
        K, M = a.shape
    if b_major == L.MAJOR_K:
        N, Kb = b.shape
    else:
        Kb, N = b.shape
    if K != Kb:
        raise RuntimeError(f"gemm: reduction extents differ ({K} vs {Kb})")
    assert a.dtype == bf16 and b.dtype == bf16
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    ldc = _row_major(out, "out")
    if preact is not None:   # bf16 pre-activation, or the uint8 codes of QuickGELU' (EPI_QUICKGELU_D8: same extents, pitch in bytes)
        assert _row_major(preact, "preact") == ldc
        assert preact.dtype == (torch.uint8 if epilogue == L.EPI_QUICKGELU_D8 else bf16)
    if epilogue == L.EPI_QUICKGELU_BWD_D8:
        assert aux is not None and aux.dtype == torch.uint8
    ctx, st = _ctx_stream(a)
    prof = GEMM_PROFILE
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    L.check(L.load().b200clip_gemm_bf16(
        ctx, a.data_ptr(), lda, a_major, b.data_ptr(), ldb, b_major, out.data_ptr(), ldc,
        L.DT_F32 if out.dtype == f32 else L.DT_BF16, _ptr(bias), _ptr(aux),
        _row_major(aux, "aux") if aux is not None else 0, _ptr(preact), _ptr(scale), _ptr(colsum), M, N, K, epilogue,
        split_k, 1 if accumulate else 0, st), "gemm_bf16")
    if prof is not None:
        e1.record()
        nbytes = (a.numel() + b.numel()) * 2 + out.numel() * out.element_size()
        nbytes += aux.numel() * aux.element_size() if aux is not None else 0
        nbytes += preact.numel() * preact.element_size() if preact is not None else 0
        prof.append((e0, e1, 2.0 * M * N * K, (a_major, b_major, M, N, K, nbytes)))
    return out


def linear_fwd(x, w, bias=None, *, epilogue=L.EPI_NONE, aux=None, preact=None, out=None, out_dtype=bf16):
    """y = x @ w.T + bias  (w stored [N,K] like nn.Linear.weight)."""
    return gemm(x, w, bias=bias, aux=aux, preact=preact, epilogue=epilogue, out=out, out_dtype=out_dtype)


def linear_dgrad(dy, w, *, epilogue=L.EPI_NONE, aux=None, out=None, colsum=None):
    """dx = dy @ w   (w stored [N',K']; read as an MN-major B operand, no transpose copy).
    ``colsum`` (fp32 [K']) accumulates the column sums of dx."""
    return gemm(dy, w, b_major=L.MAJOR_MN, epilogue=epilogue, aux=aux, out=out, colsum=colsum)


def linear_wgrad(dy, x, out):
    """out[N',K'] (fp32) += dy^T @ x   (both operands read MN-major; split-K with fp32 atomics)."""
    return gemm(dy, x, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, out=out, split_k=0, accumulate=True)


# ------------------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x, gamma, beta, *, rows=None, row_index=None, neg_row=None, add=None, add_period=0, out=None,
                  pre_out=None, want_stats=False, eps=1e-5, out_dtype=bf16):
    ldx = _row_major(x, "x")
    d = x.shape[1]
    if rows is None:
        rows = row_index.numel() if row_index is not None else x.shape[0]
    if out is None:
        out = torch.empty((rows, d), device=x.device, dtype=out_dtype)
    mean = rstd = None
    if want_stats:
        mean = torch.empty(rows, device=x.device, dtype=f32)
        rstd = torch.empty(rows, device=x.device, dtype=f32)
    ctx, st = _ctx_stream(x)
    L.check(L.load().b200clip_layernorm_fwd(
        ctx, x.data_ptr(), ldx, _ptr(row_index), _ptr(neg_row), _ptr(add), add_period, gamma.data_ptr(), beta.data_ptr(),
        out.data_ptr(), _row_major(out, "out"), _ptr(pre_out), _ptr(mean), _ptr(rstd), rows, d, eps, _dt(x), _dt(out),
        st),
        "layernorm_fwd")
    return (out, mean, rstd) if want_stats else out


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, *, row_index=None, dres=None, dx=None, dx_colsum=None):
    rows, d = dy.shape
    if dx is None:
        dx = torch.empty((rows, d), device=dy.device, dtype=bf16)
    ctx, st = _ctx_stream(dy)
    L.check(L.load().b200clip_layernorm_bwd(
        ctx, dy.data_ptr(), _row_major(dy, "dy"), x.data_ptr(), _row_major(x, "x"), _ptr(row_index),
        gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _ptr(dres),
        _row_major(dres, "dres") if dres is not None else 0, dx.data_ptr(), _row_major(dx, "dx"),
        _ptr(dgamma), _ptr(dbeta), _ptr(dx_colsum), rows, d, _dt(x), st), "layernorm_bwd")
    return dx


# ------------------------------------------------------------------------------------- attention
def attn_fwd(qkv, B, S, H, causal, out=None, want_lse=False, cu=None):
    """``cu`` (int32 [B+1]): packed rows -- sample b owns rows cu[b] .. cu[b+1]-1 (<= S) of ``qkv``; rows
    past cu[B] (the surplus of a static row count) are zero-filled in ``out`` by the kernel."""
    if cu is not None:
        rows = qkv.shape[0]
        assert qkv.dtype == bf16 and qkv.is_contiguous() and qkv.shape[1] == 3 * H * 64 and cu.dtype == i32
        if out is None:
            out = torch.empty((rows, H * 64), device=qkv.device, dtype=bf16)
        lse = torch.empty(rows * H, device=qkv.device, dtype=f32) if want_lse else None
        ctx, st = _ctx_stream(qkv)
        L.check(L.load().b200clip_attn_fwd_varlen(ctx, qkv.data_ptr(), out.data_ptr(), _ptr(lse), cu.data_ptr(), B, S, H,
                                                  rows, 1 if causal else 0, st), "attn_fwd_varlen")
        return (out, lse) if want_lse else out
    assert qkv.dtype == bf16 and qkv.is_contiguous() and qkv.shape == (B * S, 3 * H * 64)
    if out is None:
        out = torch.empty((B * S, H * 64), device=qkv.device, dtype=bf16)
    lse = torch.empty(B * H * S, device=qkv.device, dtype=f32) if want_lse else None
    ctx, st = _ctx_stream(qkv)
    L.check(L.load().b200clip_attn_fwd(ctx, qkv.data_ptr(), out.data_ptr(), _ptr(lse), B, S, H, 1 if causal else 0, st),
            "attn_fwd")
    return (out, lse) if want_lse else out


def attn_bwd(qkv, out, lse, dout, B, S, H, causal, dqkv=None, cu=None):
    if cu is not None:
        rows = qkv.shape[0]
        assert qkv.is_contiguous() and dout.is_contiguous() and out.is_contiguous() and dout.shape == (rows, H * 64)
        if dqkv is None:
            dqkv = torch.empty_like(qkv)
        ctx, st = _ctx_stream(qkv)
        L.check(L.load().b200clip_attn_bwd_varlen(ctx, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), dout.data_ptr(),
                                                  dqkv.data_ptr(), cu.data_ptr(), B, S, H, rows, 1 if causal else 0, st),
                "attn_bwd_varlen")
        return dqkv
    assert qkv.is_contiguous() and dout.is_contiguous() and out.is_contiguous() and dout.shape == (B * S, H * 64)
    if dqkv is None:
        dqkv = torch.empty_like(qkv)
    ws = torch.empty(B * H * S, device=qkv.device, dtype=f32) if S > 128 else None
    ctx, st = _ctx_stream(qkv)
    L.check(L.load().b200clip_attn_bwd(ctx, qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), dout.data_ptr(),
                                       dqkv.data_ptr(), _ptr(ws), ws.numel() * 4 if ws is not None else 0, B, S, H,
                                       1 if causal else 0, st), "attn_bwd")
    return dqkv


# ------------------------------------------------------------------------------------- embedding
def embed_tokens_fwd(ids, table, pos, out_dtype=bf16):
    B, S = ids.shape
    V, d = table.shape
    assert ids.dtype == i32 and ids.is_contiguous() and table.is_contiguous() and pos.is_contiguous()
    out = torch.empty((B * S, d), device=table.device, dtype=out_dtype)
    eot = torch.empty(B, device=table.device, dtype=i32)
    ctx, st = _ctx_stream(table)
    L.check(L.load().b200clip_embed_tokens_fwd(ctx, ids.data_ptr(), table.data_ptr(), pos.data_ptr(), out.data_ptr(),
                                               _dt(out), eot.data_ptr(), B, S, d, V, st), "embed_tokens_fwd")
    return out, eot


def embed_tokens_bwd(ids, dout, dtable, dpos):
    B, S = ids.shape
    V, d = dtable.shape
    assert dout.is_contiguous() and dtable.dtype == f32 and dpos.dtype == f32
    ctx, st = _ctx_stream(dout)
    L.check(L.load().b200clip_embed_tokens_bwd(ctx, ids.data_ptr(), dout.data_ptr(), dtable.data_ptr(), dpos.data_ptr(),
                                               B, S, d, V, st), "embed_tokens_bwd")


def text_pack_plan(ids, rows_cap):
    """-> (cu int32 [B+1], eot_row int32 [B]) of the packed text layout (caption b keeps argmax + 1 rows)."""
    B, S = ids.shape
    assert ids.dtype == i32 and ids.is_contiguous()
    cu = torch.empty(B + 1, device=ids.device, dtype=i32)
    eot = torch.empty(B, device=ids.device, dtype=i32)
    ctx, st = _ctx_stream(ids)
    L.check(L.load().b200clip_text_pack_plan(ctx, ids.data_ptr(), cu.data_ptr(), eot.data_ptr(), B, S, int(rows_cap), st),
            "text_pack_plan")
    return cu, eot


def embed_tokens_packed_fwd(ids, table, pos, cu, rows_total, out_dtype=bf16):
    B, S = ids.shape
    V, d = table.shape
    assert ids.dtype == i32 and ids.is_contiguous() and table.is_contiguous() and pos.is_contiguous() and cu.dtype == i32
    out = torch.empty((int(rows_total), d), device=table.device, dtype=out_dtype)
    ctx, st = _ctx_stream(table)
    L.check(L.load().b200clip_embed_tokens_packed_fwd(ctx, ids.data_ptr(), table.data_ptr(), pos.data_ptr(), cu.data_ptr(),
                                                      out.data_ptr(), _dt(out), B, S, d, V, int(rows_total), st),
            "embed_tokens_packed_fwd")
    return out


def embed_tokens_packed_bwd(ids, dout, cu, dtable, dpos):
    B, S = ids.shape
    V, d = dtable.shape
    assert dout.is_contiguous() and dout.dtype == bf16 and dtable.dtype == f32 and dpos.dtype == f32
    ctx, st = _ctx_stream(dout)
    L.check(L.load().b200clip_embed_tokens_packed_bwd(ctx, ids.data_ptr(), dout.data_ptr(), cu.data_ptr(),
                                                      dtable.data_ptr(), dpos.data_ptr(), B, S, d, V, st),
            "embed_tokens_packed_bwd")


def gather_rows(src, idx):
    """dst[i, :] = src[idx[i], :]  (idx int32)."""
    ld = _row_major(src, "src")
    assert idx.dtype == i32 and idx.is_contiguous()
    n, d = idx.numel(), src.shape[1]
    dst = torch.empty((n, d), device=src.device, dtype=src.dtype)
    ctx, st = _ctx_stream(src)
    es = src.element_size()
    L.check(L.load().b200clip_gather_rows(ctx, src.data_ptr(), ld * es, idx.data_ptr(), dst.data_ptr(), n, d * es, st),
            "gather_rows")
    return dst


def scatter_rows(src, idx, rows):
    """zeros([rows, d]) with dst[idx[i], :] = src[i, :]."""
    assert idx.dtype == i32 and idx.is_contiguous() and src.is_contiguous()
    n, d = src.shape
    dst = torch.empty((int(rows), d), device=src.device, dtype=src.dtype)
    ctx, st = _ctx_stream(src)
    L.check(L.load().b200clip_scatter_rows(ctx, src.data_ptr(), idx.data_ptr(), dst.data_ptr(), int(rows), n,
                                           d * src.element_size(), 1, st), "scatter_rows")
    return dst


# ---------------------------------------------------------------------------------- patch embed
def im2col_patch(image, patch, ldcols=None):
    B, C, R, R2 = image.shape
    assert C == 3 and R == R2 and image.is_contiguous() and image.dtype in (bf16, f32, torch.uint8)
    g = R // patch
    k = 3 * patch * patch
    if ldcols is None:
        ldcols = (k + 63) // 64 * 64
    cols = torch.empty((B * g * g, ldcols), device=image.device, dtype=bf16)
    ctx, st = _ctx_stream(image)
    in_dt = {f32: L.DT_F32, bf16: L.DT_BF16, torch.uint8: L.DT_U8}[image.dtype]   # uint8: raw pixels, Normalize fused
    L.check(L.load().b200clip_im2col_patch(ctx, image.data_ptr(), in_dt,
                                           cols.data_ptr(), ldcols, B, R, patch, st), "im2col_patch")
    return cols


def resize_crop_u8(src, xbounds, xcoef, ybounds, ycoef, row0, rows, R):
    """src uint8 [H, W, 3] (cuda) -> uint8 [3, R, R]: Pillow-exact bicubic resize + centre crop (tables from data.py)."""
    assert src.dtype == torch.uint8 and src.dim() == 3 and src.shape[2] == 3 and src.stride(2) == 1 and src.stride(1) == 3
    H, W, _ = src.shape
    tmp = torch.empty((rows, R, 3), device=src.device, dtype=torch.uint8)
    dst = torch.empty((3, R, R), device=src.device, dtype=torch.uint8)
    ctx, st = _ctx_stream(src)
    L.check(L.load().b200clip_resize_crop_u8(ctx, src.data_ptr(), H, W, src.stride(0), xbounds.data_ptr(), xcoef.data_ptr(),
                                             xcoef.shape[1], ybounds.data_ptr(), ycoef.data_ptr(), ycoef.shape[1], row0,
                                             rows, tmp.data_ptr(), dst.data_ptr(), R, st), "resize_crop_u8")
    return dst


def colsum(x, out):
    M, N = x.shape
    ctx, st = _ctx_stream(x)
    L.check(L.load().b200clip_colsum(ctx, x.data_ptr(), _row_major(x, "x"), out.data_ptr(), M, N, st), "colsum")
    return out


def vision_assemble_bwd(dpre, B, n, dpos, dcls):
    d = dpre.shape[1]
    dpatch = torch.empty((B * (n - 1), d), device=dpre.device, dtype=bf16)
    ctx, st = _ctx_stream(dpre)
    L.check(L.load().b200clip_vision_assemble_bwd(ctx, dpre.data_ptr(), dpatch.data_ptr(), dpos.data_ptr(),
                                                  dcls.data_ptr(), B, n, d, st), "vision_assemble_bwd")
    return dpatch


# -------------------------------------------------------------------------------- heads and loss
def l2norm_fwd(x):
    B, E = x.shape
    assert x.dtype == f32 and x.is_contiguous()
    y = torch.empty_like(x)
    inv = torch.empty(B, device=x.device, dtype=f32)
    ctx, st = _ctx_stream(x)
    L.check(L.load().b200clip_l2norm_fwd(ctx, x.data_ptr(), y.data_ptr(), inv.data_ptr(), B, E, st), "l2norm_fwd")
    return y, inv


def l2norm_bwd(dy, y, inv):
    B, E = y.shape
    assert dy.dtype == f32 and dy.is_contiguous()
    dx = torch.empty((B, E), device=y.device, dtype=bf16)
    ctx, st = _ctx_stream(y)
    L.check(L.load().b200clip_l2norm_bwd(ctx, dy.data_ptr(), y.data_ptr(), inv.data_ptr(), dx.data_ptr(), B, E, st),
            "l2norm_bwd")
    return dx


def cast_f32_to_bf16(src, dst=None):
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=bf16)
    ctx, st = _ctx_stream(src)
    L.check(L.load().b200clip_cast_f32_to_bf16(ctx, src.data_ptr(), dst.data_ptr(), src.numel(), st), "cast_f32_to_bf16")
    return dst


def split_f32_to_bf16(src):
    """(hi, lo) bf16 pair with hi + lo ~= src to ~16 mantissa bits."""
    src = src.contiguous()
    hi = torch.empty(src.shape, device=src.device, dtype=bf16)
    lo = torch.empty(src.shape, device=src.device, dtype=bf16)
    ctx, st = _ctx_stream(src)
    L.check(L.load().b200clip_split_f32_to_bf16(ctx, src.data_ptr(), hi.data_ptr(), lo.data_ptr(), src.numel(), st),
            "split_f32_to_bf16")
    return hi, lo


def logits(img_n, txt_n, logit_scale):
    Bi, E = img_n.shape
    Bt = txt_n.shape[0]
    assert img_n.dtype == f32 and txt_n.dtype == f32 and logit_scale.dtype == f32
    out = torch.empty((Bi, Bt), device=img_n.device, dtype=f32)
    ctx, st = _ctx_stream(img_n)
    L.check(L.load().b200clip_logits(ctx, img_n.data_ptr(), txt_n.data_ptr(), logit_scale.data_ptr(), out.data_ptr(),
                                     Bi, Bt, E, st), "logits")
    return out


def clip_loss_workspace(device, Bl, Bg, E):
    dev = device.index if device.index is not None else torch.cuda.current_device()
    n = L.load().b200clip_clip_loss_workspace_bytes(L.ctx(dev), Bl, Bg, E)
    if n < 0:
        raise RuntimeError("clip_loss_workspace_bytes: bad arguments")
    return torch.empty(n, device=device, dtype=torch.uint8)


def clip_loss_fwd(img_all, txt_all, logit_scale, row0, Bl, ws):
    """Returns (lse_i [Bl], lse_t [Bl], loss_sum [2], correct [1]) for the local rows."""
    Bg, E = img_all.shape
    dev = img_all.device
    lse_i = torch.empty(Bl, device=dev, dtype=f32)
    lse_t = torch.empty(Bl, device=dev, dtype=f32)
    loss_sum = torch.zeros(2, device=dev, dtype=f32)
    correct = torch.zeros(1, device=dev, dtype=i32)
    ctx, st = _ctx_stream(img_all)
    L.check(L.load().b200clip_clip_loss_fwd(ctx, img_all.data_ptr(), txt_all.data_ptr(), logit_scale.data_ptr(), row0,
                                            Bl, Bg, E, lse_i.data_ptr(), lse_t.data_ptr(), loss_sum.data_ptr(),
                                            correct.data_ptr(), ws.data_ptr(), ws.numel(), st), "clip_loss_fwd")
    return lse_i, lse_t, loss_sum, correct


def clip_loss_bwd(img_all, txt_all, logit_scale, lse_i_all, lse_t_all, grad_out, row0, Bl, ws):
    Bg, E = img_all.shape
    dev = img_all.device
    d_img = torch.empty((Bl, E), device=dev, dtype=f32)
    d_txt = torch.empty((Bl, E), device=dev, dtype=f32)
    d_ls = torch.zeros(1, device=dev, dtype=f32)
    ctx, st = _ctx_stream(img_all)
    L.check(L.load().b200clip_clip_loss_bwd(ctx, img_all.data_ptr(), txt_all.data_ptr(), logit_scale.data_ptr(),
                                            lse_i_all.data_ptr(), lse_t_all.data_ptr(), _ptr(grad_out), row0, Bl, Bg, E,
                                            d_img.data_ptr(), d_txt.data_ptr(), d_ls.data_ptr(), ws.data_ptr(),
                                            ws.numel(), st), "clip_loss_bwd")
    return d_img, d_txt, d_ls


def adamw(master, param_bf16, grad, m, v, *, lr, beta1, beta2, eps, weight_decay, grad_scale, step, hyper=None):
    ctx, st = _ctx_stream(master)
    fn = L.load().b200clip_adamw_g16 if grad.dtype == bf16 else L.load().b200clip_adamw   # bf16: the reduce-scatter wire format
    L.check(fn(ctx, master.data_ptr(), _ptr(param_bf16), grad.data_ptr(), m.data_ptr(), v.data_ptr(), master.numel(), lr,
               beta1, beta2, eps, weight_decay, grad_scale, step, _ptr(hyper), st), "adamw")


# ------------------------------------------------------------------------------ fp32 check mode
def check_gemm_f32(a, w, *, b_major=L.MAJOR_K, bias=None, residual=None, quickgelu=False):
    """fp32 SIMT GEMM of the check mode: a fp32 [M,K]; w bf16 [N,K] (MAJOR_K) or [K,N] (MAJOR_MN)."""
    assert a.dtype == f32 and w.dtype == bf16
    M, K = a.shape
    N = w.shape[0] if b_major == L.MAJOR_K else w.shape[1]
    out = torch.empty((M, N), device=a.device, dtype=f32)
    ctx, st = _ctx_stream(a)
    L.check(L.load().b200clip_check_gemm_f32(ctx, a.data_ptr(), _row_major(a, "a"), w.data_ptr(), _row_major(w, "w"),
                                             b_major, _ptr(bias), _ptr(residual),
                                             _row_major(residual, "residual") if residual is not None else 0,
                                             out.data_ptr(), N, M, N, K, 1 if quickgelu else 0, st), "check_gemm_f32")
    return out


def check_attn_fwd_f32(qkv, B, S, H, causal):
    assert qkv.dtype == f32 and qkv.is_contiguous()
    out = torch.empty((B * S, H * 64), device=qkv.device, dtype=f32)
    ctx, st = _ctx_stream(qkv)
    L.check(L.load().b200clip_check_attn_fwd_f32(ctx, qkv.data_ptr(), out.data_ptr(), B, S, H, 1 if causal else 0, st),
            "check_attn_fwd_f32")
    return out


def check_im2col_f32(image, patch, ldcols):
    B, _, R, _ = image.shape
    g = R // patch
    cols = torch.empty((B * g * g, ldcols), device=image.device, dtype=f32)
    ctx, st = _ctx_stream(image)
    L.check(L.load().b200clip_check_im2col_f32(ctx, image.data_ptr(), cols.data_ptr(), ldcols, B, R, patch, st),
            "check_im2col_f32")
    return cols
