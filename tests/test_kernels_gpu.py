"""Kernel-level parity tests (GPU): every C-ABI entry point against a plain fp32 PyTorch
statement of the same operator on the same seeded inputs.  Tolerances are bf16-level:
inputs are bf16-exact, accumulation is fp32, outputs are rounded to bf16 once."""
import math

import numpy as np
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

bf16, f32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def ops():
    from construction_clip_b200 import ops as O
    return O


def _rand(shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(bf16)


def _close(got, ref, atol, rtol, what=""):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} mismatches, max err {err.max().item():.4g} "
                           f"(ref max {ref.abs().max().item():.4g}); first bad idx {bad.nonzero()[0].tolist()}")


GEMM_SHAPES = [
    (128, 128, 64), (128, 256, 64), (128, 128, 256), (256, 512, 192), (1600, 2304, 768), (1232, 512, 2048),
    (77, 136, 72), (3000, 768, 3072), (32, 512, 768), (6400, 3072, 768),
    (20000, 768, 512), (19999, 520, 200),   # CTA-pair (cta_group::2) kernel: >= 74 tiles of 256 x 256, ragged edges
    (6400, 768, 768), (9856, 512, 1536), (6390, 776, 520),   # 128 x 192 tiles (75-78 pair tiles = 2 waves otherwise)
    # 4-CTA clusters (two pairs, B multicast): >= 8 tiles of 512 x 256 per cluster (operand ring and both TMEM stages
    # wrap several times), K not a multiple of 64, rows that leave the second pair of the last tile empty
    (51200, 768, 256), (40000 + 130, 1024, 328),
]


@pytest.fixture
def quad(monkeypatch):
    """The 4-CTA-cluster GEMM (B multicast) is off by default (measured slower); the large shapes run with it on."""
    monkeypatch.setenv("B200CLIP_GEMM_QUAD", "1")


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_nt(ops, quad, M, N, K):
    a, b = _rand((M, K), seed=1), _rand((N, K), seed=2)
    got = ops.gemm(a, b)
    ref = a.float() @ b.float().t()
    _close(got, ref, 2e-2 * math.sqrt(K / 64), 1e-2, f"gemm NT {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_nn_dgrad(ops, quad, M, N, K):
    from construction_clip_b200 import lib as L
    a, b = _rand((M, K), seed=3), _rand((K, N), seed=4)   # B stored [K, N]  (MN-major)
    got = ops.gemm(a, b, b_major=L.MAJOR_MN)
    ref = a.float() @ b.float()
    _close(got, ref, 2e-2 * math.sqrt(K / 64), 1e-2, f"gemm NN {M}x{N}x{K}")


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (768, 768, 6400), (2304, 768, 1600),
                                   (512, 2048, 9856), (136, 72, 77 * 8),
                                   (3072, 768, 6400), (1024, 512, 20000)])   # 4-CTA clusters (M a multiple of 512)
def test_gemm_tn_wgrad(ops, quad, M, N, K):
    from construction_clip_b200 import lib as L
    a, b = _rand((K, M), seed=5), _rand((K, N), seed=6)   # both stored [K, *]  (MN-major)
    out = torch.zeros((M, N), device="cuda", dtype=f32)
    ops.gemm(a, b, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, out=out, split_k=0, accumulate=True)
    ref = a.float().t() @ b.float()
    _close(out, ref, 1e-3 * math.sqrt(K), 1e-3, f"gemm TN {M}x{N}x{K}")
    # accumulate semantics: a second call adds on top
    ops.gemm(a, b, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, out=out, split_k=3, accumulate=True)
    _close(out, 2 * ref, 2e-3 * math.sqrt(K), 1e-3, f"gemm TN accumulate {M}x{N}x{K}")


@pytest.mark.parametrize("bn", [96, 128, 160, 192, 224, 256])
def test_gemm_every_tile_width(ops, bn, monkeypatch):
    """The single-CTA kernel at every tile width (B200CLIP_FORCE_BN pins it), forward (B K-major) and dgrad
    (B MN-major: 64-column swizzle atoms, the last one half used at 96 / 160 / 224), with ragged M and an N that
    is not a multiple of the width, bias + fp32 residual / QuickGELU' epilogues with the fused column sum."""
    from construction_clip_b200 import lib as L
    monkeypatch.setenv("B200CLIP_FORCE_BN", str(bn))
    monkeypatch.setenv("B200CLIP_GEMM_PAIR", "0")
    M, N, K = 1000, 808, 320
    a, w = _rand((M, K), seed=21), _rand((N, K), 0.05, seed=22)
    bias = _rand((N,), seed=23)
    auxf = torch.randn(M, N, device="cuda")
    base = a.float() @ w.float().t() + bias.float()
    _close(ops.gemm(a, w, bias=bias), base, 3e-2, 1e-2, f"bias bn={bn}")
    _close(ops.gemm(a, w, bias=bias, epilogue=L.EPI_RESIDUAL, aux=auxf, out_dtype=f32), base + auxf, 1e-3, 1e-3,
           f"residual fp32 bn={bn}")
    pre = torch.empty((M, N), device="cuda", dtype=bf16)
    got = ops.gemm(a, w, bias=bias, epilogue=L.EPI_QUICKGELU, preact=pre)
    _close(pre, base, 3e-2, 1e-2, f"preact bn={bn}")
    _close(got, base * torch.sigmoid(1.702 * base), 3e-2, 1e-2, f"quickgelu bn={bn}")
    # dgrad: dx[M, Kd] = dy[M, N] . W[N, Kd]  (W read MN-major)
    Kd = 808
    dy, wd = _rand((M, 320), seed=24), _rand((320, Kd), 0.05, seed=25)
    aux = _rand((M, Kd), seed=26)
    sg = torch.sigmoid(1.702 * aux.float())
    gref = (dy.float() @ wd.float()) * (sg * (1 + 1.702 * aux.float() * (1 - sg)))
    cs = torch.zeros(Kd, device="cuda")
    _close(ops.gemm(dy, wd, b_major=L.MAJOR_MN, epilogue=L.EPI_QUICKGELU_BWD, aux=aux, colsum=cs), gref, 3e-2, 1e-2,
           f"dgrad quickgelu_bwd bn={bn}")
    _close(cs, gref.sum(0), 0.5, 1e-2, f"dgrad colsum bn={bn}")
    # several tiles per CTA (persistent loop, both TMEM stages) at a tall shape
    M2 = 128 * 300 + 40
    a2 = _rand((M2, 128), seed=27)
    w2 = _rand((264, 128), 0.05, seed=28)
    _close(ops.gemm(a2, w2), a2.float() @ w2.float().t(), 3e-2, 1e-2, f"tall bn={bn}")


@pytest.mark.parametrize("M", [1000, 6400, 20000, -20000])   # 6400 rows -> 128 x 192 tiles, 20000 rows -> CTA-pair kernel
def test_gemm_epilogues(ops, M, monkeypatch):
    if M < 0:   # the same through the 4-CTA-cluster kernel (two pairs, B multicast)
        monkeypatch.setenv("B200CLIP_GEMM_QUAD", "1")
        M = -M
    from construction_clip_b200 import lib as L
    N, K = 768, 512
    a, w = _rand((M, K), seed=7), _rand((N, K), 0.05, seed=8)
    bias, aux = _rand((N,), seed=9), _rand((M, N), seed=10)
    base = a.float() @ w.float().t() + bias.float()
    _close(ops.gemm(a, w, bias=bias), base, 3e-2, 1e-2, "bias")
    cs0 = torch.zeros(N, device="cuda")
    _close(ops.gemm(a, w, bias=bias, colsum=cs0), base, 3e-2, 1e-2, "bias + colsum")
    _close(cs0, base.sum(0), 2e-2 * math.sqrt(M), 1e-2, "fused colsum of C (plain epilogue)")
    pre = torch.empty((M, N), device="cuda", dtype=bf16)
    got = ops.gemm(a, w, bias=bias, epilogue=L.EPI_QUICKGELU, preact=pre)
    _close(pre, base, 3e-2, 1e-2, "preact")
    _close(got, base * torch.sigmoid(1.702 * base), 3e-2, 1e-2, "quickgelu")
    _close(ops.gemm(a, w, bias=bias, epilogue=L.EPI_QUICKGELU), base * torch.sigmoid(1.702 * base), 3e-2, 1e-2,
           "quickgelu without the pre-activation output (inference)")
    nb = a.float() @ w.float().t()
    _close(ops.gemm(a, w, epilogue=L.EPI_QUICKGELU), nb * torch.sigmoid(1.702 * nb), 3e-2, 1e-2, "quickgelu, no bias")
    _close(ops.gemm(a, w, bias=bias, epilogue=L.EPI_RESIDUAL, aux=aux), base + aux.float(), 3e-2, 1e-2, "residual")
    auxf = torch.randn(M, N, device="cuda")
    _close(ops.gemm(a, w, bias=bias, epilogue=L.EPI_RESIDUAL, aux=auxf, out_dtype=f32), base + auxf, 1e-3, 1e-3,
           "residual fp32 stream")
    s = torch.sigmoid(1.702 * aux.float())
    gref = (a.float() @ w.float().t()) * (s * (1 + 1.702 * aux.float() * (1 - s)))
    cs = torch.zeros(N, device="cuda")
    _close(ops.gemm(a, w, epilogue=L.EPI_QUICKGELU_BWD, aux=aux, colsum=cs), gref, 3e-2, 1e-2, "quickgelu_bwd")
    _close(cs, gref.sum(0), 0.5, 1e-2, "fused colsum of C")
    # the 8-bit-derivative pair: the forward saves round(210 quickgelu'(x) + 22), the backward multiplies by its decode
    codes = torch.empty((M, N), device="cuda", dtype=torch.uint8)
    got = ops.gemm(a, w, bias=bias, epilogue=L.EPI_QUICKGELU_D8, preact=codes)
    _close(got, base * torch.sigmoid(1.702 * base), 3e-2, 1e-2, "quickgelu (d8)")
    sb = torch.sigmoid(1.702 * base)
    dref = sb * (1 + 1.702 * base * (1 - sb))
    dgot = (codes.float() - 22.0) / 210.0
    # half a grid step, plus the slope of the derivative times the bf16-level error of the accumulator
    assert (dgot - dref).abs().max().item() <= 0.5 / 210 + 4e-3, (dgot - dref).abs().max().item()
    assert (dgot - dref).abs().mean().item() <= 2e-3
    cs8 = torch.zeros(N, device="cuda")
    g8 = ops.gemm(a, w, epilogue=L.EPI_QUICKGELU_BWD_D8, aux=codes, colsum=cs8)
    ref8 = (a.float() @ w.float().t()) * dgot
    _close(g8, ref8, 3e-2, 1e-2, "quickgelu_bwd (d8)")
    _close(cs8, ref8.sum(0), 0.5, 1e-2, "fused colsum of C (d8)")
    sc = torch.tensor([0.37], device="cuda")
    _close(ops.gemm(a, w, scale=sc, out_dtype=f32), 0.37 * (a.float() @ w.float().t()), 1e-3, 1e-3, "scale f32")
    # strided A (row pitch > K): the CLS-row gather pattern x[:, 0, :]
    big = _rand((64, 5 * K), seed=11)
    view = big[:, :K]
    _close(ops.gemm(view, w), view.float() @ w.float().t(), 3e-2, 1e-2, "strided A")


def test_split_f32_to_bf16(ops):
    """hi + lo carries an fp32 GEMM operand at ~16 mantissa bits (the feature-projection input)."""
    x = torch.randn(1000, 768, device="cuda") * 3
    hi, lo = ops.split_f32_to_bf16(x)
    assert hi.dtype == bf16 and lo.dtype == bf16
    assert torch.equal(hi, x.to(bf16))
    rel = ((hi.float() + lo.float() - x).abs() / x.abs().clamp_min(1e-6)).max().item()
    assert rel < 2.0 ** -15, rel


@pytest.mark.parametrize("rows,d", [(1600, 768), (1232, 512), (77, 128), (515, 1024), (20000, 768), (3, 2048)])
def test_layernorm(ops, rows, d):
    x = _rand((rows, d), 2.0, seed=1)
    g = (1 + 0.1 * torch.randn(d, device="cuda")).to(bf16)
    b = (0.1 * torch.randn(d, device="cuda")).to(bf16)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, want_stats=True)
    ref = torch.nn.functional.layer_norm(x.float(), (d,), g.float(), b.float(), 1e-5)
    _close(y, ref, 2e-2, 1e-2, "ln fwd")
    _close(mean, x.float().mean(-1), 1e-5, 1e-5, "mean")
    # backward vs autograd
    xr = x.float().requires_grad_(True)
    gr, br = g.float().requires_grad_(True), b.float().requires_grad_(True)
    dy = _rand((rows, d), seed=2)
    dres = _rand((rows, d), seed=3)
    torch.nn.functional.layer_norm(xr, (d,), gr, br, 1e-5).backward(dy.float())
    dg = torch.zeros(d, device="cuda")
    db = torch.zeros(d, device="cuda")
    dcs = torch.zeros(d, device="cuda")
    dx = ops.layernorm_bwd(dy, x, g, mean, rstd, dg, db, dres=dres, dx_colsum=dcs)
    _close(dx, xr.grad + dres.float(), 3e-2, 2e-2, "ln dx")
    _close(dcs, (xr.grad + dres.float()).sum(0), 2e-2 * math.sqrt(rows), 1e-2, "ln dx colsum")
    # fp32 input (the residual stream) / fp32 output variants
    xf = torch.randn(rows, d, device="cuda") * 2
    y32, m32, r32 = ops.layernorm_fwd(xf, g, b, want_stats=True)
    _close(y32, torch.nn.functional.layer_norm(xf, (d,), g.float(), b.float(), 1e-5), 2e-2, 1e-2, "ln fwd f32 in")
    yf = ops.layernorm_fwd(x, g, b, out_dtype=f32)
    _close(yf, ref, 1e-4, 1e-4, "ln fwd f32 out")
    xfr = xf.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xfr, (d,), g.float(), b.float(), 1e-5).backward(dy.float())
    dg2, db2 = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    _close(ops.layernorm_bwd(dy, xf, g, m32, r32, dg2, db2), xfr.grad, 3e-2, 2e-2, "ln dx f32 in")
    _close(dg, gr.grad, 1e-2 * math.sqrt(rows), 1e-2, "ln dgamma")
    _close(db, br.grad, 1e-2 * math.sqrt(rows), 1e-2, "ln dbeta")


def test_layernorm_gather_and_assemble(ops):
    B, n, d = 6, 50, 768
    x = _rand((B * n, d), seed=1)
    g, b = _rand((d,), seed=2), _rand((d,), seed=3)
    idx = (torch.arange(B, device="cuda", dtype=torch.int32) * n)
    y = ops.layernorm_fwd(x, g, b, row_index=idx)
    ref = torch.nn.functional.layer_norm(x.float()[idx.long()], (d,), g.float(), b.float(), 1e-5)
    _close(y, ref, 2e-2, 1e-2, "ln gather")
    # vision token assembly: row (b,t): t==0 -> cls+pos[0] ; else patch[b,t-1] + pos[t]
    patch = _rand((B * (n - 1), d), seed=4)
    poscls = _rand((n, d), seed=5)
    cls = _rand((d,), seed=6)
    ridx = torch.full((B, n), -1, dtype=torch.int32, device="cuda")
    ridx[:, 1:] = (torch.arange(B, device="cuda")[:, None] * (n - 1) + torch.arange(n - 1, device="cuda")[None]).int()
    pre = torch.empty((B * n, d), device="cuda", dtype=bf16)
    y = ops.layernorm_fwd(patch, g, b, rows=B * n, row_index=ridx.reshape(-1).contiguous(), neg_row=cls, add=poscls,
                          add_period=n,
                          pre_out=pre)
    full = torch.cat([cls.float().expand(B, 1, d), patch.float().view(B, n - 1, d)], 1) + poscls.float()
    _close(pre, full.view(-1, d), 2e-2, 1e-2, "assemble pre")
    _close(y, torch.nn.functional.layer_norm(pre.float(), (d,), g.float(), b.float(), 1e-5), 2e-2, 1e-2, "assemble ln")


def _attn_ref(qkv, B, S, H, causal):
    d = H * 64
    q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    s = q @ k.transpose(-1, -2) / 8.0
    if causal:
        s = s + torch.full((S, S), float("-inf"), device=qkv.device).triu_(1)
    o = torch.softmax(s, -1) @ v
    return o.permute(0, 2, 1, 3).reshape(B * S, d)


@pytest.mark.parametrize("B,S,H,causal", [(2, 50, 12, False), (3, 77, 8, True), (1, 64, 2, False), (2, 16, 2, True),
                                          (1, 128, 2, True), (40, 50, 12, False), (33, 77, 8, True), (2, 7, 1, True),
                                          # >= 4 work items per persistent CTA: both input buffers of the backward in both
                                          # mbarrier phases (S <= 64), refill of the single buffer (S > 64)
                                          (128, 50, 12, False), (160, 77, 8, True), (300, 33, 5, True)])
def test_attention_fwd_bwd(ops, B, S, H, causal):
    qkv = _rand((B * S, 3 * H * 64), 1.5, seed=S)
    out, lse = ops.attn_fwd(qkv, B, S, H, causal, want_lse=True)
    qr = qkv.float().requires_grad_(True)
    ref = _attn_ref(qr, B, S, H, causal)
    _close(out, ref, 2e-2, 2e-2, "attn fwd")
    dout = _rand((B * S, H * 64), seed=S + 1)
    ref.backward(dout.float())
    dqkv = ops.attn_bwd(qkv, out, lse, dout, B, S, H, causal)
    _close(dqkv, qr.grad, 4e-2, 4e-2, "attn bwd")


@pytest.mark.parametrize("lens,S,H,causal", [([77, 5, 33, 64, 1, 76, 17], 77, 8, True), ([50, 3, 20], 50, 2, False),
                                             ([9] * 40 + [70, 2, 31], 77, 8, True),
                                             # double-buffered backward (S_max <= 64), ~4 work items per CTA
                                             ([int(x) for x in np.random.RandomState(3).randint(1, 65, 300)], 64, 4, True),
                                             ([int(x) for x in np.random.RandomState(4).randint(1, 51, 150)], 50, 8, False)])
def test_attention_varlen(ops, lens, S, H, causal):
    """Packed rows: every sample attends inside its own cu[b] .. cu[b+1]-1 rows only."""
    B, d = len(lens), H * 64
    cu = torch.tensor([0] + list(np.cumsum(lens)), device="cuda", dtype=torch.int32)
    M = int(cu[-1])
    qkv = _rand((M, 3 * d), 1.5, seed=M)
    dout = _rand((M, d), seed=M + 1)
    out, lse = ops.attn_fwd(qkv, B, S, H, causal, want_lse=True, cu=cu)
    dqkv = ops.attn_bwd(qkv, out, lse, dout, B, S, H, causal, cu=cu)
    for b, n in enumerate(lens):
        a = int(cu[b])
        qr = qkv[a:a + n].float().requires_grad_(True)
        ref = _attn_ref(qr, 1, n, H, causal)
        _close(out[a:a + n], ref, 2e-2, 2e-2, f"varlen attn fwd sample {b} (len {n})")
        ref.backward(dout[a:a + n].float())
        _close(dqkv[a:a + n], qr.grad, 4e-2, 4e-2, f"varlen attn bwd sample {b} (len {n})")


@pytest.mark.parametrize("B,S,H,causal", [(2, 197, 12, False), (1, 257, 16, False), (1, 577, 4, False), (2, 200, 2, True),
                                          (3, 129, 1, False)])
def test_attention_fwd_long(ops, B, S, H, causal):
    """KV-streaming forward (ViT-B/16: 197 tokens, ViT-L/14: 257, ViT-L/14@336px: 577)."""
    qkv = _rand((B * S, 3 * H * 64), 1.5, seed=S)
    out, lse = ops.attn_fwd(qkv, B, S, H, causal, want_lse=True)
    ref = _attn_ref(qkv, B, S, H, causal)
    _close(out, ref, 2e-2, 2e-2, "attn fwd long")
    q, k, _ = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    sc = q @ k.transpose(-1, -2) / 8.0
    if causal:
        sc = sc + torch.full((S, S), float("-inf"), device="cuda").triu_(1)
    _close(lse.view(B, H, S), torch.logsumexp(sc, -1) * 1.4426950408889634, 2e-3, 1e-3, "lse (log2)")
    # blocked two-pass backward (dK/dV pass + dQ pass)
    qr = qkv.float().requires_grad_(True)
    dout = _rand((B * S, H * 64), seed=S + 1)
    _attn_ref(qr, B, S, H, causal).backward(dout.float())
    dqkv = ops.attn_bwd(qkv, out, lse, dout, B, S, H, causal)
    _close(dqkv, qr.grad, 4e-2, 4e-2, "attn bwd long")


def test_embed_tokens(ops):
    B, S, d, V = 9, 77, 512, 49408
    g = torch.Generator(device="cuda").manual_seed(0)
    ids = torch.randint(1, 49406, (B, S), device="cuda", generator=g, dtype=torch.int32)
    lens = torch.randint(3, 77, (B,), device="cuda", generator=g)
    for b in range(B):
        ids[b, 0] = 49406
        ids[b, int(lens[b])] = 49407
        ids[b, int(lens[b]) + 1:] = 0
    table, pos = _rand((V, d), 0.02, seed=1), _rand((S, d), 0.01, seed=2)
    out, eot = ops.embed_tokens_fwd(ids, table, pos)
    ref = table.float()[ids.long()] + pos.float()[None]
    _close(out, ref.view(-1, d), 1e-3, 1e-2, "embed fwd")
    out32, _ = ops.embed_tokens_fwd(ids, table, pos, out_dtype=f32)
    _close(out32, ref.view(-1, d), 1e-7, 1e-6, "embed fwd f32")
    assert torch.equal(eot.long(), torch.arange(B, device="cuda") * S + ids.long().argmax(-1))
    dout = _rand((B * S, d), seed=3)
    dout.view(B, S, d)[0, 40:] = 0
    dt = torch.zeros((V, d), device="cuda")
    dp = torch.zeros((S, d), device="cuda")
    ops.embed_tokens_bwd(ids, dout, dt, dp)
    rt = torch.zeros((V, d), device="cuda").index_add_(0, ids.long().view(-1), dout.float())
    _close(dt, rt, 1e-3, 1e-3, "embed dtable")
    _close(dp, dout.float().view(B, S, d).sum(0), 1e-3, 1e-3, "embed dpos")


@pytest.mark.parametrize("R,p,dtype", [(224, 32, f32), (224, 16, bf16), (224, 14, f32), (64, 32, bf16)])
def test_im2col(ops, R, p, dtype):
    B = 3
    img = torch.randn(B, 3, R, R, device="cuda").to(dtype)
    cols = ops.im2col_patch(img, p)
    g = R // p
    ref = img.float().view(B, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B * g * g, 3 * p * p)
    k = 3 * p * p
    _close(cols[:, :k], ref.to(bf16), 0, 0, "im2col")
    assert (cols[:, k:] == 0).all()


def test_colsum_l2norm_cast_assemble_bwd(ops):
    x = _rand((5000, 768), seed=1)
    out = torch.zeros(768, device="cuda")
    ops.colsum(x, out)
    _close(out, x.float().sum(0), 0.5, 1e-3, "colsum")
    f = torch.randn(37, 512, device="cuda")
    y, inv = ops.l2norm_fwd(f)
    _close(y, f / f.norm(dim=1, keepdim=True), 1e-6, 1e-5, "l2norm")
    fr = f.clone().requires_grad_(True)
    dy = torch.randn(37, 512, device="cuda")
    (fr / fr.norm(dim=1, keepdim=True)).backward(dy)
    _close(ops.l2norm_bwd(dy, y, inv), fr.grad, 1e-3, 1e-2, "l2norm bwd")
    src = torch.randn(100003, device="cuda")
    _close(ops.cast_f32_to_bf16(src[:100000]), src[:100000].to(bf16), 0, 0, "cast")
    B, n, d = 7, 50, 768
    dpre = _rand((B * n, d), seed=2)
    dpos = torch.zeros((n, d), device="cuda")
    dcls = torch.zeros(d, device="cuda")
    dpatch = ops.vision_assemble_bwd(dpre, B, n, dpos, dcls)
    v = dpre.float().view(B, n, d)
    _close(dpatch, v[:, 1:].reshape(-1, d), 0, 0, "assemble dpatch")
    _close(dpos, v.sum(0), 1e-3, 1e-3, "assemble dpos")
    _close(dcls, v[:, 0].sum(0), 1e-3, 1e-3, "assemble dcls")


@pytest.mark.parametrize("Bg,E,row0,Bl", [(64, 512, 0, 64), (200, 512, 0, 200), (1024, 512, 256, 128), (96, 768, 32, 64),
                                          (9, 512, 0, 9), (27, 512, 9, 9)])   # the reference trains with batch 9
def test_clip_loss(ops, Bg, E, row0, Bl):
    torch.manual_seed(Bg)
    img = torch.nn.functional.normalize(torch.randn(Bg, E, device="cuda"), dim=1)
    txt = torch.nn.functional.normalize(img * 0.4 + torch.randn(Bg, E, device="cuda") * 0.05, dim=1)
    ls = torch.tensor([math.log(1 / 0.07)], device="cuda")
    lg = ops.logits(img, txt, ls)
    _close(lg, ls.exp() * img @ txt.t(), 1e-4, 1e-5, "logits")
    ir, tr, lr = img.clone().requires_grad_(True), txt.clone().requires_grad_(True), ls.clone().requires_grad_(True)
    L_ref = lr.exp() * ir @ tr.t()
    lab = torch.arange(Bg, device="cuda")
    loss_ref = (torch.nn.functional.cross_entropy(L_ref, lab) + torch.nn.functional.cross_entropy(L_ref.t(), lab)) / 2
    loss_ref.backward()
    ws = ops.clip_loss_workspace(img.device, Bl, Bg, E)
    # full-batch LSE vectors come from running the forward over every row block (what the
    # all-gather provides in the multi-GPU path)
    lse_i_all = torch.empty(Bg, device="cuda")
    lse_t_all = torch.empty(Bg, device="cuda")
    tot = torch.zeros(2, device="cuda")
    correct = 0
    for r0 in range(0, Bg, Bl):
        bl = min(Bl, Bg - r0)
        a, b, s, c = ops.clip_loss_fwd(img, txt, ls, r0, bl, ws)
        lse_i_all[r0:r0 + bl], lse_t_all[r0:r0 + bl] = a, b
        tot += s
        correct += int(c)
    loss = (tot[0] + tot[1]) / (2 * Bg)
    assert abs(loss.item() - loss_ref.item()) <= 1e-4 * max(1.0, abs(loss_ref.item())), (loss.item(), loss_ref.item())
    assert correct == int((L_ref.argmax(1) == lab).sum())
    d_img, d_txt, d_ls = ops.clip_loss_bwd(img, txt, ls, lse_i_all, lse_t_all, None, row0, Bl, ws)
    scale = max(ir.grad.abs().max().item(), 1e-6)
    _close(d_img, ir.grad[row0:row0 + Bl], 2e-2 * scale, 2e-2, "d_img")
    _close(d_txt, tr.grad[row0:row0 + Bl], 2e-2 * scale, 2e-2, "d_txt")
    if Bl == Bg:
        assert abs(d_ls.item() - lr.grad.item()) <= 1e-3 * max(1.0, abs(lr.grad.item())), (d_ls.item(), lr.grad.item())


def _transformers_adamw_step(p, g, m, v, step, lr, b1, b2, eps, wd):
    """Literal port of transformers.optimization.AdamW.step (correct_bias=True), the optimiser of CLIP/train.py:143."""
    m.mul_(b1).add_(g, alpha=1.0 - b1)
    v.mul_(b2).addcmul_(g, g, value=1.0 - b2)
    denom = v.sqrt().add_(eps)
    step_size = lr * math.sqrt(1.0 - b2 ** step) / (1.0 - b1 ** step)
    p.addcdiv_(m, denom, value=-step_size)
    if wd > 0.0:
        p.add_(p, alpha=-lr * wd)


def test_adamw(ops):
    """Fused AdamW == transformers.AdamW (NOT torch.optim.AdamW: eps placement differs), including tiny
    gradients (|g| ~ eps) where the two formulas disagree by up to ~4x in the first steps."""
    n = 100000
    p = torch.randn(n, device="cuda")
    g = torch.randn(n, device="cuda") * 0.1
    g[: n // 2] *= 1e-4   # |g| ~ 1e-5 against eps = 1e-6
    ref, rm, rv = p.double().clone(), torch.zeros(n, device="cuda", dtype=torch.float64), torch.zeros(n, device="cuda", dtype=torch.float64)
    master, m, v = p.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    shadow = torch.empty(n, device="cuda", dtype=bf16)
    hyper = torch.zeros(3, device="cuda")
    for step in range(1, 6):
        _transformers_adamw_step(ref, g.double(), rm, rv, step, 1e-3, 0.9, 0.98, 1e-6, 0.2)
        if step < 4:
            ops.adamw(master, shadow, g, m, v, lr=1e-3, beta1=0.9, beta2=0.98, eps=1e-6, weight_decay=0.2, grad_scale=1.0,
                      step=step)
        else:  # the CUDA-graph form: lr and bias corrections from device memory
            hyper.copy_(torch.tensor([1e-3, 1.0 - 0.9 ** step, 1.0 - 0.98 ** step]))
            ops.adamw(master, shadow, g, m, v, lr=0.0, beta1=0.9, beta2=0.98, eps=1e-6, weight_decay=0.2, grad_scale=1.0,
                      step=0, hyper=hyper)
    _close(master, ref.float(), 1e-5, 1e-5, "adamw")
    # the first update of a tiny-gradient weight must NOT be lr * sign(g) (torch.optim.AdamW's behaviour)
    _close(shadow, master.to(bf16), 0, 0, "adamw shadow")
