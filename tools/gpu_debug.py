"""Step-by-step GPU bring-up script with progressive, flushed output (debug aid)."""
import faulthandler
import sys
import time

faulthandler.enable()
faulthandler.dump_traceback_later(100, exit=True)
T0 = time.time()


def log(*a):
    print(f"[{time.time() - T0:7.2f}s]", *a, flush=True)


log("import torch")
import torch
log("torch", torch.__version__, "cuda available", torch.cuda.is_available())
x = torch.zeros(4, device="cuda")
torch.cuda.synchronize()
log("cuda init ok", torch.cuda.get_device_name(0))
sys.path.insert(0, ".")
from construction_clip_b200 import lib as L, ops as O
log("lib loaded")
h = L.ctx(0)
log("ctx created", hex(h))
bf16 = torch.bfloat16

steps = sys.argv[1:] or ["cast", "ln", "gemm1", "gemm2", "gemm3", "attn"]


def sync(what):
    torch.cuda.synchronize()
    log("   ok:", what)


for s in steps:
    log("step", s)
    if s == "cast":
        src = torch.randn(100000, device="cuda")
        d = O.cast_f32_to_bf16(src)
        sync("cast")
        log("   max err", (d.float() - src).abs().max().item())
    elif s == "ln":
        x = torch.randn(1600, 768, device="cuda").to(bf16)
        g = torch.ones(768, device="cuda", dtype=bf16)
        b = torch.zeros(768, device="cuda", dtype=bf16)
        y = O.layernorm_fwd(x, g, b)
        sync("ln")
        ref = torch.nn.functional.layer_norm(x.float(), (768,))
        log("   max err", (y.float() - ref).abs().max().item())
    elif s.startswith("gemm"):
        M, N, K = {"gemm1": (128, 128, 64), "gemm2": (128, 256, 256), "gemm3": (1600, 2304, 768)}[s]
        a = torch.randn(M, K, device="cuda").to(bf16)
        w = torch.randn(N, K, device="cuda").to(bf16)
        log("   launching NT", M, N, K)
        c = O.gemm(a, w)
        sync("gemm NT")
        ref = a.float() @ w.float().t()
        err = (c.float() - ref).abs()
        log("   NT max err", err.max().item(), "ref max", ref.abs().max().item())
        if err.max().item() > 1.0:
            log("   c[0,:8]", c[0, :8].float().tolist())
            log("   r[0,:8]", ref[0, :8].tolist())
            bad = (err > 1.0)
            log("   bad rows", bad.any(1).nonzero().flatten()[:20].tolist(), "bad cols", bad.any(0).nonzero().flatten()[:40].tolist())
        wt = w.t().contiguous()
        c = O.gemm(a, wt, b_major=L.MAJOR_MN)
        sync("gemm NN")
        err = (c.float() - ref).abs()
        log("   NN max err", err.max().item())
        at = a.t().contiguous()
        out = torch.zeros(M, N, device="cuda")
        O.gemm(at, wt, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, out=out, split_k=1, accumulate=False)
        sync("gemm TN")
        err = (out - ref).abs()
        log("   TN max err", err.max().item())
    elif s == "attn":
        for (B, S, H, causal) in [(2, 50, 12, False), (3, 77, 8, True)]:
            qkv = torch.randn(B * S, 3 * H * 64, device="cuda").to(bf16)
            log("   launching attn fwd", B, S, H, causal)
            o, lse = O.attn_fwd(qkv, B, S, H, causal, want_lse=True)
            sync("attn fwd")
            q, k, v = qkv.float().view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)
            sc = q @ k.transpose(-1, -2) / 8.0
            if causal:
                sc = sc + torch.full((S, S), float("-inf"), device="cuda").triu_(1)
            ref = (torch.softmax(sc, -1) @ v).permute(0, 2, 1, 3).reshape(B * S, H * 64)
            log("   attn fwd max err", (o.float() - ref).abs().max().item())
            do = torch.randn(B * S, H * 64, device="cuda").to(bf16)
            dq = O.attn_bwd(qkv, o, lse, do, B, S, H, causal)
            sync("attn bwd")
            log("   attn bwd absmax", dq.float().abs().max().item())
log("done")
