"""The non-GEMM kernels of one config-2 layer (and the epilogue-bound GEMM shapes), two launches each:
target for `ncu --set full` and for quick CUDA-event timing (`--time`).
usage: ncu_micro.py [--time] [--only=fc_fwd,out_fwd]"""
import sys
import torch
sys.path.insert(0, ".")
from construction_clip_b200 import lib as L, ops as O
bf16, f32 = torch.bfloat16, torch.float32
dev = "cuda"
B = 1024
cases = []


def tower(name, S, H, causal):
    d = H * 64
    M = B * S
    qkv = (torch.randn(M, 3 * d, device=dev) * 0.5).to(bf16)
    dout = (torch.randn(M, d, device=dev) * 0.1).to(bf16)
    out, lse = O.attn_fwd(qkv, B, S, H, causal, want_lse=True)
    cases.append((name + " attn_fwd", lambda: O.attn_fwd(qkv, B, S, H, causal, want_lse=True),
                  M * 4 * d * 2))
    dqkv = torch.empty_like(qkv)
    cases.append((name + " attn_bwd", lambda: O.attn_bwd(qkv, out, lse, dout, B, S, H, causal, dqkv=dqkv),
                  M * (3 + 1 + 1 + 3) * d * 2))
    x = torch.randn(M, d, device=dev)
    g = torch.randn(d, device=dev).to(bf16)
    bta = torch.randn(d, device=dev).to(bf16)
    y = torch.empty(M, d, device=dev, dtype=bf16)
    _, mean, rstd = O.layernorm_fwd(x, g, bta, out=y, want_stats=True)
    cases.append((name + " ln_fwd", lambda: O.layernorm_fwd(x, g, bta, out=y, want_stats=True), M * d * 6))
    dg = torch.zeros(d, device=dev)
    db = torch.zeros(d, device=dev)
    dc = torch.zeros(d, device=dev)
    dx = torch.empty(M, d, device=dev, dtype=bf16)
    cases.append((name + " ln_bwd", lambda: O.layernorm_bwd(dout, x, g, mean, rstd, dg, db, dres=y, dx=dx, dx_colsum=dc),
                  M * d * (2 + 4 + 2 + 2)))
    cs = torch.zeros(3 * d, device=dev)
    cases.append((name + " colsum", lambda: O.colsum(qkv, cs), M * 3 * d * 2))
    # epilogue-bound GEMMs
    xb = y
    w_fc = torch.randn(4 * d, d, device=dev).to(bf16) * 0.05
    b_fc = torch.randn(4 * d, device=dev).to(bf16)
    pre = torch.empty(M, 4 * d, device=dev, dtype=bf16)
    act = torch.empty(M, 4 * d, device=dev, dtype=bf16)
    cases.append((name + " fc_fwd", lambda: O.gemm(xb, w_fc, bias=b_fc, epilogue=L.EPI_QUICKGELU, preact=pre, out=act),
                  2 * M * 4 * d * d))
    w_pr = torch.randn(d, 4 * d, device=dev).to(bf16) * 0.05
    dfc = torch.empty(M, 4 * d, device=dev, dtype=bf16)
    cb = torch.zeros(4 * d, device=dev)
    cases.append((name + " proj_dgrad", lambda: O.gemm(dout, w_pr, b_major=L.MAJOR_MN, epilogue=L.EPI_QUICKGELU_BWD,
                                                       aux=pre, colsum=cb, out=dfc), 2 * M * 4 * d * d))
    codes = torch.empty(M, 4 * d, device=dev, dtype=torch.uint8)
    cases.append((name + " fc_fwd_d8", lambda: O.gemm(xb, w_fc, bias=b_fc, epilogue=L.EPI_QUICKGELU_D8, preact=codes, out=act),
                  2 * M * 4 * d * d))
    cases.append((name + " proj_dgrad_d8", lambda: O.gemm(dout, w_pr, b_major=L.MAJOR_MN, epilogue=L.EPI_QUICKGELU_BWD_D8,
                                                          aux=codes, colsum=cb, out=dfc), 2 * M * 4 * d * d))
    w_o = torch.randn(d, d, device=dev).to(bf16) * 0.05
    b_o = torch.randn(d, device=dev).to(bf16)
    xo = torch.empty(M, d, device=dev)
    cases.append((name + " out_fwd", lambda: O.gemm(xb, w_o, bias=b_o, epilogue=L.EPI_RESIDUAL, aux=x, out=xo),
                  2 * M * d * d))


tower("vis", 50, 12, False)
tower("txt", 77, 8, True)
timed = "--time" in sys.argv
only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
for name, fn, work in cases:
    if only and not any(o in name for o in only[0].split(",")):
        continue
    if timed:
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / n
        unit = "TFLOP/s" if "fwd" in name and "fc" in name or "dgrad" in name or "out_fwd" in name else "GB/s"
        rate = work / us * (1e-6 if unit == "TFLOP/s" else 1e-3)
        print(f"{name:16s} {us:8.1f} us  {rate:8.1f} {unit}")
    else:
        fn()
torch.cuda.synchronize()
print("ok")
