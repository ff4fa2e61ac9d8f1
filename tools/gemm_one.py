"""One GEMM shape, a few launches (for ncu --set full).  usage: gemm_one.py case [rows]"""
import sys
import torch
sys.path.insert(0, ".")
from construction_clip_b200 import lib as L, ops as O
bf16, f32 = torch.bfloat16, torch.float32
case = sys.argv[1]
dev = "cuda"
M, d = (78848, 512) if case.startswith("txt") else (51200, 768)
if len(sys.argv) > 2:
    M = int(sys.argv[2])
x = torch.randn(M, d, device=dev).to(bf16)
x4 = torch.randn(M, 4 * d, device=dev).to(bf16)
xf = torch.randn(M, d, device=dev)
if case.endswith("fc_fwd"):
    w = torch.randn(4 * d, d, device=dev).to(bf16); b = torch.randn(4 * d, device=dev).to(bf16)
    pre = torch.empty(M, 4 * d, device=dev, dtype=bf16)
    fn = lambda: O.gemm(x, w, bias=b, epilogue=L.EPI_QUICKGELU, preact=pre)
elif case.endswith("out_fwd"):
    w = torch.randn(d, d, device=dev).to(bf16); b = torch.randn(d, device=dev).to(bf16)
    fn = lambda: O.gemm(x, w, bias=b, epilogue=L.EPI_RESIDUAL, aux=xf, out_dtype=f32)
elif case.endswith("fc_dgrad"):
    w = torch.randn(4 * d, d, device=dev).to(bf16)
    fn = lambda: O.gemm(x4, w, b_major=L.MAJOR_MN)
elif case.endswith("proj_dgrad"):
    w = torch.randn(d, 4 * d, device=dev).to(bf16)
    fn = lambda: O.gemm(x, w, b_major=L.MAJOR_MN, epilogue=L.EPI_QUICKGELU_BWD, aux=x4)
elif case.endswith("fc_wgrad"):
    g = torch.zeros(4 * d, d, device=dev)
    fn = lambda: O.gemm(x4, x, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, out=g, split_k=0, accumulate=True)
elif case.endswith("qkv_fwd"):
    w = torch.randn(3 * d, d, device=dev).to(bf16); b = torch.randn(3 * d, device=dev).to(bf16)
    fn = lambda: O.gemm(x, w, bias=b)
for _ in range(4):
    fn()
torch.cuda.synchronize()
print("ok")
