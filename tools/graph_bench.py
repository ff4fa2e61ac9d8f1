"""Where does a small-batch step go?  Times every kernel of one ViT-B/32 layer (both towers) the way the
training step runs it -- back to back inside a CUDA-graph replay, warm caches, no host in the loop -- and
compares the sum with the measured graph-replayed step.  usage: graph_bench.py [pairs] [txt_rows]"""
import sys
import torch
sys.path.insert(0, ".")
from construction_clip_b200 import lib as L, ops as O

bf16, f32, i32 = torch.bfloat16, torch.float32, torch.int32
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
_g = max(256, -(-(B * 77 // 32) // 256) * 256)
TXT_ROWS = int(sys.argv[2]) if len(sys.argv) > 2 else -(-int(B * 41.5) // _g) * _g
REPS = 24


def graph_time(fn, reps=REPS, iters=20):
    """us per call of fn when `reps` calls are captured in one graph and replayed."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (iters * reps)


rows = []
total = {"gemm": 0.0, "other": 0.0}
for tower, M, d, S, causal in (("vis", B * 50, 768, 50, False), ("txt", TXT_ROWS, 512, 77, True)):
    H = d // 64
    x = torch.randn(M, d, device=dev).to(bf16)
    x4 = torch.randn(M, 4 * d, device=dev).to(bf16)
    x3 = torch.randn(M, 3 * d, device=dev).to(bf16)
    xf = torch.randn(M, d, device=dev)
    gam = torch.ones(d, device=dev, dtype=bf16)
    for name, N, K in (("qkv", 3 * d, d), ("out", d, d), ("fc", 4 * d, d), ("proj", d, 4 * d)):
        w = torch.randn(N, K, device=dev).to(bf16)
        bias = torch.randn(N, device=dev).to(bf16)
        a = x if K == d else x4
        dy = {d: x, 3 * d: x3, 4 * d: x4}[N]
        fl = 2.0 * M * N * K
        out_bf = torch.empty(M, N, device=dev, dtype=bf16)
        out_f = torch.empty(M, N, device=dev, dtype=f32)
        if name == "fc":
            pre = torch.empty(M, N, device=dev, dtype=bf16)
            t = graph_time(lambda: O.gemm(a, w, bias=bias, epilogue=L.EPI_QUICKGELU, preact=pre, out=out_bf))
        elif name in ("out", "proj"):
            t = graph_time(lambda: O.gemm(a, w, bias=bias, epilogue=L.EPI_RESIDUAL, aux=xf, out=out_f))
        else:
            t = graph_time(lambda: O.gemm(a, w, bias=bias, out=out_bf))
        rows.append((tower, name + " fwd", t, fl))
        dx = torch.empty(M, K, device=dev, dtype=bf16)
        if name == "proj":
            t = graph_time(lambda: O.gemm(dy, w, b_major=L.MAJOR_MN, epilogue=L.EPI_QUICKGELU_BWD, aux=x4, out=dx))
        else:
            t = graph_time(lambda: O.gemm(dy, w, b_major=L.MAJOR_MN, out=dx))
        rows.append((tower, name + " dgrad", t, fl))
        g = torch.zeros(N, K, device=dev)
        t = graph_time(lambda: O.gemm(dy, a, a_major=L.MAJOR_MN, b_major=L.MAJOR_MN, out=g, split_k=0, accumulate=True))
        rows.append((tower, name + " wgrad", t, fl))
    # the non-GEMM kernels of the layer
    mean = torch.zeros(M, device=dev)
    rstd = torch.ones(M, device=dev)
    dg, db, cs = torch.zeros(d, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)
    h = torch.empty(M, d, device=dev, dtype=bf16)
    t = graph_time(lambda: O.layernorm_fwd(xf, gam, gam, out=h))
    rows.append((tower, "ln fwd (x2)", 2 * t, 0))
    dxo = torch.empty(M, d, device=dev, dtype=bf16)
    t = graph_time(lambda: O.layernorm_bwd(x, xf, gam, mean, rstd, dg, db, dres=x, dx=dxo, dx_colsum=cs))
    rows.append((tower, "ln bwd (x2)", 2 * t, 0))
    if tower == "vis":
        qkv = torch.randn(M, 3 * d, device=dev).to(bf16)
        o = torch.empty(M, d, device=dev, dtype=bf16)
        lse = torch.empty(M * H, device=dev)
        dq = torch.empty_like(qkv)
        t = graph_time(lambda: O.attn_fwd(qkv, B, S, H, causal, out=o))
        rows.append((tower, "attn fwd", t, 0))
        o, lse = O.attn_fwd(qkv, B, S, H, causal, want_lse=True)
        t = graph_time(lambda: O.attn_bwd(qkv, o, lse, x, B, S, H, causal, dqkv=dq))
        rows.append((tower, "attn bwd", t, 0))
    else:
        lens = torch.randint(4, 78, (B,), device=dev)
        lens = (lens.float() * (M - 64) / lens.sum().item()).long().clamp(4, 77)
        cu = torch.zeros(B + 1, device=dev, dtype=i32)
        cu[1:] = torch.cumsum(lens, 0).to(i32)
        cu.clamp_(max=M)
        qkv = torch.randn(M, 3 * d, device=dev).to(bf16)
        o = torch.empty(M, d, device=dev, dtype=bf16)
        dq = torch.empty_like(qkv)
        t = graph_time(lambda: O.attn_fwd(qkv, B, S, H, causal, out=o, cu=cu))
        rows.append((tower, "attn fwd", t, 0))
        o, lse = O.attn_fwd(qkv, B, S, H, causal, want_lse=True, cu=cu)
        t = graph_time(lambda: O.attn_bwd(qkv, o, lse, x, B, S, H, causal, dqkv=dq, cu=cu))
        rows.append((tower, "attn bwd", t, 0))
    t = graph_time(lambda: O.colsum(x, cs))
    rows.append((tower, "colsum", t, 0))

for tower, name, t, fl in rows:
    kind = "gemm" if fl else "other"
    total[kind] += t
    extra = f"{fl / t / 1e6:8.1f} TFLOP/s" if fl else ""
    print(f"{tower:4s} {name:12s} {t:8.1f} us {extra}")
print(f"per layer (both towers): GEMM {total['gemm']:.1f} us + other {total['other']:.1f} us; x12 layers = "
      f"{12 * (total['gemm'] + total['other']) / 1e3:.2f} ms (GEMM {12 * total['gemm'] / 1e3:.2f} ms)")

# the real step, graph replayed, two streams / one stream
from construction_clip_b200.model import CLIP, CONFIGS
from construction_clip_b200.train import ClipTrainer
from oracle import clip_oracle as ORC
torch.manual_seed(567)
m = CLIP(CONFIGS["ViT-B/32"]).to(dev)
ls = m.logit_scale.data.float().clone()
m = m.to(bf16)
m.logit_scale.data = ls
img = ORC.synth_images(B, 224).to(dev)
tok = ORC.synth_tokens(B).to(i32).to(dev)
for two in (True, False):
    tr = ClipTrainer(m.train())
    tr.two_streams = two
    tr.enable_cuda_graph()
    for _ in range(4):
        tr.step(img, tok)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tr.step(img, tok)
    e1.record()
    torch.cuda.synchronize()
    print(f"step (graph replay, {'two streams' if two else 'one stream'}): {e0.elapsed_time(e1) / 20:.3f} ms, rows {tr.text_rows(tok)}")
    tr.enable_cuda_graph(False)
    del tr
