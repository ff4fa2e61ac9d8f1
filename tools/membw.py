"""Device-memory bandwidth probes (write-only, read-only, copy) to put the epilogue-bound GEMM shapes
in context.  usage: membw.py"""
import torch

dev = "cuda"
n = 1 << 29  # 512 Mi elements
x = torch.empty(n, device=dev, dtype=torch.bfloat16)  # 1 GiB
y = torch.empty(n, device=dev, dtype=torch.bfloat16)


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


gb = n * 2 / 1e9
s = t(lambda: x.fill_(1.0))
print(f"write-only  fill_      {gb / s:8.1f} GB/s")
s = t(lambda: x.zero_())
print(f"write-only  zero_      {gb / s:8.1f} GB/s")
s = t(lambda: y.copy_(x))
print(f"copy        copy_      {2 * gb / s:8.1f} GB/s (read + write)")
xf = x.view(torch.int32)
s = t(lambda: xf.sum())
print(f"read-only   sum        {gb / s:8.1f} GB/s")
s = t(lambda: torch.add(x, y, out=y))
print(f"2 reads + 1 write add  {3 * gb / s:8.1f} GB/s")
