"""Where does the host-fed step (ClipTrainer.step_from_host) lose time against the device-resident step?
usage: e2e_probe.py [pairs]"""
import sys
import time
import torch
sys.path.insert(0, ".")
from construction_clip_b200.model import CLIP, CONFIGS
from construction_clip_b200.train import ClipTrainer
from oracle import clip_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda", 0)
torch.manual_seed(567)
m = CLIP(CONFIGS["ViT-B/32"]).to(dev)
ls = m.logit_scale.data.float().clone()
m = m.to(torch.bfloat16)
m.logit_scale.data = ls
tr = ClipTrainer(m.train()).enable_cuda_graph(True)
host = [(O.synth_images(B, 224, seed=s).pin_memory(), O.synth_tokens(B, seed=s).to(torch.int32).pin_memory())
        for s in range(3)]
img, tok = host[0][0].to(dev), host[0][1].to(dev)


def timed(fn, n=20):
    for _ in range(4):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    t_cpu = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_cpu


d, c = timed(lambda i: tr.step(img, tok))
print(f"device-resident step      : {d:7.3f} ms/step   (CPU enqueue {c:.3f} ms)")
d, c = timed(lambda i: tr.step_from_host(*host[i % 3], next_batch=host[(i + 1) % 3]))
print(f"step_from_host + prefetch : {d:7.3f} ms/step   (CPU enqueue {c:.3f} ms)")
cs = torch.cuda.Stream()


def h2d_only(i):
    with torch.cuda.stream(cs):
        host[i % 3][0].to(dev, non_blocking=True)


d, c = timed(h2d_only)
print(f"H2D copy alone            : {d:7.3f} ms (default-stream events), CPU {c:.3f} ms; "
      f"{host[0][0].numel() * 4 / 1e6:.0f} MB")
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10):
    with torch.cuda.stream(cs):
        host[i % 3][0].to(dev, non_blocking=True)
cs.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"H2D bandwidth             : {host[0][0].numel() * 4 / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms per batch)")

# ---- CPU time of each statement of step_from_host (same sequence, instrumented)
import collections
acc = collections.defaultdict(float)
copy_stream = torch.cuda.Stream()
loss_host = torch.zeros(64, dtype=torch.float32).pin_memory()
cur = torch.cuda.current_stream(dev)


def tick(name, t0):
    t1 = time.perf_counter()
    acc[name] += t1 - t0
    return t1


staged = None
N = 20
torch.cuda.synchronize()
for i in range(N + 2):
    if i == 2:
        acc.clear()
    t = time.perf_counter()
    img_h, txt_h = host[i % 3]
    if staged is None:
        with torch.cuda.stream(copy_stream):
            staged = (img_h.to(dev, non_blocking=True), txt_h.to(dev, non_blocking=True), torch.cuda.Event())
            staged[2].record(copy_stream)
    img_d, txt_d, ev = staged
    t = tick("unpack", t)
    cur.wait_event(ev)
    img_d.record_stream(cur)
    txt_d.record_stream(cur)
    t = tick("wait_event+record_stream", t)
    nh = host[(i + 1) % 3]
    with torch.cuda.stream(copy_stream):
        a = nh[0].to(dev, non_blocking=True)
        t = tick("H2D image .to()", t)
        b = nh[1].to(dev, non_blocking=True)
        t = tick("H2D tokens .to()", t)
        e = torch.cuda.Event()
        e.record(copy_stream)
    staged = (a, b, e)
    t = tick("event record", t)
    loss = tr.step(img_d, txt_d)
    t = tick("tr.step (graph replay)", t)
    loss_host[i % 64:i % 64 + 1].copy_(loss.reshape(1), non_blocking=True)
    t = tick("D2H loss", t)
    del img_d, txt_d
    t = tick("free", t)
torch.cuda.synchronize()
for k, v in acc.items():
    print(f"  {k:28s} {v / N * 1e3:8.3f} ms / step")
