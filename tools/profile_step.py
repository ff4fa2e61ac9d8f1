"""Warm per-kernel timeline of one training step with torch.profiler (CUPTI): real (not
serialised / cold-cache) durations, GPU busy time vs wall time.  usage: profile_step.py [pairs]"""
import sys
import collections
import torch
sys.path.insert(0, ".")
from construction_clip_b200.model import CLIP, CONFIGS
from construction_clip_b200.train import ClipTrainer
from oracle import clip_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
cfg = CONFIGS["ViT-B/32"]
torch.manual_seed(567)
m = CLIP(cfg).to(dev)
ls = m.logit_scale.data.float().clone()
m = m.to(torch.bfloat16)
m.logit_scale.data = ls
tr = ClipTrainer(m.train())
img = O.synth_images(B, 224).to(dev)
tok = O.synth_tokens(B).to(torch.int32).to(dev)
for _ in range(3):
    tr.step(img, tok)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    tr.step(img, tok)
e1.record()
torch.cuda.synchronize()
print(f"wall per step (eager, no profiler): {e0.elapsed_time(e1) / 3:.2f} ms")
tr.enable_cuda_graph(True)
for _ in range(3):
    tr.step(img, tok)
torch.cuda.synchronize()
e0.record()
for _ in range(10):
    tr.step(img, tok)
e1.record()
torch.cuda.synchronize()
print(f"wall per step (CUDA graph replay): {e0.elapsed_time(e1) / 10:.2f} ms")
tr.enable_cuda_graph(False)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    tr.step(img, tok)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
agg = collections.defaultdict(lambda: [0, 0.0])
t0 = min(e.time_range.start for e in ev)
t1 = max(e.time_range.end for e in ev)
busy = 0.0
for e in ev:
    d = e.time_range.end - e.time_range.start
    name = e.name.split("(")[0].replace("void ", "")[:80]
    agg[name][0] += 1
    agg[name][1] += d
    busy += d
print(f"GPU span {(t1 - t0) / 1e3:.2f} ms, sum of kernel durations {busy / 1e3:.2f} ms, {len(ev)} device activities")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{t / 1e3:8.3f} ms {100 * t / busy:5.1f}% n={n:4d} avg={t / n:8.1f} us  {k}")
