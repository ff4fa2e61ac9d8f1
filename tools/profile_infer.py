"""One no-grad forward of BASELINE config 3 (ViT-B/16, 1024 images) or 4 (ViT-L/14, 512 pairs) bracketed by
cudaProfilerStart/Stop: target of `ncu --profile-from-start off`.  usage: profile_infer.py <3|4> [batch]"""
import sys
import torch
sys.path.insert(0, ".")
from construction_clip_b200.model import CLIP, CONFIGS
from oracle import clip_oracle as O  # synthetic inputs only

cfgno = int(sys.argv[1]) if len(sys.argv) > 1 else 4
name = {3: "ViT-B/16", 4: "ViT-L/14"}[cfgno]
B = int(sys.argv[2]) if len(sys.argv) > 2 else (1024 if cfgno == 3 else 512)
dev = torch.device("cuda", 0)
torch.manual_seed(567)
m = CLIP(CONFIGS[name]).to(dev)
ls = m.logit_scale.data.float().clone()
m = m.to(torch.bfloat16)
m.logit_scale.data = ls
m.eval()
img = O.synth_images(B, 224).to(dev).to(torch.bfloat16)
tok = O.synth_tokens(B).to(torch.int32).to(dev)


def fwd():
    with torch.no_grad():
        a = m.encode_image(img)
        b = m.encode_text(tok) if cfgno == 4 else None
    return a, b


for _ in range(2):
    fwd()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
fwd()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
torch.cuda.profiler.start()
fwd()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"ok config {cfgno} {name} batch {B}: {ms:.2f} ms per forward = {B / ms * 1e3:.0f} /s")
