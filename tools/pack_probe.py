"""Eager training-step time with the text tower packed to the caption lengths (experimental,
towers.PACK_TEXT) against the full 77 positions, same process.  usage: pack_probe.py [pairs]"""
import sys
import torch
sys.path.insert(0, ".")
from construction_clip_b200 import towers
from construction_clip_b200.model import CLIP, CONFIGS
from construction_clip_b200.train import ClipTrainer
from oracle import clip_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
torch.manual_seed(567)
m = CLIP(CONFIGS["ViT-B/32"]).to(dev)
ls = m.logit_scale.data.float().clone()
m = m.to(torch.bfloat16)
m.logit_scale.data = ls
tr = ClipTrainer(m.train())
img = O.synth_images(B, 224).to(dev)
tok = O.synth_tokens(B).to(torch.int32).to(dev)
for flag in (False, True, False, True):
    towers.PACK_TEXT = flag
    for _ in range(2):
        loss = tr.step(img, tok)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        loss = tr.step(img, tok)
    e1.record()
    torch.cuda.synchronize()
    print(f"PACK_TEXT={int(flag)}: {e0.elapsed_time(e1) / 4:7.2f} ms/step (eager), loss {float(loss):.4f}")
towers.PACK_TEXT = False
