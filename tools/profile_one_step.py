"""One eager training step (ViT-B/32, 1024 pairs, one GPU) bracketed by cudaProfilerStart/Stop: the
target of the ncu passes in tools/collect_profiles.sh (`ncu --profile-from-start off ...`).
Two warm-up steps run outside the profiled range.  usage: profile_one_step.py [pairs]"""
import sys
import torch
sys.path.insert(0, ".")
from construction_clip_b200 import lib as L
from construction_clip_b200.model import CLIP, CONFIGS
from construction_clip_b200.train import ClipTrainer
from oracle import clip_oracle as O  # synthetic inputs only

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
torch.manual_seed(567)
m = CLIP(CONFIGS["ViT-B/32"]).to(dev)
ls = m.logit_scale.data.float().clone()
m = m.to(torch.bfloat16)
m.logit_scale.data = ls
tr = ClipTrainer(m.train())
tr.two_streams = False  # serialised launch order = the order of the list
img = O.synth_images(B, 224).to(dev)
tok = O.synth_tokens(B).to(torch.int32).to(dev)
for _ in range(2):
    tr.step(img, tok)
torch.cuda.synchronize()
n0 = L.launch_count()
torch.cuda.profiler.start()
loss = tr.step(img, tok)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"ok loss={float(loss):.4f} library launches in the profiled step: {L.launch_count() - n0}")
