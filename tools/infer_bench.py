"""Forward-only throughput of BASELINE configs 3 and 4 (not bench lines: parity-test configs whose
speed is recorded for DESIGN.md).  usage: infer_bench.py"""
import sys
import torch
sys.path.insert(0, ".")
from construction_clip_b200.model import CLIP, CONFIGS
from oracle import clip_oracle as O

dev = torch.device("cuda", 0)


def build(name):
    torch.manual_seed(567)
    m = CLIP(CONFIGS[name]).to(dev)
    ls = m.logit_scale.data.float().clone()
    m = m.to(torch.bfloat16)
    m.logit_scale.data = ls
    return m.eval()


def timeit(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


with torch.no_grad():
    # config 3: ViT-B/16 embedding extraction, 4096 images per GPU (in chunks of 1024)
    m = build("ViT-B/16")
    img = torch.randn(1024, 3, 224, 224, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: [m.encode_image(img) for _ in range(4)])
    f = O.flops_image(O.CONFIGS["ViT-B/16"])
    print(f"config3 ViT-B/16 encode_image: 4096 images in {ms:.1f} ms = {4096 / ms * 1e3:.0f} img/s, "
          f"{4096 / ms * 1e3 * f / 1e12:.0f} TFLOP/s")
    del m
    # config 4: ViT-L/14 encode_image + encode_text, batch 512
    m = build("ViT-L/14")
    img = torch.randn(512, 3, 224, 224, device=dev, dtype=torch.bfloat16)
    tok = O.synth_tokens(512).to(dev)
    ms = timeit(lambda: (m.encode_image(img), m.encode_text(tok)))
    f = O.flops_pair(O.CONFIGS["ViT-L/14"])
    print(f"config4 ViT-L/14 encode_image+encode_text: 512 pairs in {ms:.1f} ms = {512 / ms * 1e3:.0f} pairs/s, "
          f"{512 / ms * 1e3 * f / 1e12:.0f} TFLOP/s")
    del m
    # config 1: ViT-B/32 zero-shot predict 32 x 16
    m = build("ViT-B/32")
    img = torch.randn(32, 3, 224, 224, device=dev)
    tok = O.synth_tokens(16, max_len=12).to(dev)
    ms = timeit(lambda: m(img, tok), iters=20)
    print(f"config1 ViT-B/32 model(image[32], text[16]): {ms:.3f} ms per call")
