"""CUPTI timeline of one graph-replayed training step on rank 0 of an N-rank run (torchrun): where does the step
go when the collectives overlap the backward?  Prints span, per-kernel-family busy time, every NCCL kernel with
its start offset / duration, and the idle gaps of the compute streams.  usage (under torchrun): dist_profile.py"""
import collections
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from construction_clip_b200.model import CLIP, CONFIGS
    from construction_clip_b200.train import ClipTrainer
    from oracle import clip_oracle as ORC
    bl = int(os.environ.get("PAIRS", 1024)) // world
    torch.manual_seed(567)
    m = CLIP(CONFIGS["ViT-B/32"]).to(dev)
    ls = m.logit_scale.data.float().clone()
    m = m.to(torch.bfloat16)
    m.logit_scale.data = ls
    if world > 1:
        for p in m.parameters():
            dist.broadcast(p.data, 0)
    tr = ClipTrainer(m.train(), lr=1e-5)
    tr.enable_cuda_graph()
    img = ORC.synth_images(bl, 224, seed=567 + rank).to(dev)
    tok = ORC.synth_tokens(bl, seed=567 + rank).to(torch.int32).to(dev)
    for _ in range(5):
        tr.step(img, tok)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        tr.step(img, tok)
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 10
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            tr.step(img, tok)
        torch.cuda.synchronize()
    if rank == 0:
        ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                    key=lambda e: e.time_range.start)
        t0, t1 = ev[0].time_range.start, ev[-1].time_range.end
        half = (t0 + t1) / 2
        ev = [e for e in ev if e.time_range.start >= half - 1]     # the second step only
        t0, t1 = ev[0].time_range.start, max(e.time_range.end for e in ev)
        print(f"world {world}, {bl} pairs/GPU: step {step_ms:.3f} ms (events); profiled span {(t1 - t0) / 1e3:.3f} ms, {len(ev)} activities")
        fam = collections.defaultdict(lambda: [0, 0.0])
        for e in ev:
            n = e.name
            k = ("nccl" if "nccl" in n.lower() else "gemm" if "gemm" in n else "attn" if "attn" in n else
                 "layernorm" if "layernorm" in n else "adamw" if "adamw" in n else "memset/memcpy" if "Mem" in n else "other")
            fam[k][0] += 1
            fam[k][1] += e.time_range.end - e.time_range.start
        for k, (c, t) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
            print(f"  {k:14s} n={c:4d} sum {t / 1e3:8.3f} ms")
        # union of busy intervals of all non-NCCL kernels (how much of the span has NO compute kernel running)
        iv = sorted((e.time_range.start, e.time_range.end) for e in ev if "nccl" not in e.name.lower())
        busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
        gaps = []
        for s, e_ in iv[1:]:
            if s > cur_e:
                busy += cur_e - cur_s
                gaps.append((s - cur_e, cur_e - t0))
                cur_s, cur_e = s, e_
            else:
                cur_e = max(cur_e, e_)
        busy += cur_e - cur_s
        print(f"  compute-kernel union busy {busy / 1e3:.3f} ms of span {(t1 - t0) / 1e3:.3f} ms; largest gaps (us @ offset ms): "
              + ", ".join(f"{g:.0f}@{o / 1e3:.2f}" for g, o in sorted(gaps, reverse=True)[:8]))
        print("  NCCL kernels (offset ms, duration us):")
        for e in ev:
            if "nccl" in e.name.lower():
                print(f"    {(e.time_range.start - t0) / 1e3:7.3f}  {e.time_range.end - e.time_range.start:8.1f}  {e.name[:70]}")
    tr.enable_cuda_graph(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
