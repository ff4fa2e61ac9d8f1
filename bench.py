#!/usr/bin/env python
"""Headline benchmark: CLIP ViT-B/32 contrastive fine-tune step (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model ViT-B/32]

One "step" = CLIP/train.py:157-171 on one synthetic batch: zero grads, both towers forward,
all-gathered symmetric InfoNCE, full backward, gradient all-reduce (N > 1), AdamW.  Global batch
1024 image-text pairs (strong scaling: 1024 / N pairs per rank), bf16 tensor-core arithmetic with
fp32 accumulation, synthetic 224x224 images and 77-token prompts, random-init weights (seed 567).

Printed JSON (one line, rank 0): `value` = pairs/s with inputs resident in HBM; `e2e` = the same
metric through ClipTrainer.step_from_host with HOST (pinned) inputs, the host->device copy of every
step's batch and a device->host read of every step's loss inside the timed region; `roofline` =
the tcgen05 GEMM kernel family timed live with CUDA events inside the timed region;
`cpu_baseline` = the CPU oracle (restated upstream CLIP, fp32, all host threads) on a bounded
sample.  `--impl reference` times that CPU implementation alone (the reference's own code path
for this metric is `clip` on the host cores).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner
# on stdout when the box sets NCCL_DEBUG), so file descriptor 1 is pointed at stderr for the whole
# run and the result line goes to a private duplicate of the original stdout.
_RESULT_OUT = None


def _claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()

GLOBAL_BATCH = int(os.environ.get("BENCH_GLOBAL_BATCH", "1024"))   # 1024 = BASELINE configs[1]; other values are tuning runs, not the headline
SEED = 567
CPU_SAMPLE_PAIRS = 32


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"bf16_burst": p.get("bf16_tflops", 1590.0), "bf16_sustained": p.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": p.get("hbm_gbs", 6650.0), "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(steps, warmup, pairs=CPU_SAMPLE_PAIRS, model_name="ViT-B/32"):
    """Oracle (fp32 PyTorch restatement of upstream CLIP) train step on the host cores."""
    import torch
    from oracle import clip_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    model = O.build(model_name, seed=SEED).train()
    cfg = O.CONFIGS[model_name]
    img = O.synth_images(pairs, cfg.image_resolution, seed=SEED)
    tok = O.synth_tokens(pairs, seed=SEED)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, eps=1e-6, weight_decay=0.0)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        model.zero_grad()
        lpi, lpt = model(img, tok)
        loss = O.clip_loss(lpi, lpt)
        loss.backward()
        opt.step()
        loss.item()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return pairs / sec, sec, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
    rate, sec, cores = cpu_oracle_rate(steps, warmup, model_name=args.model)
    line = {
        "impl": "reference", "metric": "image-text pairs/sec (contrastive fine-tune step)", "value": rate,
        "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"CLIP {args.model} contrastive fine-tune step (CLIP/train.py:157-171), CPU oracle, "
                               f"{CPU_SAMPLE_PAIRS}-pair bounded sample of the 1024-pair global batch",
                   "global_batch": CPU_SAMPLE_PAIRS, "sample_of_global_batch": GLOBAL_BATCH, "parallelism": "cpu"},
        "cpu_baseline": {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{CPU_SAMPLE_PAIRS} pairs/step, fwd+loss+bwd+AdamW, fp32, median of {steps}"},
        "e2e": {"value": rate, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def torch_eager_bf16_rate(model_name, pairs, steps, warmup, dev):
    """The GPU-side bar: the restated upstream module run by PyTorch eager in bf16 on the same
    B200 -- cuBLAS GEMMs, nn.MultiheadAttention's SDPA path, ATen LayerNorm / softmax / CE, fused torch AdamW --
    i.e. what the reference's CLIP/train.py:157-171 executes (it ships no kernel of its own), bf16 for fp16.
    Secondary figure, taken after every headline number; none of this repository's kernels run here."""
    import torch
    from oracle import clip_oracle as ORC
    cfg = ORC.CONFIGS[model_name]
    model = ORC.build(model_name, seed=SEED).to(dev).train()

    def _convert(l):   # upstream clip.model.convert_weights with bf16 for fp16: LayerNorm and embeddings stay fp32
        if isinstance(l, (torch.nn.Conv2d, torch.nn.Linear)):
            l.weight.data = l.weight.data.to(torch.bfloat16)
            if l.bias is not None:
                l.bias.data = l.bias.data.to(torch.bfloat16)
        if isinstance(l, torch.nn.MultiheadAttention):
            for attr in ("in_proj_weight", "in_proj_bias"):
                t = getattr(l, attr)
                if t is not None:
                    t.data = t.data.to(torch.bfloat16)
        for name in ("text_projection", "proj"):
            if hasattr(l, name) and getattr(l, name) is not None:
                getattr(l, name).data = getattr(l, name).data.to(torch.bfloat16)

    model.apply(_convert)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-5, eps=1e-6, weight_decay=0.0, fused=True)
    img = ORC.synth_images(pairs, cfg.image_resolution, seed=SEED).to(dev).to(torch.bfloat16)
    tok = ORC.synth_tokens(pairs, seed=SEED).to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        lpi, lpt = model(img, tok)
        loss = ORC.clip_loss(lpi, lpt)
        loss.backward()
        opt.step()
        return loss

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    out = {"value": pairs / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "pairs_per_step": pairs,
           "loss": float(loss.item()), "torch": torch.__version__,
           "what": "PyTorch eager bf16 of the restated upstream module (cuBLAS + SDPA + ATen + fused AdamW), same step, "
                   "same B200, device-resident inputs; not this repository's kernels"}
    del model, opt, img, tok
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CLIP hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from construction_clip_b200 import lib as L, ops as O
    from construction_clip_b200.model import CLIP, CONFIGS
    from construction_clip_b200.train import ClipTrainer
    from oracle import clip_oracle as ORC  # FLOP model + synthetic-input recipe only (never on the timed path)

    cfg = CONFIGS[args.model]
    assert GLOBAL_BATCH % world == 0
    bl = GLOBAL_BATCH // world
    torch.manual_seed(SEED)
    model = CLIP(cfg).to(dev)
    ls = model.logit_scale.data.float().clone()
    model = model.to(torch.bfloat16)
    model.logit_scale.data = ls
    model.train()
    if world > 1:  # identical replicas
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    trainer = ClipTrainer(model, lr=1e-5, warmup_steps=5000, total_steps=100000)
    # The step is replayed as ONE CUDA graph (two tower streams + NCCL collectives captured): ~500
    # launches / step of host work disappear, which matters as soon as the per-GPU batch is small.
    # B200CLIP_GRAPH=0 forces eager launches.
    use_graph = os.environ.get("B200CLIP_GRAPH", "1") == "1"
    if use_graph:
        trainer.enable_cuda_graph()

    # synthetic inputs: two alternating host batches, pinned
    n_host = 2
    host = []
    for i in range(n_host):
        img = ORC.synth_images(bl, cfg.image_resolution, seed=SEED + 17 * rank + 1000 * i)
        tok = ORC.synth_tokens(bl, seed=SEED + 17 * rank + 1000 * i).to(torch.int32)
        host.append((img.pin_memory(), tok.pin_memory()))
    dev_batches = [(im.to(dev), tk.to(dev)) for im, tk in host]
    h2d_bytes = host[0][0].numel() * 4 + host[0][1].numel() * 4
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident throughput (`value`) ----------------
    for i in range(args.warmup):
        trainer.step(*dev_batches[i % n_host])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for i in range(args.steps):
        loss = trainer.step(*dev_batches[i % n_host])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = L.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    final_loss = float(loss.item())
    ms_per_step = ms_total / args.steps
    value = GLOBAL_BATCH * args.steps / (ms_total * 1e-3)

    # Roofline of the dominant kernel family (the tcgen05 GEMM).  In the timed region the two towers run
    # on two streams (and, for small per-GPU batches, inside a CUDA-graph replay), so kernels overlap
    # and cannot be bracketed one by one; their durations are therefore taken with CUDA events around
    # every GEMM launch in serialised (single-stream, eager) steps run right after the timed region.
    trainer_graph_ok = trainer._use_graph and len(trainer._graphs) > 0   # False: capture failed, eager fallback ran
    graphs_captured = len(trainer._graphs)
    from construction_clip_b200 import towers as TW
    text_rows = [trainer.text_rows(tk) for _, tk in dev_batches]   # static rows of the packed text tower per batch
    real_rows = [int((tk.argmax(-1) + 1).sum().item()) for _, tk in dev_batches]
    trainer.enable_cuda_graph(False)
    trainer.two_streams = False
    n0 = L.launch_count()
    trainer.step(*dev_batches[0])            # settle
    torch.cuda.synchronize()
    n_instr = 2
    O.GEMM_PROFILE = []
    i0, i1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i0.record()
    for i in range(n_instr):
        trainer.step(*dev_batches[i % n_host])
    i1.record()
    torch.cuda.synchronize()
    gemm_prof, O.GEMM_PROFILE = O.GEMM_PROFILE, None
    instr_ms = i0.elapsed_time(i1)
    if trainer_graph_ok:
        launches = (L.launch_count() - n0) // (n_instr + 1) * args.steps   # launches one replay contains x steps
    graph_captured = bool(trainer_graph_ok)
    trainer.two_streams = True
    trainer.enable_cuda_graph(use_graph)
    gemm_ms = sum(a.elapsed_time(b) for a, b, _, _ in gemm_prof)
    gemm_flops = sum(f for _, _, f, _ in gemm_prof)
    exec_flops_step = gemm_flops / n_instr   # executed GEMM FLOPs of one step on this rank
    gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    gemm_alg_bytes = sum(info[5] for _, _, _, info in gemm_prof) / max(1, len(gemm_prof))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_gemm_traffic.json")   # ncu capture of the current step (tools/summarize_dram.py)
    if not os.path.exists(tpath):
        tpath = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    if world == 1 and os.path.exists(tpath):   # ncu capture of the same workload (N = 1), committed under profiles/
        with open(tpath) as fh:
            traffic = json.load(fh).get("dram_bytes_per_launch")

    # ---------------- end to end through the host-facing API (`e2e`) ----------------
    # Host batches in the wire format of a loader that leaves ToTensor + Normalize to the GPU (uint8 [B,3,R,R]
    # pixels, construction_clip_b200.data.preprocess_uint8; the kernels normalise inside the patch-embedding
    # im2col) -- a quarter of the bytes of the fp32 tensors upstream's `preprocess` produces.  The fp32 form
    # (what the reference's DataLoader yields today, CLIP/train.py:56,159) is timed as well: `e2e_fp32_host`.
    def e2e_leg(batches):
        for i in range(max(2, args.warmup)):
            trainer.step_from_host(*batches[i % n_host], next_batch=batches[(i + 1) % n_host])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            slot = trainer.step_from_host(*batches[i % n_host], next_batch=batches[(i + 1) % n_host])
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        _ = float(slot.item())
        return ms

    host_u8 = []
    for i in range(n_host):
        g8 = torch.Generator().manual_seed(SEED + 17 * rank + 1000 * i)
        host_u8.append((torch.randint(0, 256, (bl, 3, cfg.image_resolution, cfg.image_resolution), generator=g8,
                                      dtype=torch.uint8).pin_memory(), host[i][1]))
    e2e_ms = e2e_leg(host_u8)
    e2e_value = GLOBAL_BATCH * args.steps / (e2e_ms * 1e-3)
    h2d_u8 = host_u8[0][0].numel() + host_u8[0][1].numel() * 4
    e2e32_ms = e2e_leg(host)
    e2e32_value = GLOBAL_BATCH * args.steps / (e2e32_ms * 1e-3)

    def shutdown():
        # Drop captured graphs (they hold NCCL kernels) BEFORE the communicator goes away, and leave
        # through os._exit: tearing down NCCL with live graph state has been seen to hang at exit.
        trainer.enable_cuda_graph(False)
        import gc
        gc.collect()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        shutdown()
        return

    # secondary: the library bar on the same GPU (N = 1 only; after every number of ours has been taken)
    variants = {}
    if world == 1 and os.environ.get("B200CLIP_BENCH_VARIANTS", "1") != "0":
        try:
            trainer.enable_cuda_graph(False)
            trainer.grads = trainer.master = trainer.m = trainer.v = None   # hand the memory back first
            del dev_batches
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            variants["torch_eager_bf16"] = torch_eager_bf16_rate(args.model, GLOBAL_BATCH, steps=min(args.steps, 5),
                                                                 warmup=3, dev=dev)
            variants["torch_eager_bf16"]["ours_over_torch_eager"] = value / variants["torch_eager_bf16"]["value"]
        except Exception as exc:  # noqa: BLE001 -- a secondary figure must never break the contract line
            variants["torch_eager_bf16"] = {"error": repr(exc)[:300]}

    # which transport carried the features / log-sum-exps between the towers and the loss (N > 1)
    from construction_clip_b200 import peer as PEER
    feature_exchange = "none (one rank)"
    if world > 1:
        exs = PEER.active()
        if exs:
            if any(e.timed_out() for e in exs):
                raise RuntimeError("peer exchange: a wait for a peer timed out; the measured step is invalid")
            feature_exchange = "peer-memory all-gather over NVLink (csrc/peer.cu), one kernel per rank, in the step's graph"
        else:
            feature_exchange = f"NCCL ({PEER.mode()})"

    peaks = load_peaks()
    f_pair = 3.0 * ORC.flops_pair(ORC.CONFIGS[args.model])          # algorithmic FLOPs per trained pair
    step_tflops_per_gpu = value / world * f_pair / 1e12
    cpu_baseline = None
    if world == 1:  # reported at N=1 only (under torchrun the host threads are partitioned between ranks)
        cpu_rate, cpu_sec, cores = cpu_oracle_rate(steps=3, warmup=1, model_name=args.model)
        cpu_baseline = {"value": cpu_rate, "unit": "pairs/s", "cores": cores, "kind": "port",
                        "sample": f"oracle (fp32 restated upstream CLIP) train step on {CPU_SAMPLE_PAIRS} pairs, "
                                  f"median of 3 after 1 warm-up ({cpu_sec:.2f} s/step)"}

    line = {
        "metric": "image-text pairs/sec (contrastive fine-tune step)", "value": value, "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": f"CLIP {args.model} contrastive fine-tune step (CLIP/train.py:157-171): fwd both towers, "
                        f"all-gathered symmetric InfoNCE, bwd, grad all-reduce, AdamW; global batch {GLOBAL_BATCH} "
                        f"({bl}/GPU), 224x224 images, 77-token prompts, random-init weights seed {SEED}",
            "global_batch": GLOBAL_BATCH, "per_gpu_batch": bl, "parallelism": f"dp{world}",
            "cuda_graph": graph_captured, "cuda_graphs_captured": graphs_captured,
            "packed_text": {
                "on": bool(TW.PACK_TEXT and text_rows[0] is not None),
                "why": "under upstream's causal mask nothing after a caption's EOT reaches the pooled feature, and those "
                       "positions get exactly-zero gradients: the text tower runs on sum(EOT position + 1) rows (rounded "
                       "up to a bucket) instead of per_gpu_batch x 77; same features / loss / gradients "
                       "(tests/test_model_gpu.py::test_packed_text_train_step_vs_oracle_b64)",
                "caption_lengths": "U{3..76} tokens + EOT (SURVEY 8(d) synthetic recipe)",
                "rows_executed_per_batch": text_rows, "rows_real_per_batch": real_rows, "rows_unpacked": bl * 77},
            "l2_policy": "inputs and per-step activations (>10 GB/step) exceed the 126 MB L2; no explicit flush",
            "feature_exchange": feature_exchange,
            "mlp_saved_for_backward": "uint8 code of QuickGELU'(c_fc(x))" if TW.QGELU_D8 else "bf16 pre-activation",
            "final_loss": final_loss,
            "algorithmic_gflop_per_pair": f_pair / 1e9,
            "step_tflops_per_gpu": step_tflops_per_gpu,
            "step_frac_of_bf16_sustained_peak": step_tflops_per_gpu / peaks["bf16_sustained"],
            "step_frac_of_bf16_burst_peak": step_tflops_per_gpu / peaks["bf16_burst"],
            # the same step counted by the GEMM FLOPs the kernels actually EXECUTED (packed text tower, pooled last
            # block): sum of 2MNK over every GEMM launch of one step, attention excluded (< 1 %)
            "executed_gemm_gflop_per_pair": exec_flops_step / bl / 1e9,
            "executed_over_algorithmic_flops": exec_flops_step / bl / f_pair,
            "step_executed_tflops_per_gpu": value / world * (exec_flops_step / bl) / 1e12,
            "step_executed_frac_of_bf16_burst_peak": value / world * (exec_flops_step / bl) / 1e12 / peaks["bf16_burst"],
        },
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_u8, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms / args.steps,
                "host_format": "pinned uint8 [B,3,224,224] pixels + int32 [B,77] tokens per rank; ToTensor + Normalize fused "
                               "into the patch-embedding im2col (construction_clip_b200.data.preprocess_uint8)"},
        "e2e_fp32_host": {"value": e2e32_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes,
                          "d2h_bytes_per_step": 4, "ms_per_step": e2e32_ms / args.steps,
                          "host_format": "pinned fp32 [B,3,224,224] normalised tensors (upstream preprocess) + int32 tokens"},
        "gpu_launches": launches,
        "clocks": clocks,
        "variants": variants,
        "roofline": {
            "bound": "tensor", "achieved": gemm_tflops, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
            "frac": gemm_tflops / peaks["bf16_sustained"], "traffic": traffic,
            "traffic_unit": f"DRAM bytes per GEMM launch (ncu, profiles/{os.path.basename(tpath)})",
            "algorithmic_bytes_per_launch_avg": gemm_alg_bytes,
            "kernel": "gemm_bf16_kernel (tcgen05, all fwd/dgrad/wgrad launches)",
            "peak_source": f"{peaks['source']} (sustained cuBLAS bf16; kernel timed inside a long step)",
            "launches_timed": len(gemm_prof),
            "how": "CUDA events around every GEMM launch in 2 serialised eager steps right after the timed region "
                   "(the timed region overlaps the two towers on two streams"
                   + (" inside a CUDA-graph replay)" if graph_captured else ")"),
            "share_of_step": gemm_ms / instr_ms,
            "algorithmic_flops_per_launch_avg": gemm_flops / max(1, len(gemm_prof)),
        },
        "cpu_baseline": cpu_baseline,
    }
    emit(line)
    shutdown()


# ------------------------------------------------------------------------------------------------
# Secondary modes: the other BASELINE.json configs (`--config 1|3|4|5`).  The headline stays config 2.
SECONDARY = {
    # config: (model, what one "step" is, units per step per GPU, unit name)
    1: ("ViT-B/32", "model(image[32], text[16]) -> logits.softmax.argmax (CLIP/predict.py:40-54)", 32, "images/s"),
    3: ("ViT-B/16", "encode_image on 4096 images per GPU in chunks of 1024 (parse_coco-style prefix extraction)", 4096, "images/s"),
    4: ("ViT-L/14", "encode_image + encode_text at batch 512 per GPU", 512, "pairs/s"),
    5: ("ViT-L/14@336px", "contrastive fine-tune step, 512 pairs per GPU (global 4096 on 8 GPUs), global fused logits/loss, "
                          "activation recompute", 512, "pairs/s"),
}


def run_secondary(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CLIP hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from construction_clip_b200 import lib as L
    from construction_clip_b200.model import CLIP, CONFIGS
    from oracle import clip_oracle as ORC  # FLOP model + synthetic-input recipe + the config-1 CPU baseline

    name, what, units, unit = SECONDARY[args.config]
    cfg = CONFIGS[name]
    torch.manual_seed(SEED)
    model = CLIP(cfg).to(dev)
    ls = model.logit_scale.data.float().clone()
    model = model.to(torch.bfloat16)
    model.logit_scale.data = ls
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    R = cfg.image_resolution
    bpg = int(os.environ.get("BENCH_PAIRS_PER_GPU", str(units)))   # tuning runs only; default = the BASELINE shape
    trainer = None
    if args.config == 1:
        model.eval()
        img_h = ORC.synth_images(32, R, seed=SEED).pin_memory()
        tok_h = ORC.synth_tokens(16, seed=SEED, min_len=3, max_len=12).to(torch.int32).pin_memory()
        img_d, tok_d = img_h.to(dev), tok_h.to(dev)
        f_unit = (32 * ORC.flops_image(ORC.CONFIGS[name]) + 16 * ORC.flops_text(ORC.CONFIGS[name])) / 32
        bpg = 32

        def step_dev():
            with torch.no_grad():
                lpi, _ = model(img_d, tok_d)
                return lpi.softmax(dim=-1).argmax(dim=1)

        def step_host():
            with torch.no_grad():
                lpi, _ = model(img_h.to(dev, non_blocking=True), tok_h.to(dev, non_blocking=True))
                return lpi.softmax(dim=-1).argmax(dim=1).cpu()
        h2d, d2h = img_h.numel() * 4 + tok_h.numel() * 4, 32 * 8
    elif args.config == 3:
        model.eval()
        chunk = min(1024, bpg)
        img_h = ORC.synth_images(chunk, R, seed=SEED + rank).to(torch.bfloat16).pin_memory()
        img_d = img_h.to(dev)
        f_unit = ORC.flops_image(ORC.CONFIGS[name])

        def step_dev():
            with torch.no_grad():
                return [model.encode_image(img_d) for _ in range(bpg // chunk)][-1]

        def step_host():
            with torch.no_grad():
                return [model.encode_image(img_h.to(dev, non_blocking=True)).cpu() for _ in range(bpg // chunk)][-1]
        h2d, d2h = (bpg // chunk) * img_h.numel() * 2, bpg * cfg.embed_dim * 2
    elif args.config == 4:
        model.eval()
        img_h = ORC.synth_images(bpg, R, seed=SEED + rank).to(torch.bfloat16).pin_memory()
        tok_h = ORC.synth_tokens(bpg, seed=SEED + rank).to(torch.int32).pin_memory()
        img_d, tok_d = img_h.to(dev), tok_h.to(dev)
        f_unit = ORC.flops_pair(ORC.CONFIGS[name])

        def step_dev():
            with torch.no_grad():
                return model.encode_image(img_d), model.encode_text(tok_d)

        def step_host():
            with torch.no_grad():
                a = model.encode_image(img_h.to(dev, non_blocking=True))
                b = model.encode_text(tok_h.to(dev, non_blocking=True))
                return a.cpu(), b.cpu()
        h2d, d2h = img_h.numel() * 2 + tok_h.numel() * 4, 2 * bpg * cfg.embed_dim * 2
    else:
        from construction_clip_b200.train import ClipTrainer
        model.train()
        trainer = ClipTrainer(model, lr=1e-5, warmup_steps=5000, total_steps=100000, recompute=True)
        img_h = ORC.synth_images(bpg, R, seed=SEED + 17 * rank).to(torch.bfloat16).pin_memory()
        tok_h = ORC.synth_tokens(bpg, seed=SEED + 17 * rank).to(torch.int32).pin_memory()
        img_d, tok_d = img_h.to(dev), tok_h.to(dev)
        f_unit = 3.0 * ORC.flops_pair(ORC.CONFIGS[name])

        def step_dev():
            return trainer.step(img_d, tok_d)

        def step_host():
            return trainer.step_from_host(img_h, tok_h, next_batch=(img_h, tok_h))
        h2d, d2h = img_h.numel() * 2 + tok_h.numel() * 4, 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, out

    for _ in range(max(3, args.warmup)):
        step_dev()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n0 = L.launch_count()
    ms, out = timed(step_dev, args.steps)
    launches = L.launch_count() - n0
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_host()
    e2e_ms, _ = timed(step_host, args.steps)
    torch.cuda.synchronize()
    final = None
    if args.config == 5:
        final = float(out.item())
    if rank == 0:
        peaks = load_peaks()
        value = world * bpg / (ms * 1e-3)
        tf = value / world * f_unit / 1e12
        cpu_baseline = None
        if args.config == 1 and world == 1:   # BASELINE configs[0] is the reference's own CPU-runnable case
            torch.set_num_threads(os.cpu_count() or 1)
            orc = ORC.build(name, seed=SEED).eval()
            ci, ct = ORC.synth_images(32, R, seed=SEED), ORC.synth_tokens(16, seed=SEED, min_len=3, max_len=12)
            ts = []
            with torch.no_grad():
                for i in range(4):
                    t0 = time.perf_counter()
                    orc(ci, ct)[0].softmax(dim=-1).argmax(dim=1)
                    ts.append(time.perf_counter() - t0)
            sec = statistics.median(ts[1:])
            cpu_baseline = {"value": 32 / sec, "unit": unit, "cores": torch.get_num_threads(), "kind": "port",
                            "sample": f"oracle fp32 forward on the exact config-1 shapes, median of 3 after 1 warm-up ({sec:.3f} s/call)"}
        emit({
            "metric": f"BASELINE config {args.config}: {unit.split('/')[0]} per second", "value": value, "unit": unit,
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"BASELINE config {args.config} (secondary mode, not the headline): CLIP {name}, {what}; "
                                   f"{bpg} per GPU, random-init weights seed {SEED}",
                       "per_gpu_batch": bpg, "parallelism": f"dp{world}", "cuda_graph": False,
                       "algorithmic_gflop_per_unit": f_unit / 1e9, "tflops_per_gpu": tf,
                       "frac_of_bf16_burst_peak": tf / peaks["bf16_burst"],
                       "frac_of_bf16_sustained_peak": tf / peaks["bf16_sustained"], "final_loss": final,
                       "l2_policy": "activations exceed the 126 MB L2 (config 1: 3 warm-up calls, weights stay L2 resident as in "
                                    "a real predict loop)"},
            "e2e": {"value": world * bpg / (e2e_ms * 1e-3), "unit": unit, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_burst"], "unit": "TFLOP/s",
                         "frac": tf / peaks["bf16_burst"], "traffic": None,
                         "kernel": "whole path (tcgen05 GEMM family dominates); algorithmic FLOPs x rate",
                         "peak_source": f"{peaks['source']} (burst cuBLAS bf16)"},
            "cpu_baseline": cpu_baseline,
        })
    if world > 1:
        if trainer is not None:
            trainer.enable_cuda_graph(False)
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--model", default="ViT-B/32")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json config: 2 (default) = the headline train step; 1/3/4/5 = secondary modes")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl != "reference" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # A multi-rank run normally ends within two minutes.  If a rank ever stalls in a collective, fail loudly instead of
        # holding N GPUs until the launcher's own limit (BENCH_WATCHDOG_S=0 disables).
        import threading
        limit = float(os.environ.get("BENCH_WATCHDOG_S", "1200"))
        if limit > 0:
            def _abort():
                sys.stderr.write(f"bench.py: no result after {limit:.0f} s on rank {os.environ.get('RANK')}; aborting\n")
                sys.stderr.flush()
                os._exit(3)
            t = threading.Timer(limit, _abort)
            t.daemon = True
            t.start()
    if args.impl == "reference":
        run_reference(args)
    elif args.config != 2:
        run_secondary(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
