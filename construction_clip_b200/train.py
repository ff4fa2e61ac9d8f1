"""Contrastive fine-tune step of CLIP/train.py:157-171 as one fused call, data parallel over
``torch.distributed`` (one process per GPU, NCCL over NVLink).

  model.zero_grad(); logits = model(image, text); loss = (CE(lpi)+CE(lpt))/2;
  loss.backward(); optimizer.step(); scheduler.step()

becomes ``ClipTrainer.step(image, text)``:
  * each rank encodes its slice of the global batch (both towers, hand-written kernels),
  * the L2-normalised embeddings are all-gathered (the ONE data-path collective of the forward),
  * the fused similarity + softmax-CE kernel computes this rank's rows of the global Bg x Bg
    problem without writing the logits; row log-sum-exps are all-gathered so that every rank can
    form the EXACT gradient of the global loss w.r.t. its own embeddings (no gradient collective
    for activations),
  * backward kernels fill a flat fp32 gradient buffer per tower; chunk by chunk, as soon as the
    backward pass has finished the blocks a chunk covers, it is reduce-scattered, this rank's
    1/N shard is updated by the fused AdamW kernel (fp32 master weights and moments sharded,
    ZeRO-1 style) and the updated bf16 weights (= the tensors the forward kernels read) are
    all-gathered in place -- on a side stream, under the rest of the backward.  With
    ``shard_optimizer=False`` the gradients are sum-all-reduced in full instead.

``clip_contrastive_loss`` exposes the same fused, distributed loss as an autograd op for callers
that keep their own optimiser (``loss.backward()`` then fills ``.grad`` of the local replica with
the gradient of the GLOBAL loss for the local samples; sum-reduce across ranks).
"""
from __future__ import annotations

import math

import torch
import torch.distributed as dist

from . import ops as O
from . import towers as T

bf16, f32 = torch.bfloat16, torch.float32

import os as _os
from . import peer as _peer
FEATURE_GATHER_BY_ALLREDUCE = _os.environ.get("B200CLIP_FEATURE_GATHER", "peer") != "allgather"


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _all_gather_rows(x, group, world):
    """[Bl, E] per rank -> [world*Bl, E] (rank-major), NCCL all-gather."""
    if world == 1:
        return x
    out = torch.empty((world * x.shape[0], *x.shape[1:]), device=x.device, dtype=x.dtype)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


class _LossState:
    """Forward of the fused loss on normalised local features; keeps what backward needs."""

    def __init__(self, img_f, txt_f, logit_scale, group):
        world, rank = _world(group)
        self.world, self.rank, self.group = world, rank, group
        Bl, E = img_f.shape
        self.Bl, self.Bg, self.E = Bl, Bl * world, E
        self.row0 = rank * Bl
        self.ls = logit_scale.detach().to(f32).reshape(1).contiguous()
        self.img_n, self.inv_i = O.l2norm_fwd(img_f.contiguous())
        self.txt_n, self.inv_t = O.l2norm_fwd(txt_f.contiguous())
        # The two exchanges between the towers and the loss.  Default: the library's own peer-memory all-gather over
        # NVLink (peer.py / csrc/peer.cu: one kernel per rank, flags + direct loads from every peer, results written
        # straight into the layouts the loss kernels read); NCCL when the exchange is unavailable or switched off.
        ex = _peer.get(group, img_f.device, 2 * Bl * E * 4) if world > 1 else None
        if ex is not None:
            self.img_all = torch.empty((self.Bg, E), device=img_f.device, dtype=f32)
            self.txt_all = torch.empty((self.Bg, E), device=img_f.device, dtype=f32)
            ex.allgather([self.img_n, self.txt_n], [self.img_all, self.txt_all])
        elif world > 1:
            if FEATURE_GATHER_BY_ALLREDUCE:
                # A ring all-gather of this tiny message pays one hop per rank (147 us at 8 GPUs in
                # profiles/r02_timeline_n8.txt).  Summing zero-padded copies gives the same matrix bit for bit (x + 0
                # is exact) through NCCL's latency-optimised all-reduce.
                both = torch.zeros((self.Bg, 2 * E), device=img_f.device, dtype=f32)
                both[self.row0:self.row0 + Bl, :E] = self.img_n
                both[self.row0:self.row0 + Bl, E:] = self.txt_n
                dist.all_reduce(both, group=group)
            else:
                both = _all_gather_rows(torch.cat([self.img_n, self.txt_n], dim=1), group, world)  # [Bg, 2E]
            self.img_all = both[:, :E].contiguous()
            self.txt_all = both[:, E:].contiguous()
        else:
            self.img_all, self.txt_all = self.img_n, self.txt_n
        self.ws = O.clip_loss_workspace(img_f.device, Bl, self.Bg, E)
        lse_i, lse_t, loss_sum, correct = O.clip_loss_fwd(self.img_all, self.txt_all, self.ls, self.row0, Bl, self.ws)
        if ex is not None:
            self.lse_i_all = torch.empty(self.Bg, device=img_f.device, dtype=f32)
            self.lse_t_all = torch.empty(self.Bg, device=img_f.device, dtype=f32)
            stats_all = torch.empty((world, 4), device=img_f.device, dtype=f32)
            stats = torch.cat([loss_sum, correct.to(f32), correct.to(f32)])   # 4 floats = one 16-byte unit
            ex.allgather([lse_i, lse_t, stats], [self.lse_i_all, self.lse_t_all, stats_all])
            stats = stats_all.sum(0)
            loss_sum, self.correct = stats[:2], stats[2]
        elif world > 1:
            lse = _all_gather_rows(torch.stack([lse_i, lse_t], dim=1), group, world)  # [Bg, 2]
            self.lse_i_all, self.lse_t_all = lse[:, 0].contiguous(), lse[:, 1].contiguous()
            stats = torch.cat([loss_sum, correct.to(f32)])
            dist.all_reduce(stats, group=group)
            loss_sum, self.correct = stats[:2], stats[2]
        else:
            self.lse_i_all, self.lse_t_all = lse_i, lse_t
            self.correct = correct.to(f32)[0]
        self.loss = (loss_sum[0] + loss_sum[1]) / (2.0 * self.Bg)  # global mean loss, same on every rank

    def backward(self, grad_out=None):
        """-> (d_img_f bf16 [Bl,E], d_txt_f bf16 [Bl,E], d_logit_scale fp32 [1], local share)."""
        d_img_n, d_txt_n, d_ls = O.clip_loss_bwd(self.img_all, self.txt_all, self.ls, self.lse_i_all, self.lse_t_all,
                                                 grad_out, self.row0, self.Bl, self.ws)
        d_img = O.l2norm_bwd(d_img_n, self.img_n, self.inv_i)
        d_txt = O.l2norm_bwd(d_txt_n, self.txt_n, self.inv_t)
        return d_img, d_txt, d_ls


class _FusedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img_f, txt_f, logit_scale, group):
        st = _LossState(img_f, txt_f, logit_scale, group)
        ctx.st = st
        ctx.ls_dtype = logit_scale.dtype
        ctx.mark_non_differentiable(st.correct)
        return st.loss, st.correct

    @staticmethod
    def backward(ctx, grad_loss, _grad_correct):
        g = grad_loss.detach().to(f32).reshape(1).contiguous()
        d_img, d_txt, d_ls = ctx.st.backward(g)
        ctx.st = None
        return d_img.float(), d_txt.float(), d_ls.reshape(()).to(ctx.ls_dtype), None


def clip_contrastive_loss(model, image, text, group=None, return_correct=False):
    """Fused symmetric InfoNCE over the GLOBAL batch (all ranks of ``group``); differentiable."""
    img_f = model._features("visual", image)
    txt_f = model._features("text", text)
    loss, correct = _FusedLossFn.apply(img_f, txt_f, model.logit_scale, group)
    return (loss, correct) if return_correct else loss


# ------------------------------------------------------------------------------------------------
def _plan_chunks(store, world, min_elems=1 << 20):
    """Cuts a tower's flat parameter range into contiguous chunks (a, b, ready_layer).  ``ready_layer``
    = L >= 0: every gradient inside is final once the backward of transformer block L is done;
    -1: only at the end of the tower's backward (embeddings, conv1, ln_pre and the first blocks).
    Head parameters (ln_post / proj / ln_final / text_projection) are final before the first block;
    tiny runs are merged into a neighbour that becomes final no earlier.  Every chunk length divides
    by ``world`` (reduce-scatter / all-gather in equal shards), else the plan is one chunk."""
    layers = 1 + max((int(n.split("resblocks.")[1].split(".")[0]) for n, *_ in store.entries if "resblocks." in n),
                     default=-1)
    bounds = sorted({layers // 6, layers // 3 + layers // 12, 2 * layers // 3}) if layers >= 6 else []
    bounds = [L for L in bounds if L > 0]
    head = ("ln_post.", "proj", "ln_final.", "text_projection")

    def group(name):  # 0 = final at the end, g >= 1 = final after block bounds[g-1]
        if "resblocks." in name:
            i = int(name.split("resblocks.")[1].split(".")[0])
            return sum(i >= L for L in bounds)
        return len(bounds) if name.startswith(head) else 0

    runs = []  # [start, end, group]
    ends = [o for _, _, o, _ in store.entries][1:] + [store.total]
    for (name, _, o, _), e in zip(store.entries, ends):
        g = group(name)
        if runs and runs[-1][2] == g:
            runs[-1][1] = e
        else:
            runs.append([o, e, g])
    merged = True
    while merged and len(runs) > 1:  # fold small runs into a neighbour (the union is final at the later of the two)
        merged = False
        for i, (a, b, g) in enumerate(runs):
            if b - a < min_elems:
                j = i - 1 if i > 0 and (i == len(runs) - 1 or runs[i - 1][2] <= runs[i + 1][2]) else i + 1
                lo, hi = min(i, j), max(i, j)
                runs[lo:hi + 1] = [[runs[lo][0], runs[hi][1], min(runs[lo][2], runs[hi][2])]]
                merged = True
                break
    out = []
    for a, b, g in runs:
        tag = bounds[g - 1] if g > 0 else -1
        if out and out[-1][2] == tag:
            out[-1] = (out[-1][0], b, tag)
        else:
            out.append((a, b, tag))
    if any((b - a) % world for a, b, _ in out):
        return [(0, store.total, -1)]
    return out


class ClipTrainer:
    """Owns flat fp32 gradients / master weights / Adam moments for both towers and runs the
    whole training step with library kernels.  Hyper-parameters default to the reference's
    ``AdamW(lr=1e-5)`` from transformers (betas 0.9/0.999, eps 1e-6, no weight decay) and
    ``get_linear_schedule_with_warmup(5000, total)`` (CLIP/train.py:143-147)."""

    def __init__(self, model, lr=1e-5, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, warmup_steps=5000,
                 total_steps=None, group=None, device=None, shard_optimizer=True, recompute=False):
        self.model = model
        self.cfg = model.cfg
        self.group = group
        self.world, self.rank = _world(group)
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.warmup_steps, self.total_steps = warmup_steps, total_steps
        self.step_count = 0
        self.recompute = bool(recompute)   # activation recompute per block (towers.blocks_fwd): 1/9 of the activation memory
        self.stores = {k: model._store(k, self.device) for k in ("visual", "text")}
        # ZeRO-1 style optimiser sharding for N > 1: gradients are REDUCE-SCATTERED (each rank receives
        # the sum of its 1/N slice), AdamW runs on that slice only (fp32 master / moments are 1/N the
        # size) and the updated bf16 weights are all-gathered in place -- less NVLink traffic than an
        # all-reduce (fp32 down, bf16 back) and 1/N of the optimiser's HBM traffic.
        # The flat buffers are cut into a few contiguous CHUNKS whose gradients become final at different
        # points of the backward pass (layers finish last-to-first); a chunk's collectives + update run
        # on a side stream as soon as it is final, under the rest of the backward (see _plan_chunks).
        self.sharded = bool(shard_optimizer) and self.world > 1
        # wire format of the sharded gradient reduce-scatter: bf16 halves the NVLink bytes (605 -> 302 MB per step for
        # ViT-B/32) and the optimiser's gradient read; the fp32 accumulators are cast chunk by chunk right before the
        # collective.  B200CLIP_GRAD_WIRE=fp32 keeps the exact fp32 sum.
        import os as _os
        self.grad_wire_bf16 = self.sharded and _os.environ.get("B200CLIP_GRAD_WIRE", "bf16") != "fp32"
        self.gwire = {}
        self.grads, self.master, self.m, self.v, self.chunks, self.chunk_ready = {}, {}, {}, {}, {}, {}
        self._comm_streams = {}
        # the two towers' collectives are issued from two streams; a second communicator keeps the text
        # tower's early chunks from queueing behind the vision tower's last one (one FIFO per communicator)
        self._tower_group = {"visual": group, "text": group}
        if self.sharded:
            ranks = dist.get_process_group_ranks(group) if group is not None else list(range(self.world))
            self._tower_group["text"] = dist.new_group(ranks=ranks)
        for k, st in self.stores.items():
            st.sync()
            self.grads[k] = torch.zeros(st.total, device=self.device, dtype=f32)
            if self.grad_wire_bf16:
                self.gwire[k] = torch.zeros(st.total, device=self.device, dtype=bf16)
            ranges = _plan_chunks(st, self.world) if self.sharded else [(0, st.total, -1)]
            chunks, moff = [], 0
            for a, b, ready in ranges:
                n = (b - a) // self.world if self.sharded else b - a
                chunks.append((a, b, moff, n, ready))  # flat range, offset / length of this rank's slice in master/m/v
                moff += n
            self.chunks[k] = chunks
            self.chunk_ready[k] = {}
            for j, c in enumerate(chunks):
                if c[4] >= 0:
                    self.chunk_ready[k].setdefault(c[4], []).append(j)
            full = st.master_f32()  # fp32 parameters (clip.load) keep their precision; bf16 ones are upcast
            self.master[k] = torch.cat([full[a + self.rank * n * int(self.sharded):][:n] for a, b, _, n, _ in chunks])
            del full
            self.m[k] = torch.zeros(moff, device=self.device, dtype=f32)
            self.v[k] = torch.zeros(moff, device=self.device, dtype=f32)
        self.G = {k: self.stores[k].grad_views(self.grads[k]) for k in self.stores}
        self.ls_master = model.logit_scale.detach().to(f32).reshape(1).clone()
        self.ls_m = torch.zeros(1, device=self.device, dtype=f32)
        self.ls_v = torch.zeros(1, device=self.device, dtype=f32)
        self.last_correct = None
        self.two_streams = True
        self._tower_streams = None
        # weight-gradient GEMMs on a second stream per tower (towers.blocks_bwd), B200CLIP_WGRAD_STREAM=1
        import os
        # (opt-in: measured 8.09 vs 8.13 ms per 128-pair step, 47.99 vs 47.69 ms at 1024 pairs -- no gain: a persistent
        # GEMM CTA owns its SM's shared memory, so a second GEMM only gets SMs as CTAs of the first retire)
        self.wgrad_streams = os.environ.get("B200CLIP_WGRAD_STREAM", "0") == "1"
        self._wgrad_streams = None
        self._hyper_live = False  # True while a CUDA-graph capture / warm-up wants device-side lr
        # CUDA-graph mode (enable_cuda_graph): step-dependent scalars live in device memory
        self._use_graph = False
        self._hyper = torch.zeros(3, device=self.device, dtype=f32)
        self._hyper_host = torch.zeros((64, 3), dtype=f32).pin_memory() if self.device.type == "cuda" else None
        self._hyper_slot = 0
        self._hyper_events = [None] * 64
        self._rows_hint = None
        self._rows_cache = {}
        self._graphs = {}        # (input shapes / dtypes, packed text rows) -> captured step
        self._static_in = {}     # input shapes / dtypes -> static input buffers shared by those graphs
        self._graph_pool = None
        # Parameters that are not views of the bf16 shadow go stale when this trainer updates the weights;
        # CLIP.state_dict() calls write_back() while `dirty` (CLIP/train.py:210-216 saves model.state_dict()).
        import weakref
        self._has_unlinked = any(st.unlinked() for st in self.stores.values())
        self.dirty = False
        self._grads_reduced = False
        model._trainer = weakref.ref(self)

    def enable_cuda_graph(self, on=True):
        """Capture forward + loss + backward + gradient collectives + AdamW once per input shape and replay it:
        removes the ~500 launches / step of host work, which dominates when the per-GPU batch is
        small (strong scaling at 8 GPUs)."""
        self._use_graph = bool(on)
        if not on:
            self._graphs, self._static_in, self._graph_pool = {}, {}, None
        return self

    def current_lr(self):
        # lr used by the k-th optimizer.step() (k = 0, 1, ...) under LambdaLR: lr * lambda(k)
        s = self.step_count - 1
        if self.warmup_steps and s < self.warmup_steps:
            return self.lr * s / max(1, self.warmup_steps)
        if self.total_steps:
            return self.lr * max(0.0, (self.total_steps - s) / max(1, self.total_steps - self.warmup_steps))
        return self.lr

    def text_rows(self, text):
        """Static row count of the packed text tower for this batch (towers.PACK_TEXT), or None when the tower
        runs unpacked: the real count sum(EOT position + 1) rounded up to a bucket (~3 % of B x 77), so that a
        handful of CUDA graphs covers every batch.  The count comes from the host copy of the tokens when
        ``step_from_host`` has seen one, from a small cache keyed by the token tensor's identity and version
        for device-resident batches that are fed repeatedly, and otherwise costs one host sync."""
        B, S = text.shape
        if not T.PACK_TEXT or S > 128 or B * S < T.PACK_MIN_ROWS:
            return None
        hint, self._rows_hint = self._rows_hint, None
        if hint is None:
            # the cache holds a reference to the tensor: its address cannot be recycled for other tokens while the
            # entry lives, and the version counter catches in-place edits
            key = (text.data_ptr(), text._version, tuple(text.shape), text.stride(), text.dtype)
            hit = self._rows_cache.get(key)
            if hit is None:
                hint = int((text.argmax(-1) + 1).sum().item())
                if len(self._rows_cache) >= 8:
                    self._rows_cache.pop(next(iter(self._rows_cache)))
                self._rows_cache[key] = (hint, text)
            else:
                hint = hit[0]
        g = max(256, -(-(B * S // 32) // 256) * 256)
        return min(B * S, -(-hint // g) * g)

    def forward_backward(self, image, text, fused_update=False, text_rows="auto"):
        """Fills the flat gradient buffers with d(global loss)/d(params); returns the loss tensor.
        ``fused_update`` (used by ``step``): with a sharded optimiser each tower's gradient
        reduce-scatter, AdamW on the local shard and weight all-gather are issued on the tower's own
        stream right after its backward; otherwise the gradients are sum-all-reduced in full.

        The two towers are independent until the loss, so they run on two CUDA streams: whenever a
        persistent GEMM of one tower leaves SMs idle in its last (partial) wave, CTAs of the other
        tower's kernel fill them -- at 128 pairs / GPU most GEMMs are 1.0x .. 4.1x waves of tiles."""
        cfg = self.cfg
        dev = self.device
        for k in self.grads:
            self.grads[k].zero_()
        self._grads_reduced = False
        Wv, Wt = self.stores["visual"].W, self.stores["text"].W
        two = self.two_streams and dev.type == "cuda"
        main = torch.cuda.current_stream(dev)
        if two:
            if self._tower_streams is None:
                self._tower_streams = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
            sv, stt = self._tower_streams
            sv.wait_stream(main)
            stt.wait_stream(main)
        else:
            sv = stt = main
        with torch.cuda.stream(sv):
            img_f, saved_i = T.vision_fwd(Wv, cfg, image, True, recompute=self.recompute)
        if text_rows == "auto":
            text_rows = self.text_rows(text)
        with torch.cuda.stream(stt):
            txt_f, saved_t = T.text_fwd(Wt, cfg, text, True, rows=text_rows, recompute=self.recompute)
        if two:
            main.wait_stream(sv)
            main.wait_stream(stt)
            img_f.record_stream(main)
            txt_f.record_stream(main)
        st = _LossState(img_f, txt_f, self.ls_master, self.group)
        d_img, d_txt, d_ls = st.backward(None)
        if two:
            sv.wait_stream(main)
            stt.wait_stream(main)
            d_img.record_stream(sv)
            d_txt.record_stream(stt)
        work = []
        hyper = self._hyper if self._hyper_live else None
        fused = fused_update and self.sharded
        wg = (None, None)
        if self.wgrad_streams and two:
            if self._wgrad_streams is None:
                self._wgrad_streams = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
            wg = self._wgrad_streams
        for k, s_tower, bwd, W_, saved_, dfeat, s_wg in (("visual", sv, T.vision_bwd, Wv, saved_i, d_img, wg[0]),
                                                         ("text", stt, T.text_bwd, Wt, saved_t, d_txt, wg[1])):
            with torch.cuda.stream(s_tower):
                if s_wg is not None:
                    s_wg.wait_stream(s_tower)
                bwd(W_, self.G[k], cfg, saved_, dfeat, self._chunk_callback(k, s_tower, hyper, s_wg) if fused else None,
                    wgrad_stream=s_wg)
                if self.world > 1:
                    if fused:
                        self._sharded_update(k, hyper, ready=-1)  # what only becomes final at the end
                        if k in self._comm_streams:
                            s_tower.wait_stream(self._comm_streams[k])
                    else:
                        work.append(dist.all_reduce(self.grads[k], group=self.group, async_op=True))
                        self._grads_reduced = True
        if two:
            main.wait_stream(sv)
            main.wait_stream(stt)
        if self.world > 1:
            work.append(dist.all_reduce(d_ls, group=self.group, async_op=True))
            for w in work:
                w.wait()
        self.d_ls = d_ls
        self.last_correct = st.correct
        return st.loss

    def _adam_args(self, hyper):
        b1, b2 = self.betas
        return dict(lr=self.current_lr() if hyper is None else 0.0, beta1=b1, beta2=b2, eps=self.eps,
                    weight_decay=self.wd, grad_scale=1.0, step=self.step_count if hyper is None else 0, hyper=hyper)

    def _sharded_update(self, k, hyper, ready=None, only=None, reduced=False):
        """reduce-scatter(grad) -> AdamW on the local shard -> all-gather(bf16 weights) on the current
        stream, for the chunks of tower ``k`` selected by ``only`` (indices) / ``ready`` (their
        ready-layer tag) -- all of them by default.  ``reduced``: the flat gradients already hold the
        sum over ranks (forward_backward all-reduced them), so this rank's slice is used as it is."""
        st = self.stores[k]
        g = self.grads[k]
        for j, (a, b, moff, n, tag) in enumerate(self.chunks[k]):
            if (only is not None and j not in only) or (ready is not None and tag != ready):
                continue
            lo = a + self.rank * n
            grp = self._tower_group[k]
            gs = g[lo:lo + n]
            if not reduced:
                if self.grad_wire_bf16:
                    gw = self.gwire[k]
                    O.cast_f32_to_bf16(g[a:b], gw[a:b])
                    dist.reduce_scatter_tensor(gw[lo:lo + n], gw[a:b], group=grp)
                    gs = gw[lo:lo + n]
                else:
                    dist.reduce_scatter_tensor(gs, g[a:b], group=grp)
            O.adamw(self.master[k][moff:moff + n], st.w[lo:lo + n], gs, self.m[k][moff:moff + n],
                    self.v[k][moff:moff + n], **self._adam_args(hyper))
            dist.all_gather_into_tensor(st.w[a:b], st.w[lo:lo + n], group=grp)

    def _chunk_callback(self, k, tower_stream, hyper, wgrad_stream=None):
        """Called by the tower's backward after each transformer block: chunks that just became final
        are reduced / updated / re-gathered on a side stream while the backward goes on."""
        table = self.chunk_ready[k]
        if not table:
            return None
        if k not in self._comm_streams:
            self._comm_streams[k] = torch.cuda.Stream(device=self.device)
        side = self._comm_streams[k]

        def done(layer):
            idx = table.get(layer)
            if idx:
                side.wait_stream(tower_stream)
                if wgrad_stream is not None:   # the weight gradients of the finished layers come from this stream
                    side.wait_stream(wgrad_stream)
                with torch.cuda.stream(side):
                    self._sharded_update(k, hyper, only=idx)
        return done

    def optimizer_step(self, hyper=None, towers=True, _count=True):
        """AdamW on both flat buffers + logit_scale.  ``hyper`` (device float[3]) carries lr and the
        bias corrections when the step is replayed from a CUDA graph.  Works after either form of
        ``forward_backward``: with a sharded optimiser the gradients are reduce-scattered here unless
        ``forward_backward(fused_update=False)`` has already sum-all-reduced them (then each rank
        just takes its slice); ``step`` fuses the tower updates into ``forward_backward`` and calls
        this with ``towers=False`` for logit_scale only."""
        if _count and hyper is None:
            self.step_count += 1
        self.dirty = self._has_unlinked
        if towers:
            for k, st in self.stores.items():
                if self.sharded:
                    self._sharded_update(k, hyper, reduced=self._grads_reduced)
                else:
                    O.adamw(self.master[k], st.w, self.grads[k], self.m[k], self.v[k], **self._adam_args(hyper))
        O.adamw(self.ls_master, None, self.d_ls, self.ls_m, self.ls_v, **self._adam_args(hyper))
        with torch.no_grad():
            self.model.logit_scale.copy_(self.ls_master.reshape(()))

    # ------------------------------------------------------------------------------ CUDA graph
    def _push_hyper(self):
        b1, b2 = self.betas
        ev = self._hyper_events[self._hyper_slot]
        if ev is not None:   # the copy that last read this pinned slot (64 steps ago) must have executed
            ev.synchronize()
        row = self._hyper_host[self._hyper_slot]
        row[0] = self.current_lr()
        row[1] = 1.0 - b1 ** self.step_count
        row[2] = 1.0 - b2 ** self.step_count
        self._hyper.copy_(row, non_blocking=True)
        if self._hyper_events[self._hyper_slot] is None:
            self._hyper_events[self._hyper_slot] = torch.cuda.Event()
        self._hyper_events[self._hyper_slot].record()
        self._hyper_slot = (self._hyper_slot + 1) % 64

    def _state_tensors(self):
        out = [self.ls_master, self.ls_m, self.ls_v]
        for k in self.stores:
            out += [self.master[k], self.m[k], self.v[k], self.stores[k].w]
        return out

    def _capture(self, g_img, g_txt, rows, warm):
        """Captures one step on the static inputs (g_img, g_txt) with `rows` packed text rows.  All graphs of
        a trainer share one memory pool (they never run concurrently).  ``warm``: first capture for this
        input shape -- run two eager steps before (and undo them)."""
        dev = self.device
        if self._graph_pool is None:
            self._graph_pool = torch.cuda.graph_pool_handle()
        if warm:
            # warm-up on a side stream (allocator / NCCL / lazy module loading / per-shape index caches), then restore the state
            backup = [t.clone() for t in self._state_tensors()]
            count = self.step_count
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            try:  # whatever happens (a failed warm-up falls back to eager), the optimiser state is restored
                with torch.cuda.stream(side):
                    for _ in range(2):
                        self.step_count += 1
                        self._push_hyper()
                        self._hyper_live = True
                        self.forward_backward(g_img, g_txt, fused_update=True, text_rows=rows)
                        self.optimizer_step(hyper=self._hyper, towers=not self.sharded, _count=False)
                        self._hyper_live = False
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
            finally:
                self._hyper_live = False
                with torch.no_grad():
                    for t, b in zip(self._state_tensors(), backup):
                        t.copy_(b)
                    self.model.logit_scale.copy_(self.ls_master.reshape(()))
                self.step_count = count
                del backup
        backup = [t.clone() for t in self._state_tensors()]   # capture must be side-effect free even if it dies
        graph = torch.cuda.CUDAGraph()
        try:
            with torch.cuda.graph(graph, pool=self._graph_pool):
                self._hyper_live = True
                loss = self.forward_backward(g_img, g_txt, fused_update=True, text_rows=rows)
                self.optimizer_step(hyper=self._hyper, towers=not self.sharded, _count=False)
                self._hyper_live = False
                g_loss = loss.reshape(1).clone()
                g_correct = self.last_correct.reshape(1).clone()
        except Exception:
            self._hyper_live = False
            torch.cuda.synchronize(dev)
            with torch.no_grad():
                for t, b in zip(self._state_tensors(), backup):
                    t.copy_(b)
            raise
        finally:
            del backup
        return graph, g_loss, g_correct

    def _graph_step(self, image, text):
        rows = self.text_rows(text)
        in_key = (tuple(image.shape), image.dtype, tuple(text.shape), text.dtype)
        static = self._static_in.get(in_key)
        warm = static is None
        if static is None:
            static = self._static_in[in_key] = (image.detach().clone(), text.detach().clone())
        g_img, g_txt = static
        g_img.copy_(image, non_blocking=True)
        g_txt.copy_(text, non_blocking=True)
        entry = self._graphs.get((in_key, rows))
        if entry is None:
            try:
                entry = self._graphs[(in_key, rows)] = self._capture(g_img, g_txt, rows, warm)
            except Exception as e:  # capture is an optimisation: fall back to eager launches
                import warnings
                warnings.warn(f"CUDA-graph capture of the training step failed ({e!r}); running eagerly")
                self.enable_cuda_graph(False)
                self._hyper_live = False
                self._rows_hint = rows
                return self.step(image, text)
        graph, g_loss, g_correct = entry
        self.step_count += 1
        self._push_hyper()
        graph.replay()
        self.last_correct = g_correct[0]
        return g_loss[0]

    def write_back(self):
        """Copies the fp32 master weights into parameters that are not views of the bf16 shadow
        (e.g. after ``model.float()``); linked bf16 parameters are already up to date."""
        with torch.no_grad():
            for k, st in self.stores.items():
                todo = st.unlinked()
                if not todo:
                    continue
                full = self.master[k]
                if self.sharded:   # collective: every rank must get here (CLIP.state_dict() on all ranks)
                    full = torch.empty(st.total, device=self.device, dtype=f32)
                    for a, b, moff, n, _ in self.chunks[k]:
                        dist.all_gather_into_tensor(full[a:b], self.master[k][moff:moff + n], group=self.group)
                for name, p, o, s in todo:
                    n = math.prod(s)
                    src = full[o:o + n].view(s)
                    if tuple(s) != tuple(p.shape):
                        src = src[:, :math.prod(p.shape[1:])]
                    p.copy_(src.reshape(p.shape))
                    st._seen[name] = (p.data_ptr(), p._version)   # the shadow already holds these values
        self.dirty = False

    # ---------------------------------------------------------------------------- host-fed steps
    def step_from_host(self, image_host, text_host, next_batch=None):
        """The step as a data-loader consumer sees it (CLIP/train.py:159 `image.to(device)`):
        inputs are HOST tensors (pinned for asynchronous copies); the host->device copy of THIS
        step's batch happens on a copy stream, overlapped with the previous step's compute, and the
        loss is copied back to a pinned host scalar.  Returns that pinned host tensor (valid after
        the next synchronisation).  ``next_batch`` optionally prefetches the following batch."""
        dev = self.device
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged = None
            self._loss_host = torch.zeros(64, dtype=f32).pin_memory()
            self._loss_slot = 0
        cur = torch.cuda.current_stream(dev)

        def stage(img_h, txt_h):
            with torch.cuda.stream(self._copy_stream):
                img_d = img_h.to(dev, non_blocking=True)
                txt_d = txt_h.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            return img_h, img_d, txt_d, ev

        if self._staged is None or self._staged[0] is not image_host:
            self._staged = stage(image_host, text_host)
        _, img_d, txt_d, ev = self._staged
        cur.wait_event(ev)
        img_d.record_stream(cur)
        txt_d.record_stream(cur)
        if T.PACK_TEXT:  # row count of the packed text tower from the host copy of the tokens: no device sync
            self._rows_hint = int((text_host.argmax(-1) + 1).sum())
        # enqueue this step's GPU work FIRST: cudaMemcpyAsync of a large pinned batch can hold the host
        # thread for about the DMA time (measured 1.9 ms for 77 MB, more when copies queue up), and that
        # must not delay the launch of the step the GPU is waiting for
        loss = self.step(img_d, txt_d)
        self._staged = stage(*next_batch) if next_batch is not None else None
        slot = self._loss_host[self._loss_slot:self._loss_slot + 1]
        self._loss_slot = (self._loss_slot + 1) % 64
        slot.copy_(loss.reshape(1), non_blocking=True)
        return slot

    def step(self, image, text):
        """One optimisation step on this rank's slice (image [Bl,3,R,R], text [Bl,77]) of the
        global batch; returns the global mean loss as a device tensor (no host sync)."""
        self.dirty = self._has_unlinked
        if self._use_graph:   # one graph per (input shape, packed-text row bucket)
            return self._graph_step(image, text)
        self.step_count += 1
        loss = self.forward_backward(image, text, fused_update=True)
        self.optimizer_step(towers=not self.sharded, _count=False)
        return loss
