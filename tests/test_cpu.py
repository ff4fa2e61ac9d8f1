"""CPU suite (`-m "not gpu"`): the oracle against its golden vectors and an independent
implementation (HuggingFace CLIPModel), the host logic (clip API, state-dict compatibility,
tokenizer, sharded-loss algorithm under 2-rank gloo) and the C-ABI library's symbol table.  No
compute call into libb200clip is made here (there is no GPU)."""
import ctypes
import os
import re
import subprocess
import sys
import warnings

import numpy as np
import pytest
import torch

from helpers import SEED, golden, oracle_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------------ oracle
@pytest.mark.parametrize("case,name,ni,nt", [("tiny_fwd_6x4", "tiny", 6, 4), ("vitb32_fwd_4x3", "ViT-B/32", 4, 3)])
def test_oracle_matches_golden_forward(case, name, ni, nt):
    from oracle import clip_oracle as O
    g = golden(case)
    orc = oracle_model(name)
    img = O.synth_images(ni, O.CONFIGS[name].image_resolution, seed=SEED)
    tok = O.synth_tokens(nt, seed=SEED, min_len=3, max_len=12)
    assert np.array_equal(tok.numpy(), g["tokens"])
    np.testing.assert_allclose([img.double().sum().item(), img.double().abs().sum().item()], g["image_checksum"], rtol=1e-9)
    with torch.no_grad():
        lpi, lpt = orc(img, tok)
        fi, ft = orc.encode_image(img), orc.encode_text(tok)
    np.testing.assert_allclose(lpi.numpy(), g["logits_per_image"], atol=2e-5)
    np.testing.assert_allclose(fi.numpy(), g["image_features"], atol=2e-5)
    np.testing.assert_allclose(ft.numpy(), g["text_features"], atol=2e-5)
    assert torch.equal(lpt, lpi.t())


def test_oracle_matches_golden_train_tiny():
    from oracle import clip_oracle as O
    g = golden("tiny_train_8")
    orc = oracle_model("tiny")
    img = O.synth_images(8, 64, seed=SEED)
    tok = O.synth_tokens(8, seed=SEED, min_len=3, max_len=12)
    lpi, lpt = orc(img, tok)
    loss = O.clip_loss(lpi, lpt)
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-5
    gn = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    for n, p in orc.named_parameters():
        assert abs(p.grad.double().norm().item() - gn[n]) <= 1e-4 * max(gn[n], 1e-6), n
    np.testing.assert_allclose(orc.visual.ln_post.weight.grad.numpy(), g["grad::visual.ln_post.weight"], atol=1e-6)


@pytest.mark.parametrize("name", ["tiny", "ViT-B/32"])
def test_oracle_matches_huggingface(name):
    """Independent second implementation of the same published model."""
    from oracle import clip_oracle as O
    cfg = O.CONFIGS[name]
    orc = O.build(name, seed=SEED, jitter=0.05)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hf = O.build_hf(cfg)
    res = hf.load_state_dict(O.to_hf_state_dict(orc.state_dict(), cfg), strict=False)
    assert not res.unexpected_keys and all("position_ids" in k for k in res.missing_keys)
    img = O.synth_images(3, cfg.image_resolution, seed=SEED)
    tok = O.synth_tokens(5, seed=SEED)
    with torch.no_grad():
        lpi, _ = orc(img, tok)
        out = hf(input_ids=tok, pixel_values=img)
    assert (lpi - out.logits_per_image).abs().max().item() < 2e-5
    fi = orc.encode_image(img)
    cos = torch.nn.functional.cosine_similarity(fi, out.image_embeds, dim=1)
    assert cos.min().item() > 0.999999
    # the loss formula of CLIP/train.py:162-166 == transformers' clip_loss
    from transformers.models.clip.modeling_clip import clip_loss
    sq = orc(O.synth_images(5, cfg.image_resolution, seed=1), tok)
    assert abs(O.clip_loss(*sq).item() - clip_loss(sq[1]).item()) < 1e-6


def test_oracle_semantics_edge_cases():
    """Causal mask does NOT mask padding; EOT pooling = first arg-max; QuickGELU constant."""
    from oracle import clip_oracle as O
    orc = oracle_model("tiny")
    tok = O.synth_tokens(2, seed=3, min_len=5, max_len=5)
    with torch.no_grad():
        a = orc.encode_text(tok)
        tok2 = tok.clone()
        tok2[:, 40:] = 7      # tokens after EOT cannot influence the pooled EOT row (causal) ...
        b = orc.encode_text(tok2)
        tok3 = tok.clone()
        tok3[:, 2] = 9        # ... but tokens before it do
        c = orc.encode_text(tok3)
    assert torch.allclose(a, b, atol=1e-6) and not torch.allclose(a, c, atol=1e-4)
    x = torch.linspace(-3, 3, 7)
    assert torch.allclose(O.QuickGELU()(x), x * torch.sigmoid(1.702 * x))
    assert abs(orc.logit_scale.item() - np.log(1 / 0.07)) < 1e-6
    assert abs(O.flops_pair(O.CONFIGS["ViT-B/32"]) / 1e9 - 14.705) < 0.01
    assert abs(O.flops_image(O.CONFIGS["ViT-L/14@336px"]) / 1e9 - 381.92) < 0.05


# ------------------------------------------------------------------------------------ clip API
def test_clip_api_surface_and_state_dict():
    import clip
    assert clip.available_models() == ["ViT-B/32", "ViT-B/16", "ViT-L/14", "ViT-L/14@336px"]
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        model, preprocess = clip.load("ViT-B/32", device="cpu")
    assert any("RANDOM" in str(x.message) for x in w)
    from oracle import clip_oracle as O
    osd = O.build("ViT-B/32").state_dict()
    sd = model.state_dict()
    assert len(sd) == 302 and list(sd.keys()) == list(osd.keys())
    assert all(sd[k].shape == osd[k].shape for k in sd)
    assert abs(sum(p.numel() for p in model.parameters()) / 1e6 - 151.28) < 0.01
    model.load_state_dict(osd)            # strict, like CLIP/predict.py:14-16
    assert model.visual.input_resolution == 224 and model.context_length == 77 and model.vocab_size == 49408
    assert len(model.visual.transformer.resblocks) == 12 and len(model.transformer.resblocks) == 12
    assert model.dtype == torch.float32
    for attr in ("encode_image", "encode_text", "forward", "token_embedding", "positional_embedding", "ln_final",
                 "text_projection", "logit_scale"):
        assert hasattr(model, attr)
    from PIL import Image
    t = preprocess(Image.new("RGB", (300, 200), (10, 200, 30)))
    assert t.shape == (3, 224, 224) and t.dtype == torch.float32
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.zeros(1, 3, 224, 224), torch.zeros(1, 77, dtype=torch.int))
    with pytest.raises(RuntimeError):
        clip.load("RN50", device="cpu")
    with pytest.raises(RuntimeError):
        clip.load("not-a-model", device="cpu")
    # build_model infers every hyper-parameter from tensor shapes (upstream behaviour)
    from clip.model import build_model
    m2 = build_model(O.build("tiny").state_dict())
    assert m2.cfg.vision_width == 128 and m2.cfg.vision_layers == 2 and m2.cfg.embed_dim == 64 and m2.cfg.vision_patch_size == 32


def test_tokenizer_with_synthetic_vocab(tmp_path, monkeypatch):
    """The real merge table is upstream data that is absent offline; the algorithm is exercised
    with a synthetic one (byte-level BPE, SOT/EOT framing, zero padding, overflow error)."""
    import gzip
    merges = ["#version: synthetic", "h e", "l l", "he ll", "o </w>", "hell o</w>", "w o", "r l", "wo rl", "worl d</w>"]
    merges += [f"x{i} y{i}" for i in range(49152 - 256 - 2 - len(merges) + 1)]
    path = tmp_path / "bpe.txt.gz"
    with gzip.open(path, "wb") as fh:
        fh.write("\n".join(merges).encode())
    monkeypatch.setenv("CLIP_BPE_PATH", str(path))
    import clip
    import clip.clip as cc
    cc._tokenizer = None
    out = clip.tokenize(["hello world", "Hello   WORLD"])
    assert out.shape == (2, 77) and out.dtype == torch.int32
    assert out[0, 0] == 49406 and torch.equal(out[0], out[1])
    n = int((out[0] != 0).sum())
    assert n == 4 and out[0, n - 1] == 49407 and out[0].argmax() == n - 1
    tok = cc._tokenizer
    assert tok.decode(out[0, 1:n - 1].tolist()).strip() == "hello world"
    with pytest.raises(RuntimeError, match="too long"):
        clip.tokenize("a b c d e f g h i j k l m n o p q r s t u v w x y z " * 4)
    t = clip.tokenize("a b c d e f g h i j k l m n o p q r s t u v w x y z " * 4, truncate=True)
    assert t[0, -1] == 49407
    cc._tokenizer = None


def test_tokenize_without_vocab_fails_loudly(monkeypatch):
    import clip
    import clip.clip as cc
    monkeypatch.setenv("CLIP_BPE_PATH", "/nonexistent/bpe.txt.gz")
    cc._tokenizer = None
    with pytest.raises(FileNotFoundError, match="CLIP_BPE_PATH"):
        clip.tokenize("hello")


# --------------------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    from construction_clip_b200 import build as B, lib as L
    path = B.build()
    hdr = open(os.path.join(ROOT, "include", "b200clip.h")).read()
    declared = set(re.findall(r"\b(b200clip_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("b200clip_ctx")
    assert len(declared) >= 24
    cdll = ctypes.CDLL(str(path))
    for name in sorted(declared):
        assert hasattr(cdll, name), f"{name} declared in b200clip.h but not exported"
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = L.load()
    assert lib.b200clip_abi_version() == L.ABI_VERSION
    # without a GPU the context cannot be created and says so (no silent fallback)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
            L.ctx(0)


def test_ops_reject_cpu_tensors():
    from construction_clip_b200 import ops as O
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        O.layernorm_fwd(torch.zeros(4, 64, dtype=torch.bfloat16), torch.ones(64, dtype=torch.bfloat16),
                        torch.zeros(64, dtype=torch.bfloat16))


def test_product_path_does_not_import_oracle():
    for base in ("construction_clip_b200", "clip"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith(".py"):
                    src = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src, os.path.join(dirpath, f)


# --------------------------------------------------------------------------- sharded loss (gloo)
def _sharded_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from oracle import clip_oracle as O
    from oracle import sharded_loss as S
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    Bg, E = 12, 32
    img = torch.nn.functional.normalize(torch.randn(Bg, E, dtype=torch.float64), dim=1)
    txt = torch.nn.functional.normalize(img * 0.5 + 0.3 * torch.randn(Bg, E, dtype=torch.float64), dim=1)
    ls = torch.tensor(np.log(1 / 0.07), dtype=torch.float64)
    bl = Bg // world
    row0 = rank * bl
    # what each rank holds locally, then the collectives of construction_clip_b200/train.py
    both = torch.cat([img[row0:row0 + bl], txt[row0:row0 + bl]], 1)
    gathered = torch.empty(Bg, 2 * E, dtype=torch.float64)
    dist.all_gather_into_tensor(gathered, both)
    img_all, txt_all = gathered[:, :E].contiguous(), gathered[:, E:].contiguous()
    lse_i, lse_t, loss_sum, correct = S.local_fwd(img_all, txt_all, ls, row0, bl)
    lse = torch.empty(Bg, 2, dtype=torch.float64)
    dist.all_gather_into_tensor(lse, torch.stack([lse_i, lse_t], 1))
    stats = torch.cat([loss_sum, correct.double().reshape(1)])
    dist.all_reduce(stats)
    loss = (stats[0] + stats[1]) / (2 * Bg)
    d_img, d_txt, d_ls = S.local_bwd(img_all, txt_all, ls, lse[:, 0].contiguous(), lse[:, 1].contiguous(), row0, bl)
    dist.all_reduce(d_ls)
    # single-process reference: CLIP/train.py:162-166 + autograd
    ir, tr, lr = img.clone().requires_grad_(True), txt.clone().requires_grad_(True), ls.clone().requires_grad_(True)
    L = lr.exp() * ir @ tr.t()
    ref = O.clip_loss(L, L.t())
    # clip_loss casts to fp32; redo in fp64 for a tight comparison
    lab = torch.arange(Bg)
    ref = (torch.nn.functional.cross_entropy(L, lab) + torch.nn.functional.cross_entropy(L.t(), lab)) / 2
    ref.backward()
    ok = (abs(loss.item() - ref.item()) < 1e-10 and torch.allclose(d_img, ir.grad[row0:row0 + bl], atol=1e-10)
          and torch.allclose(d_txt, tr.grad[row0:row0 + bl], atol=1e-10) and abs(d_ls.item() - lr.grad.item()) < 1e-9
          and int(stats[2].item()) == int((L.argmax(1) == lab).sum()))
    q.put((rank, bool(ok), loss.item(), ref.item()))
    dist.destroy_process_group()


def test_sharded_loss_equals_global_loss_gloo_world2():
    """N > 1 host logic on CPU: local-rows loss + all-gathers == single-device loss and autograd
    gradients (the definition of correct for the data-parallel path)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res


def test_trainer_lr_schedule_matches_reference():
    """get_linear_schedule_with_warmup(5000, total) as used at CLIP/train.py:145-147."""
    from construction_clip_b200.train import ClipTrainer
    t = ClipTrainer.__new__(ClipTrainer)
    t.lr, t.warmup_steps, t.total_steps = 1e-5, 5000, 100000
    def lam(k):
        return k / 5000 if k < 5000 else max(0.0, (100000 - k) / (100000 - 5000))
    for k in (0, 1, 2500, 4999, 5000, 50000, 99999):
        t.step_count = k + 1
        assert abs(t.current_lr() - 1e-5 * lam(k)) < 1e-15


def test_bench_reference_arm_runs_on_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--model", "tiny"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    out_lines = r.stdout.strip().splitlines()
    assert len(out_lines) == 1, f"bench.py must print exactly one stdout line, got {len(out_lines)}"
    line = json.loads(out_lines[0])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "pairs/s"


@pytest.mark.parametrize("name", ["ViT-B/32", "ViT-L/14", "tiny"])
def test_optimizer_chunk_plan(name):
    """The trainer's chunked reduce-scatter plan: contiguous cover of the flat buffer, every chunk
    divisible by the world size, and no parameter lands in a chunk that is reduced before the
    parameter's gradient is final (blocks finish last-to-first; embeddings / conv1 / ln_pre at the end)."""
    import types
    from construction_clip_b200.model import CLIP, CONFIGS
    from construction_clip_b200.train import _plan_chunks
    m = CLIP(CONFIGS[name])
    for which in ("visual", "text"):
        entries, off = [], 0
        for n, p in m._tower_named_params(which):
            entries.append((n, p, off, tuple(p.shape)))
            off += (p.numel() + 63) // 64 * 64
        unit = 64 * 840
        store = types.SimpleNamespace(entries=entries, total=(off + unit - 1) // unit * unit)
        for world in (2, 3, 4, 8):
            plan = _plan_chunks(store, world)
            assert plan[0][0] == 0 and plan[-1][1] == store.total
            for (a, b, tag), nxt in zip(plan, plan[1:] + [None]):
                assert b > a and (b - a) % world == 0
                if nxt is not None:
                    assert nxt[0] == b
            head = ("ln_post.", "proj", "ln_final.", "text_projection")
            for n, _, o, _ in entries:
                tag = next(t for a, b, t in plan if a <= o < b)
                if "resblocks." in n:
                    layer = int(n.split("resblocks.")[1].split(".")[0])
                    assert tag <= layer, (n, tag)      # reduced only after block `tag` <= this block is done
                elif not n.startswith(head):
                    assert tag == -1, (n, tag)         # embeddings, conv1, ln_pre: final at the very end


def test_positions_after_eot_are_dead_in_the_oracle():
    """The claim behind the packed text tower (towers.PACK_TEXT) and the pooled last block, proven on the
    fp32 CPU oracle: under upstream's causal mask (a) the text feature of a caption does not depend on
    anything after its EOT token -- truncating the caption at EOT, or replacing the tail by other ids,
    gives the same feature -- and (b) the loss gradient w.r.t. the token/positional embeddings of those
    positions is exactly zero."""
    from oracle import clip_oracle as O
    torch.manual_seed(0)
    orc = O.build("tiny", seed=SEED, jitter=0.05)
    tok = O.synth_tokens(6, seed=3, min_len=3, max_len=20)
    eot = tok.argmax(-1)
    full = orc.encode_text(tok)
    # (a1) garbage after EOT (ids below EOT so that arg-max pooling still finds the same position)
    junk = tok.clone()
    for b in range(tok.shape[0]):
        junk[b, int(eot[b]) + 1:] = torch.randint(1, O.SOT, (tok.shape[1] - int(eot[b]) - 1,))
    assert torch.allclose(orc.encode_text(junk), full, atol=1e-6, rtol=1e-6)
    # (a2) a caption alone, cut at EOT + 1 tokens, through the same blocks with an L x L causal mask
    for b in range(tok.shape[0]):
        L = int(eot[b]) + 1
        x = orc.token_embedding(tok[b:b + 1, :L]) + orc.positional_embedding[:L]
        x = x.permute(1, 0, 2)
        mask = torch.full((L, L), float("-inf")).triu_(1)
        for blk in orc.transformer.resblocks:
            saved = blk.attn_mask
            blk.attn_mask = mask
            try:
                x = blk(x)
            finally:
                blk.attn_mask = saved
        x = orc.ln_final(x.permute(1, 0, 2))
        feat = x[0, L - 1] @ orc.text_projection
        assert torch.allclose(feat, full[b], atol=1e-5, rtol=1e-5), b
    # (b) zero gradient into the dead positions
    emb = (orc.token_embedding(tok) + orc.positional_embedding).detach().requires_grad_(True)
    x = orc.transformer(emb.permute(1, 0, 2)).permute(1, 0, 2)
    x = orc.ln_final(x)
    f = x[torch.arange(tok.shape[0]), eot] @ orc.text_projection
    f.square().sum().backward()
    for b in range(tok.shape[0]):
        assert emb.grad[b, int(eot[b]) + 1:].abs().max().item() == 0.0
        assert emb.grad[b, :int(eot[b]) + 1].abs().max().item() > 0.0


@pytest.mark.parametrize("causal", [False, True])
def test_in_proj_bias_gradient_identities(causal):
    """The identities towers.blocks_bwd uses to get the in_proj_bias gradient without re-reading dqkv:
    the K third is zero (a key bias shifts every score of a row equally) and the V third equals the
    column sum of the gradient w.r.t. the attention output that feeds out_proj (softmax rows sum to 1)."""
    torch.manual_seed(1)
    S, B, d, H = 9, 3, 64, 2
    dh = d // H
    f64 = torch.float64
    w_in = torch.randn(3 * d, d, dtype=f64) * 0.2
    b_in = torch.randn(3 * d, dtype=f64, requires_grad=True)
    x = torch.randn(B, S, d, dtype=f64)
    q, k, v = (x @ w_in.t() + b_in).view(B, S, 3, H, dh).permute(2, 0, 3, 1, 4)
    s_ = q @ k.transpose(-1, -2) / dh ** 0.5
    if causal:
        s_ = s_ + torch.full((S, S), float("-inf"), dtype=f64).triu_(1)
    a = (torch.softmax(s_, -1) @ v).permute(0, 2, 1, 3).reshape(B, S, d)   # what out_proj consumes
    a.retain_grad()
    out = a @ (torch.randn(d, d, dtype=f64) * 0.2).t()
    (out * torch.randn_like(out)).sum().backward()
    gq, gk, gv = b_in.grad.split(d)
    assert gk.abs().max().item() < 1e-12 * max(1.0, gq.abs().max().item())
    assert torch.allclose(gv, a.grad.reshape(-1, d).sum(0), atol=1e-10, rtol=1e-10)
    assert gq.abs().max().item() > 1e-6
