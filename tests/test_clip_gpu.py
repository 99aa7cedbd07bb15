"""GPU: the CLIP interpret() engine through the C ABI vs (a) the committed reference goldens, (b) the oracle on
seeded inputs at ViT-B/32 size with stage-by-stage taps (A_l, dA_l, Abar_l), (c) size-independent properties at
the BASELINE.json batch (64)."""
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as co
from util import rel_err, text_rel_err, TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mmx():
    import mmx_b200
    return mmx_b200


def _engine(mmx, cfg, sd, max_batch):
    return mmx.ClipEngine(mmx.ClipConfig(*cfg.ref_args()), sd, max_batch=max_batch, device="cuda:0")


@pytest.mark.parametrize("tag", ["tiny", "small"])
def test_golden_reference_outputs(mmx, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"clip_{tag}.npz"))
    cfg = co.ClipConfig(*[int(v) for v in g["cfg"]])
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    images, tokens = torch.from_numpy(g["images"]).cuda(), torch.from_numpy(g["tokens"]).cuda()
    eng = _engine(mmx, cfg, sd, tokens.shape[0])
    for sl in (-1, 0, 1):
        rt, ri = mmx.interpret(images, tokens, eng, "cuda:0", sl, sl)
        assert text_rel_err(rt, g[f"distinct.sl{sl}.R_text"]) < TOL
        assert rel_err(ri, g[f"distinct.sl{sl}.R_image"]) < TOL
        rt, ri = mmx.interpret(images[:1], tokens, eng, "cuda:0", sl, sl)     # notebook semantics: one image repeated
        assert text_rel_err(rt, g[f"repeat.sl{sl}.R_text"]) < TOL
        assert rel_err(ri, g[f"repeat.sl{sl}.R_image"]) < TOL


def test_golden_logits(mmx, golden_dir):
    g = np.load(os.path.join(golden_dir, "clip_small.npz"))
    cfg = co.ClipConfig(*[int(v) for v in g["cfg"]])
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    images, tokens = torch.from_numpy(g["images"]).cuda(), torch.from_numpy(g["tokens"]).cuda()
    eng = _engine(mmx, cfg, sd, tokens.shape[0])
    mmx.interpret(images, tokens, eng, "cuda:0")
    assert rel_err(eng.tap("logits"), g["logits_per_image"]) < TOL


@pytest.fixture(scope="module")
def b32(mmx):
    cfg = co.VIT_B32
    sd = co.init_state_dict(cfg, seed=0)
    eng = _engine(mmx, cfg, sd, 64)
    return cfg, sd, eng


@pytest.mark.parametrize("sl", [-1, 0, 6])
def test_vit_b32_vs_oracle_with_stages(mmx, b32, sl):
    cfg, sd, eng = b32
    B = 3
    images, tokens = co.synthetic_inputs(cfg, B, seed=21)
    ot, oi, stg = co.clip_interpret(sd, cfg, images, tokens, sl, sl, return_stages=True)
    rt, ri = mmx.interpret(images.cuda(), tokens.cuda(), eng, "cuda:0", sl, sl)
    assert rel_err(eng.tap("logits"), stg["logits"]) < TOL
    for tower, (Akey, Gkey, Bkey, L, H) in enumerate((("A_v", "G_v", "bar_v", cfg.vision_layers, cfg.vision_heads),
                                                      ("A_t", "G_t", "bar_t", cfg.transformer_layers,
                                                       cfg.transformer_heads))):
        start = L - 1 if sl == -1 else sl
        for l in sorted({start, L - 1}):
            S = stg[Akey][l].shape[-1]
            A = eng.tap("A", tower, l)
            A_ref = stg[Akey][l].reshape(B, H, S, S).clone()
            G_ref = stg[Gkey][l].reshape(B, H, S, S).clone()
            if tower == 1 and os.environ.get("MMX_RAGGED_TEXT", "1") != "0":
                # ragged text rows: tokens after the EOT are dead under the causal mask (their A rows get zero
                # gradient in the reference and their R rows stay identity), so the engine neither computes nor
                # stages them: rows AND columns >= len are zero in the staged A / dA (A is zero there anyway for the
                # live rows; dA = dO V^T of the reference is not, but it only ever multiplies those zeros)
                for b in range(B):
                    n = int(tokens[b].argmax()) + 1
                    A_ref[b, :, n:, :] = 0
                    G_ref[b, :, n:, :] = 0
                    G_ref[b, :, :, n:] = 0
            assert rel_err(A, A_ref) < TOL, (tower, l)
            dA = eng.tap("dA", tower, l)
            assert rel_err(dA, G_ref) < TOL, (tower, l)
            assert rel_err(eng.tap("Abar", tower, l), stg[Bkey][l]) < TOL, (tower, l)
    assert text_rel_err(rt, ot) < TOL
    assert rel_err(ri, oi) < TOL


def test_batch64_properties(mmx, b32):
    """BASELINE.json size (batch 64): per-sample independence (bitwise), micro-batching equality, causal structure."""
    cfg, sd, eng = b32
    images, tokens = co.synthetic_inputs(cfg, 64, seed=33)
    ic, tc = images.cuda(), tokens.cuda()
    rt, ri = mmx.interpret(ic, tc, eng, "cuda:0", 0, 0)
    assert torch.isfinite(rt).all() and torch.isfinite(ri).all()
    # sample b alone == sample b inside the batch (no cross-sample arithmetic; needed for sharded == single-GPU)
    for b in (0, 17, 63):
        rt1, ri1 = mmx.interpret(ic[b:b + 1], tc[b:b + 1], eng, "cuda:0", 0, 0)
        assert torch.equal(rt1[0], rt[b]) and torch.equal(ri1[0], ri[b])
    # text relevance is lower-triangular (causal tower): R - I vanishes above the diagonal, diagonal >= 1
    assert (rt.triu(1) == 0).all()
    assert (rt.diagonal(dim1=1, dim2=2) >= 1).all()
    assert (ri >= 0).all()
    # rows after the EOT token receive no gradient: they stay rows of the identity
    eot = tokens.argmax(-1)
    for b in (0, 5):
        tail = rt[b, eot[b] + 1:, :].cpu()
        assert torch.equal(tail, torch.eye(cfg.context_length)[eot[b] + 1:, :])
    # an engine that micro-batches (max_batch 24 -> chunks 24/24/16) gives identical maps
    eng2 = _engine(mmx, cfg, sd, 24)
    rt2, ri2 = mmx.interpret(ic, tc, eng2, "cuda:0", 0, 0)
    assert torch.equal(rt2, rt) and torch.equal(ri2, ri)
    # host entry point (H2D / D2H inside the call) == device entry point
    ht, hi = eng.interpret_host(images, tokens, 0, 0)
    assert torch.equal(ht, rt.cpu()) and torch.equal(hi, ri.cpu())
    # default mode equals the oracle on one sample at full batch
    rtd, rid = mmx.interpret(ic, tc, eng, "cuda:0")
    ot, oi = co.clip_interpret(sd, cfg, images[40:41], tokens[40:41])
    assert text_rel_err(rtd[40:41], ot) < TOL and rel_err(rid[40:41], oi) < TOL


def test_error_behaviour(mmx, b32):
    cfg, sd, eng = b32
    images, tokens = co.synthetic_inputs(cfg, 2, seed=1)
    with pytest.raises(mmx.MmxError):
        mmx.interpret(images.cuda(), tokens.cuda(), eng, "cuda:0", start_layer=12)     # out of range
    with pytest.raises(mmx.MmxError):
        mmx.interpret(torch.cat([images, images[:1]]).cuda(), tokens.cuda(), eng, "cuda:0")   # 3 images, 2 texts
    with pytest.raises(mmx.MmxError):
        mmx.interpret(images.cuda(), tokens.cuda(), object(), "cuda:0")
    bad = dict(sd); bad.pop("ln_final.bias")
    with pytest.raises(mmx.MmxError):
        mmx.ClipEngine(mmx.ClipConfig(*cfg.ref_args()), bad, max_batch=1)


def test_ragged_text_edge_lengths(mmx):
    """EOT at the first possible position, at the very last position, and an all-zero row (argmax = 0 -> one live
    token): the packed-row text tower must reproduce the dense reference semantics in every case."""
    cfg = co.SMALL
    sd = co.init_state_dict(cfg, seed=6)
    images, tokens = co.synthetic_inputs(cfg, 5, seed=3)
    ctx, eot = cfg.context_length, cfg.vocab_size - 1
    tokens[0] = 0; tokens[0, 0] = cfg.vocab_size - 2; tokens[0, 1] = eot                    # shortest prompt
    tokens[1] = torch.randint(1, cfg.vocab_size - 2, (ctx,)); tokens[1, 0] = cfg.vocab_size - 2
    tokens[1, ctx - 1] = eot                                                                  # EOT in the last slot
    tokens[2] = 0                                                                             # degenerate: argmax = 0
    eng = _engine(mmx, cfg, sd, 5)
    for sl in (-1, 0):
        rt, ri = mmx.interpret(images.cuda(), tokens.cuda(), eng, "cuda:0", sl, sl)
        ot, oi = co.clip_interpret(sd, cfg, images, tokens, sl, sl)
        assert text_rel_err(rt, ot) < TOL and rel_err(ri, oi) < TOL
        assert torch.equal(rt[2].cpu()[1:], torch.eye(ctx)[1:])      # a one-token prompt: only row 0 can change


def test_vit_l14_336_vs_oracle(mmx):
    """BASELINE.json config 5 model (CLIP ViT-L/14@336: 24 vision blocks, 577 tokens, 16 heads; 12 text blocks of
    width 768) on 2 pairs, all layers, against the oracle; also exercises micro-batching (max_batch = 1)."""
    cfg = co.VIT_L14_336
    sd = co.init_state_dict(cfg, seed=0)
    images, tokens = co.synthetic_inputs(cfg, 2, seed=4)
    ot, oi = co.clip_interpret(sd, cfg, images, tokens, 0, 0)
    eng = _engine(mmx, cfg, sd, 1)
    rt, ri = mmx.interpret(images.cuda(), tokens.cuda(), eng, "cuda:0", 0, 0)
    assert ri.shape == (2, 576) and rt.shape == (2, 77, 77)
    assert text_rel_err(rt, ot) < TOL and rel_err(ri, oi) < TOL


class _RefLikeClip:
    """Stands in for the reference ``CLIP`` nn.Module on the GPU box (the reference tree does not travel): what
    ``interpret`` needs from it is ``state_dict()`` with the reference's key names (CLIP/clip/model.py:248-303)."""

    def __init__(self, sd, text_heads):
        import types
        self._sd = dict(sd)
        self.calls = 0
        attn = types.SimpleNamespace(num_heads=text_heads)            # nn.MultiheadAttention.num_heads of the text blocks
        self.transformer = types.SimpleNamespace(resblocks=[types.SimpleNamespace(attn=attn)])

    def state_dict(self):
        self.calls += 1
        return self._sd


def test_notebook_call_form_with_reference_module(mmx, golden_dir):
    """The notebook cells call ``interpret(model=model, image=img, texts=text, device=device)`` with the reference module
    (CLIP_explainability.ipynb cells 12-19): that exact call must work, build ONE engine per module, and give the goldens."""
    g = np.load(os.path.join(golden_dir, "clip_small.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    model, device = _RefLikeClip(sd, int(g["cfg"][8])), "cuda"
    img, text = torch.from_numpy(g["images"][:1]).to(device), torch.from_numpy(g["tokens"]).to(device)
    interpret = mmx.interpret                                            # the only line a notebook changes: the import
    R_text, R_image = interpret(model=model, image=img, texts=text, device=device)
    assert text_rel_err(R_text, g["repeat.sl-1.R_text"]) < TOL and rel_err(R_image, g["repeat.sl-1.R_image"]) < TOL
    R_text, R_image = interpret(model=model, image=img, texts=text, device=device, start_layer=0, start_layer_text=0)
    assert text_rel_err(R_text, g["repeat.sl0.R_text"]) < TOL and rel_err(R_image, g["repeat.sl0.R_image"]) < TOL
    assert model.calls == 1                                              # engine cached per module object
    batch_size = text.shape[0]
    assert R_text.shape == (batch_size, text.shape[1], text.shape[1]) and R_image[0].numel() == (img.shape[-1] // 16) ** 2


@pytest.mark.parametrize("index", [None, 0, 1, 2])
def test_example_py_interpret_golden(mmx, golden_dir, index):
    """CLIP/example.py:8-53 (one image, N texts, ``index``) vs the relevance the unmodified function computes."""
    g = np.load(os.path.join(golden_dir, "clip_example.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    model = _RefLikeClip(sd, int(g["cfg"][8]))
    image, text = torch.from_numpy(g["image"]).cuda(), torch.from_numpy(g["tokens"]).cuda()
    R, logits = mmx.interpret_example(model=model, image=image, text=text, device="cuda", index=index)
    assert rel_err(logits, g["logits_per_image"]) < TOL
    assert rel_err(R, g[f"R.index{index}"]) < TOL


def test_shape_and_token_validation(mmx, golden_dir):
    """The C ABI takes no sizes: wrong shapes / token ids must be refused before any pointer reaches it."""
    g = np.load(os.path.join(golden_dir, "clip_tiny.npz"))
    cfg = co.ClipConfig(*[int(v) for v in g["cfg"]])
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    eng = _engine(mmx, cfg, sd, 4)
    images, tokens = torch.from_numpy(g["images"]), torch.from_numpy(g["tokens"])
    bad = [(images[:, :, :16], tokens), (images, tokens[:, :-1]), (images[:2], tokens), (images[:, :2], tokens)]
    for im, tk in bad:
        with pytest.raises(mmx.MmxError):
            eng.interpret(im.cuda(), tk.cuda())
        with pytest.raises(mmx.MmxError):
            eng.interpret_host(im, tk)
    oob = tokens.clone()
    oob[0, 1] = cfg.vocab_size
    with pytest.raises(mmx.MmxError):
        eng.interpret(images.cuda(), oob.cuda())
    neg = tokens.clone()
    neg[0, 1] = -1
    with pytest.raises(mmx.MmxError):
        eng.interpret_host(images, neg)
    with pytest.raises(mmx.MmxError):
        eng.interpret_host(images, tokens, out=(torch.empty(1, 2, 2), torch.empty(1, 3)))


def test_ragged_rows_leave_no_stale_columns(mmx):
    """A short prompt after a long one in the same engine slot: every column of the staged A / dA planes beyond the prompt must
    be rewritten as zero (the planes are dense [B,H,77,80]; a ragged row covers only round_up(len, 64) columns by itself)."""
    cfg = co.VIT_B32
    sd = co.init_state_dict(cfg, seed=0)
    eng = _engine(mmx, cfg, sd, 2)
    images, _ = co.synthetic_inputs(cfg, 2, seed=3)
    def toks(n):
        t = torch.zeros(2, cfg.context_length, dtype=torch.int64)
        t[:, 0] = cfg.vocab_size - 2
        t[:, 1:1 + n] = 7
        t[:, 1 + n] = cfg.vocab_size - 1
        return t
    eng.interpret(images.cuda(), toks(74).cuda(), 0, 0)           # 76 live rows: fills columns up to 75
    rt, _ = eng.interpret(images.cuda(), toks(10).cuda(), 0, 0)   # 12 live rows
    for what in ("A", "dA"):
        t = eng.tap(what, tower=1, layer=3)
        assert float(t[..., 12:, :].abs().max()) == 0.0 and float(t[..., :, 12:].abs().max()) == 0.0, what
    ot, _ = co.clip_interpret(sd, cfg, images, toks(10), 0, 0)
    assert text_rel_err(rt, ot) < TOL


def test_inputs_still_in_flight_on_the_callers_stream(mmx, b32):
    """The call is ordered after whatever the caller enqueued on its stream (here PyTorch's default stream = the NULL
    stream of the C ABI): images that arrive by an asynchronous H2D copy issued right before interpret() must be the
    images the towers see.  (Until round 2 a NULL stream meant the engine's private stream, which did not wait.)"""
    cfg, sd, eng = b32
    images, tokens = co.synthetic_inputs(cfg, 64, seed=71)
    other, _ = co.synthetic_inputs(cfg, 64, seed=72)
    tok = tokens.cuda()
    ref_t, ref_i = eng.interpret(images.cuda(), tok, 0, 0)
    torch.cuda.synchronize()
    pinned, stale = images.pin_memory(), other.cuda()
    buf = torch.empty_like(stale)
    for _ in range(4):
        buf.copy_(stale)                                   # what a too-early read would see
        torch.cuda.synchronize()
        buf.copy_(pinned, non_blocking=True)               # 38.5 MB in flight on the current stream ...
        rt, ri = eng.interpret(buf, tok, 0, 0)             # ... when the towers are enqueued
        assert torch.equal(rt, ref_t) and torch.equal(ri, ref_i)
