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
            assert rel_err(A, stg[Akey][l].reshape(B, H, S, S)) < TOL, (tower, l)
            dA = eng.tap("dA", tower, l)
            assert rel_err(dA, stg[Gkey][l].reshape(B, H, S, S)) < TOL, (tower, l)
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
