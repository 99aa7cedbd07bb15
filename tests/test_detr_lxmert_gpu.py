"""GPU: DETR Generator.generate_ours and LXMERT GeneratorOurs.generate_ours through the
libmmx kernels vs the committed reference goldens and the oracle at the BASELINE.json shapes."""
import os

import numpy as np
import pytest
import torch

from oracle import detr_oracle as do, lxmert_oracle as lo
from util import rel_err, TOL

pytestmark = pytest.mark.gpu


def _sd(g):
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}


@pytest.mark.parametrize("norm,s10", [(True, True), (False, True), (True, False)])
def test_detr_golden(golden_dir, norm, s10):
    import mmx_b200
    g = np.load(os.path.join(golden_dir, "detr_tiny.npz"))
    eng = mmx_b200.DetrEngine(_sd(g), nhead=do.DETR_TINY.nhead, device="cuda:0")
    gen = mmx_b200.Generator(eng)
    src, pos, tq = (torch.from_numpy(g[k]) for k in ("src", "pos", "tq"))
    out = gen.generate_ours((src.cuda(), pos.cuda()), tq, use_lrp=False, normalize_self_attention=norm,
                            apply_self_in_rule_10=s10)
    assert rel_err(out, g[f"R.n{int(norm)}s{int(s10)}"]) < TOL
    # single-sample call: the reference's return shapes (tensor index -> [1,1,1,S], int -> [1,1,S])
    one = gen.generate_ours((src[:1].cuda(), pos[:1].cuda()), torch.tensor([int(tq[0])]), use_lrp=False,
                            normalize_self_attention=norm, apply_self_in_rule_10=s10)
    assert one.shape == (1, 1, 1, src.shape[2] * src.shape[3])
    assert rel_err(one.reshape(-1), g[f"R.n{int(norm)}s{int(s10)}"][0]) < TOL
    assert gen.generate_ours((src[:1].cuda(), pos[:1].cuda()), int(tq[0]), use_lrp=False).shape == (1, 1, src.shape[2] * src.shape[3])
    assert gen.generate_ours((src.cuda(), pos.cuda()), tq).shape == out.shape      # API default use_lrp=True runs (tests/test_lrp_gpu.py)


def test_detr_r50_shape_vs_oracle():
    """DETR-R50 transformer dims (d=256, 8 heads, 6+6 layers, 100 queries) on a 10x12 feature map, batch 2, checkpoint
    (packed in_proj) key format."""
    import mmx_b200
    cfg = do.DETR_R50
    sd = do.init_state_dict(cfg, seed=9)
    src, pos, tq = do.synthetic_inputs(cfg, 2, 10, 12, seed=4)
    ref, stg = do.generate_ours(sd, cfg, src, pos, tq, return_stages=True)
    packed = {k: v for k, v in do.to_checkpoint_format(sd).items() if "q_proj" not in k and "k_proj" not in k and "v_proj" not in k}
    eng = mmx_b200.DetrEngine(packed, nhead=cfg.nhead, device="cuda:0")
    out = mmx_b200.Generator(eng).generate_ours((src.cuda(), pos.cuda()), tq, use_lrp=False)
    assert rel_err(eng.pred_logits, stg["logits"]) < TOL
    rec = eng.decoder[-1].multihead_attn["rec"]
    B, H = 2, cfg.nhead
    assert rel_err(rec.get_attn(), stg["A_dc"][-1].reshape(B, H, cfg.queries, -1)) < TOL
    assert rel_err(rec.get_attn_gradients(), stg["G_dc"][-1].reshape(B, H, cfg.queries, -1)) < TOL
    assert rel_err(out, ref) < TOL


@pytest.mark.parametrize("norm,s10", [(True, True), (False, True), (True, False)])
def test_lxmert_golden(golden_dir, norm, s10):
    import mmx_b200
    g = np.load(os.path.join(golden_dir, "lxmert_tiny.npz"))
    eng = mmx_b200.LxmertEngine(_sd(g), num_heads=lo.LXMERT_TINY.heads, device="cuda:0")
    gen = mmx_b200.GeneratorOurs(eng)
    ids, feats, boxes = (torch.from_numpy(g[k]) for k in ("ids", "feats", "boxes"))
    rtt, rti = gen.generate_ours((ids.cuda(), feats.cuda(), boxes.cuda()), use_lrp=False, normalize_self_attention=norm,
                                 apply_self_in_rule_10=s10)
    key = f"n{int(norm)}s{int(s10)}"
    assert rel_err(rtt, g["Rtt." + key]) < TOL and rel_err(rti, g["Rti." + key]) < TOL
    rtt1, rti1 = gen.generate_ours((ids[:1].cuda(), feats[:1].cuda(), boxes[:1].cuda()), use_lrp=False,
                                   normalize_self_attention=norm, apply_self_in_rule_10=s10)
    assert rtt1.shape == (ids.shape[1], ids.shape[1]) and rti1.shape == (ids.shape[1], feats.shape[1])   # reference shapes
    assert rel_err(rtt1, g["Rtt." + key][0]) < TOL and rel_err(rti1, g["Rti." + key][0]) < TOL
    assert gen.R_i_i.shape == (feats.shape[1], feats.shape[1])       # left on self like the reference
    # the last cross layer's image->text copy never receives a gradient
    assert eng.x_layers[-1].cross.recs[1].get_attn_gradients() is None
    assert gen.generate_ours((ids.cuda(), feats.cuda(), boxes.cuda()))[0].shape == rtt.shape   # default use_lrp=True runs


def test_lxmert_base_shape_vs_oracle():
    """LXMERT base dims (768 hidden, 12 heads, 9/5/5 layers), 20 tokens x 36 boxes, batch 2 (BASELINE.json config 4)."""
    import mmx_b200
    cfg = lo.LxmertConfig(vocab=2000, num_labels=300)
    sd = lo.init_state_dict(cfg, seed=1)
    ids, feats, boxes = lo.synthetic_inputs(cfg, 2, 20, 36, seed=8)
    ott, oti, logits = lo.generate_ours(sd, cfg, ids, feats, boxes)
    eng = mmx_b200.LxmertEngine(sd, num_heads=cfg.heads, device="cuda:0")
    rtt, rti = mmx_b200.GeneratorOurs(eng).generate_ours((ids.cuda(), feats.cuda(), boxes.cuda()), use_lrp=False)
    assert rel_err(eng.question_answering_score, logits) < TOL
    assert rel_err(rtt, ott) < TOL and rel_err(rti, oti) < TOL
    assert (rtt[:, 0, 0] == 0).all()


@pytest.mark.parametrize("method", ["raw_attn", "rollout", "attn_gradcam"])
def test_detr_baselines_golden(golden_dir, method):
    """Baseline methods behind the same API vs the reference Generator's own outputs."""
    import mmx_b200
    g = np.load(os.path.join(golden_dir, "detr_tiny.npz"))
    eng = mmx_b200.DetrEngine(_sd(g), nhead=do.DETR_TINY.nhead, device="cuda:0")
    gen = mmx_b200.Generator(eng)
    src, pos, tq = (torch.from_numpy(g[k]) for k in ("src", "pos", "tq"))
    out = getattr(gen, "generate_" + method)((src.cuda(), pos.cuda()), tq)
    assert rel_err(out, g["base." + method]) < TOL


@pytest.mark.parametrize("method", ["raw_attn", "rollout", "attn_gradcam"])
def test_lxmert_baselines_golden(golden_dir, method):
    import mmx_b200
    g = np.load(os.path.join(golden_dir, "lxmert_tiny.npz"))
    eng = mmx_b200.LxmertEngine(_sd(g), num_heads=lo.LXMERT_TINY.heads, device="cuda:0")
    gen = mmx_b200.GeneratorBaselines(eng)
    ids, feats, boxes = (torch.from_numpy(g[k]) for k in ("ids", "feats", "boxes"))
    rtt, rti = getattr(gen, "generate_" + method)((ids.cuda(), feats.cuda(), boxes.cuda()))
    assert rel_err(rtt, g[f"base.{method}.Rtt"]) < TOL and rel_err(rti, g[f"base.{method}.Rti"]) < TOL


def _vb_inputs(g):
    return {k[4:]: torch.from_numpy(g[k]).cuda() for k in g.files if k.startswith("inp.")}


@pytest.mark.parametrize("method", ["ours", "ours.index5", "raw_attn", "rollout", "rollout.sl1", "attn_gradcam"])
def test_visualbert_golden(golden_dir, method):
    """VisualBERT single-stream generator vs the reference SelfAttentionGenerator's own outputs."""
    import mmx_b200
    from oracle import visualbert_oracle as vo
    g = np.load(os.path.join(golden_dir, "visualbert_tiny.npz"))
    eng = mmx_b200.VisualBertEngine(_sd(g), num_heads=vo.VISUALBERT_TINY.heads, device="cuda:0")
    gen = mmx_b200.SelfAttentionGenerator(eng)
    inp = _vb_inputs(g)
    if method.startswith("ours"):
        out = gen.generate_ours(inp, index=5 if method.endswith("index5") else None)
    elif method.startswith("rollout"):
        out = gen.generate_rollout(inp, start_layer=1 if method.endswith("sl1") else 0)
    else:
        out = getattr(gen, "generate_" + method)(inp)
    assert out.shape == g["R." + method].shape
    assert rel_err(out, g["R." + method]) < TOL
    one = {k: v[:1] for k, v in inp.items()}                         # the reference's call form: one sample -> [1, S]
    if method == "ours":
        assert rel_err(gen.generate_ours(one), g["R.ours"][:1]) < TOL


def test_visualbert_base_shape_vs_oracle():
    """VisualBERT base dims (768 hidden, 12 heads, 12 layers), 14 tokens + 100 boxes of 2048-d features, mmf-style
    ``model.``-prefixed checkpoint keys."""
    import mmx_b200
    from oracle import visualbert_oracle as vo
    cfg = vo.VisualBertConfig(vocab=2000, num_labels=300)
    sd = vo.init_state_dict(cfg, seed=2)
    inp = vo.synthetic_inputs(cfg, 2, 14, 100, seed=3)
    ref, scores = vo.generate_ours(sd, cfg, inp)
    eng = mmx_b200.VisualBertEngine({"model." + k: v for k, v in sd.items()}, num_heads=cfg.heads, device="cuda:0")
    out = mmx_b200.SelfAttentionGenerator(eng).generate_ours({k: v.cuda() for k, v in inp.items()})
    assert rel_err(eng.scores, scores) < TOL
    assert rel_err(out, ref) < TOL
    assert (out[torch.arange(2), eng.cls_index] == 0).all()


def test_minmax_normalize():
    import mmx_b200
    x = torch.randn(3, 17, 23, device="cuda")
    ref = torch.stack([(m - m.min()) / (m.max() - m.min()) for m in x])
    assert torch.equal(mmx_b200.minmax_normalize(x), ref)


# The no-aggregation ablations chain ~12 un-aggregated products of small matrices (the map is ~1e-8 .. 1e-12 at random
# init), so the relative errors of the factors multiply up; 5e-4 on these two ablation-only outputs.
ABL_TOL = 5e-4


def test_detr_ablation_no_agg_golden(golden_dir):
    """GeneratorAlbationNoAgg.generate_ours_abl vs the reference class's own outputs."""
    import mmx_b200
    g = np.load(os.path.join(golden_dir, "detr_tiny.npz"))
    eng = mmx_b200.DetrEngine(_sd(g), nhead=do.DETR_TINY.nhead, device="cuda:0")
    gen = mmx_b200.GeneratorAlbationNoAgg(eng)
    src, pos, tq = (torch.from_numpy(g[k]) for k in ("src", "pos", "tq"))
    out = gen.generate_ours_abl((src.cuda(), pos.cuda()), tq)
    assert rel_err(out, g["abl.noagg"]) < ABL_TOL
    out = gen.generate_ours_abl((src.cuda(), pos.cuda()), tq, apply_self_in_rule_10=False)
    assert rel_err(out, g["abl.noagg.s0"]) < ABL_TOL
    with pytest.raises(AssertionError):                   # like the reference: eq. 8-9 asserts on the un-aggregated R
        gen.generate_ours_abl((src.cuda(), pos.cuda()), tq, normalize_self_attention=True)


def test_lxmert_ablation_no_agg_golden(golden_dir):
    import mmx_b200
    g = np.load(os.path.join(golden_dir, "lxmert_tiny.npz"))
    eng = mmx_b200.LxmertEngine(_sd(g), num_heads=lo.LXMERT_TINY.heads, device="cuda:0")
    gen = mmx_b200.GeneratorOursAblationNoAggregation(eng)
    ids, feats, boxes = (torch.from_numpy(g[k]).cuda() for k in ("ids", "feats", "boxes"))
    rtt, rti = gen.generate_ours_no_agg((ids, feats, boxes), normalize_self_attention=False)
    assert rel_err(rtt, g["abl.noagg.Rtt"]) < ABL_TOL and rel_err(rti, g["abl.noagg.Rti"]) < ABL_TOL
    with pytest.raises(AssertionError):
        gen.generate_ours_no_agg((ids, feats, boxes))   # default normalize_self_attention=True asserts in the reference too


def test_otsu_masks_bit_exact(golden_dir):
    """mmx_otsu_masks vs cv2.threshold(THRESH_BINARY + THRESH_OTSU) outputs (golden) - integer / byte work: bit-exact."""
    import mmx_b200
    g = np.load(os.path.join(golden_dir, "otsu.npz"))
    masks, thr = mmx_b200.otsu_masks(torch.from_numpy(g["cams"]).cuda())
    assert np.array_equal(thr.cpu().numpy(), g["thresholds"])
    assert np.array_equal(masks.cpu().numpy(), g["masks"])
    # ragged sizes and a large map against the oracle
    from oracle import rules as R
    gen = torch.Generator().manual_seed(3)
    for n in (1, 2, 7, 255, 256, 850, 20000):
        x = torch.rand(5, n, generator=gen) ** 3
        if n == 1:
            continue                                        # a constant map is 0/0 in the reference too
        m, t = mmx_b200.otsu_masks(x.cuda())
        om, ot = R.otsu_masks(x)
        assert np.array_equal(t.cpu().numpy(), ot.numpy()) and np.array_equal(m.cpu().numpy(), om.numpy()), n


@pytest.mark.parametrize("method", ["ours_no_lrp", "ablation_no_self_in_10", "raw_attn", "rollout"])
def test_detr_mask_generator(golden_dir, method):
    """MaskGenerator.get_masks: all kept queries of one image in one batch == the per-query reference maps (golden) pushed
    through the oracle's Otsu step."""
    import mmx_b200
    from oracle import rules as R
    g = np.load(os.path.join(golden_dir, "detr_tiny.npz"))
    eng = mmx_b200.DetrEngine(_sd(g), nhead=do.DETR_TINY.nhead, device="cuda:0")
    mg = mmx_b200.MaskGenerator(eng)
    src, pos = torch.from_numpy(g["src"])[:1], torch.from_numpy(g["pos"])[:1]
    queries = torch.tensor([0, 2, 3])
    masks, thr, cams = mg.get_masks((src.cuda(), pos.cuda()), queries, method)
    assert masks.shape == (3, src.shape[-2], src.shape[-1])
    # per-query oracle maps of the same image
    kw = {"ours_no_lrp": {}, "ablation_no_self_in_10": {"apply_self_in_rule_10": False}}
    if method in kw:
        ref = do.generate_ours(_sd(g), do.DETR_TINY, src.expand(3, -1, -1, -1), pos.expand(3, -1, -1, -1), queries, **kw[method])
        assert rel_err(cams, ref) < TOL
    om, ot = R.otsu_masks(cams.cpu())
    assert np.array_equal(thr.cpu().numpy(), ot.numpy())
    assert np.array_equal(masks.reshape(3, -1).cpu().numpy(), om.numpy())
    assert mg.get_masks((src.cuda(), pos.cuda()), queries, "no_such_method") is None
    for lrp_method in ("ours_with_lrp", "transformer_att", "partial_lrp"):       # the LRP-based switches of mask_generator.py:93-110
        m2, _, c2 = mg.get_masks((src.cuda(), pos.cuda()), queries, lrp_method)
        assert m2.shape == masks.shape and torch.isfinite(c2).all()
