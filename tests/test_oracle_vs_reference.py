"""CPU, build container only: the oracle restatement against the LIVE reference imported from /root/reference
(skipped on the GPU box, where the reference tree does not exist)."""
import pytest
import torch

from oracle import clip_oracle as co
from oracle import ref_shims as rs
from util import rel_err, text_rel_err

pytestmark = pytest.mark.skipif(not rs.reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("sl", [-1, 0])
def test_vit_b32_full_size(sl):
    cfg = co.VIT_B32
    sd = co.init_state_dict(cfg, seed=0)
    images, tokens = co.synthetic_inputs(cfg, 2, seed=7)
    model = rs.build_reference_clip(cfg, sd)
    rts, ris = [], []
    for b in range(2):
        rt, ri = rs.reference_interpret(images[b:b + 1], tokens[b:b + 1], model, "cpu", sl, sl)
        rts.append(rt.detach()); ris.append(ri.detach())
    ot, oi = co.clip_interpret(sd, cfg, images, tokens, sl, sl)
    assert text_rel_err(ot, torch.cat(rts)) < 2e-5
    assert rel_err(oi, torch.cat(ris)) < 2e-5


def test_repeat_mode_small():
    cfg = co.SMALL
    sd = co.init_state_dict(cfg, seed=4)
    images, tokens = co.synthetic_inputs(cfg, 5, seed=9)
    model = rs.build_reference_clip(cfg, sd)
    rt, ri = rs.reference_interpret(images[:1], tokens, model, "cpu", 0, 0)
    ot, oi = co.clip_interpret(sd, cfg, images[:1], tokens, 0, 0)
    assert text_rel_err(ot, rt.detach()) < 1e-5 and rel_err(oi, ri.detach()) < 1e-5


@pytest.mark.parametrize("method", ["ours", "raw_attn", "rollout", "attn_gradcam"])
def test_visualbert_live_reference(method):
    """The unmodified BertEncoder / BertPredictionHeadTransform / SelfAttentionGenerator of the reference at a
    wider shape than the golden file (hidden 128, 8 heads, 4 layers, 9 tokens + 12 boxes)."""
    from oracle import visualbert_oracle as vo, ref_visualbert as rv
    cfg = vo.VisualBertConfig(hidden=128, heads=8, intermediate=256, layers=4, vocab=80, max_pos=32, visual_dim=48, num_labels=31)
    sd = vo.init_state_dict(cfg, 11)
    inp = vo.synthetic_inputs(cfg, 2, 9, 12, seed=5)
    ref = rv.generate(cfg, sd, inp, method)
    got = getattr(vo, "generate_" + method)(sd, cfg, inp)
    got = got[0] if isinstance(got, tuple) else got
    assert rel_err(got, ref) < 1e-5


def test_detr_lrp_live_reference_wider():
    """use_lrp=True at a shape other than the golden one (d=128, 4 heads, 3 encoder / 3 decoder layers, 5x6 features,
    9 queries): the restated relprop sweep against the unmodified reference generator."""
    from oracle import detr_oracle as do, ref_detr
    cfg = do.DetrConfig(128, 4, 3, 3, 192, 9, 6)
    sd = do.init_state_dict(cfg, 9)
    src, pos, tq = do.synthetic_inputs(cfg, 2, 5, 6, seed=4)
    ref = ref_detr.generate_ours(cfg, sd, src, pos, tq, use_lrp=True, normalize_self_attention=False)
    got = do.generate_ours_lrp(sd, cfg, src, pos, tq, normalize_self_attention=False)
    assert ref.abs().max() > 0
    assert rel_err(got, ref) < 2e-4


def test_lxmert_lrp_live_reference_wider():
    from oracle import lxmert_oracle as lo, ref_lxmert
    cfg = lo.LxmertConfig(hidden=128, heads=4, intermediate=160, l_layers=3, x_layers=3, r_layers=2, vocab=70, max_pos=20,
                          feat_dim=40, pos_dim=4, num_labels=17)
    sd = lo.init_state_dict(cfg, 6)
    ids, feats, boxes = lo.synthetic_inputs(cfg, 2, 9, 7, seed=3)
    rtt, rti = ref_lxmert.generate_ours(cfg, sd, ids, feats, boxes, use_lrp=True)
    gtt, gti = lo.generate_ours_lrp(sd, cfg, ids, feats, boxes)
    assert rel_err(gtt, rtt) < 1e-5 and rel_err(gti, rti) < 1e-5
