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
