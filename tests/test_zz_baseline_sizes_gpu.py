"""GPU: DETR and LXMERT end to end at the sizes BASELINE.json quotes (configs 3 and 4) against the oracle - the kernels are
held to these sizes one by one in test_primitives_gpu / test_rules_gpu, CLIP ViT-L/14@336 (config 5) in
test_clip_gpu.test_vit_l14_336_vs_oracle; this file runs the whole generators there.

Criterion.  LXMERT is well conditioned at this size (the fp32 oracle is 7e-6 from the fp64 oracle) and is held to the
north-star 1e-4 against the fp32 oracle.  DETR is not: the reference's own fp32 arithmetic is 3-4e-5 away from the same
algorithm in fp64 at 120 tokens and at 625 / 850 alike (random-init weights; measured on the build box), i.e. a third of the
1e-4 budget is the reference's own rounding noise.  The DETR maps are therefore held to the criterion of the LRP tests:
err(device, fp64 oracle) <= max(1e-4, 5 x err(fp32 oracle, fp64 oracle)); the class logits (well conditioned) to 1e-4.
(Sorted last on purpose: the file was added when the round's GPU minutes were spent; it mirrors tests that pass at the
small sizes.)"""
import pytest
import torch

from oracle import detr_oracle as do, lxmert_oracle as lo
from util import rel_err, TOL

pytestmark = pytest.mark.gpu
NOISE_FACTOR = 5.0


def _detr_case(B, h, w, seed):
    import mmx_b200
    cfg = do.DETR_R50
    sd = do.init_state_dict(cfg, seed=9)
    src, pos, tq = do.synthetic_inputs(cfg, B, h, w, seed=seed)
    r32, stg = do.generate_ours(sd, cfg, src, pos, tq, return_stages=True)
    r64 = do.generate_ours(sd, cfg, src, pos, tq, dtype=torch.float64)
    packed = {k: v for k, v in do.to_checkpoint_format(sd).items() if "q_proj" not in k and "k_proj" not in k and "v_proj" not in k}
    eng = mmx_b200.DetrEngine(packed, nhead=cfg.nhead, device="cuda:0")
    out = mmx_b200.Generator(eng).generate_ours((src.cuda(), pos.cuda()), tq, use_lrp=False)
    noise = rel_err(r32, r64)
    e64, e32, elog = rel_err(out, r64), rel_err(out, r32), rel_err(eng.pred_logits, stg["logits"])
    print(f"DETR-R50 {h}x{w} B={B}: maps vs fp64 oracle {e64:.2e}, vs fp32 oracle {e32:.2e} (fp32 oracle vs fp64 {noise:.2e}); logits {elog:.2e}")
    assert tuple(out.shape) == (B, h * w)
    assert elog < TOL
    assert e64 < max(TOL, NOISE_FACTOR * noise)


def test_detr_r50_config3_size():
    """BASELINE.json config 3: 25 x 25 backbone features (625 tokens), 100 queries, batch 16."""
    _detr_case(16, 25, 25, seed=4)


def test_detr_r50_850_tokens():
    """The 800 x 1066 input of the reference's evaluation: 25 x 34 features = 850 tokens (the 32-query attention tiles)."""
    _detr_case(2, 25, 34, seed=5)


def test_lxmert_config4_size():
    """BASELINE.json config 4: LXMERT-base, 20 tokens x 36 boxes, batch 32."""
    import mmx_b200
    cfg = lo.LxmertConfig(vocab=2000, num_labels=300)
    sd = lo.init_state_dict(cfg, seed=1)
    ids, feats, boxes = lo.synthetic_inputs(cfg, 32, 20, 36, seed=8)
    ott, oti, logits = lo.generate_ours(sd, cfg, ids, feats, boxes)
    eng = mmx_b200.LxmertEngine(sd, num_heads=cfg.heads, device="cuda:0")
    rtt, rti = mmx_b200.GeneratorOurs(eng).generate_ours((ids.cuda(), feats.cuda(), boxes.cuda()), use_lrp=False)
    e = (rel_err(eng.question_answering_score, logits), rel_err(rtt, ott), rel_err(rti, oti))
    print("LXMERT-base B=32:", " ".join(f"{x:.2e}" for x in e))
    assert max(e) < TOL
    assert (rtt[:, 0, 0] == 0).all()
