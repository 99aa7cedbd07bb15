"""CPU: the DETR and LXMERT oracle restatements reproduce the committed outputs of the UNMODIFIED reference
generators (tests/golden/detr_tiny.npz, lxmert_tiny.npz, visualbert_tiny.npz made by oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import detr_oracle as do, lxmert_oracle as lo, visualbert_oracle as vo
from util import rel_err


def _sd(g):
    return {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}


@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("s10", [True, False])
def test_detr_oracle_golden(golden_dir, norm, s10):
    g = np.load(os.path.join(golden_dir, "detr_tiny.npz"))
    r = do.generate_ours(_sd(g), do.DETR_TINY, torch.from_numpy(g["src"]), torch.from_numpy(g["pos"]), torch.from_numpy(g["tq"]),
                         normalize_self_attention=norm, apply_self_in_rule_10=s10)
    assert rel_err(r, g[f"R.n{int(norm)}s{int(s10)}"]) < 1e-5


@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("s10", [True, False])
def test_lxmert_oracle_golden(golden_dir, norm, s10):
    g = np.load(os.path.join(golden_dir, "lxmert_tiny.npz"))
    rtt, rti, _ = lo.generate_ours(_sd(g), lo.LXMERT_TINY, torch.from_numpy(g["ids"]), torch.from_numpy(g["feats"]),
                                   torch.from_numpy(g["boxes"]), normalize_self_attention=norm, apply_self_in_rule_10=s10)
    key = f"n{int(norm)}s{int(s10)}"
    assert rel_err(rtt, g["Rtt." + key]) < 1e-5 and rel_err(rti, g["Rti." + key]) < 1e-5


def _vb_inputs(g):
    return {k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("inp.")}


@pytest.mark.parametrize("method", ["ours", "ours.index5", "raw_attn", "rollout", "rollout.sl1", "attn_gradcam"])
def test_visualbert_oracle_golden(golden_dir, method):
    g = np.load(os.path.join(golden_dir, "visualbert_tiny.npz"))
    sd, inp, cfg = _sd(g), _vb_inputs(g), vo.VISUALBERT_TINY
    if method.startswith("ours"):
        r = vo.generate_ours(sd, cfg, inp, index=5 if method.endswith("index5") else None)[0]
    elif method.startswith("rollout"):
        r = vo.generate_rollout(sd, cfg, inp, start_layer=1 if method.endswith("sl1") else 0)
    else:
        r = getattr(vo, "generate_" + method)(sd, cfg, inp)
    assert rel_err(r, g["R." + method]) < 1e-5


def test_detr_ablation_no_agg_oracle_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "detr_tiny.npz"))
    args = (_sd(g), do.DETR_TINY, torch.from_numpy(g["src"]), torch.from_numpy(g["pos"]), torch.from_numpy(g["tq"]))
    assert rel_err(do.generate_ours_abl(*args), g["abl.noagg"]) < 1e-5
    assert rel_err(do.generate_ours_abl(*args, apply_self_in_rule_10=False), g["abl.noagg.s0"]) < 1e-5


def test_lxmert_ablation_no_agg_oracle_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "lxmert_tiny.npz"))
    rtt, rti = lo.generate_ours_no_agg(_sd(g), lo.LXMERT_TINY, torch.from_numpy(g["ids"]), torch.from_numpy(g["feats"]),
                                       torch.from_numpy(g["boxes"]), normalize_self_attention=False)
    assert rel_err(rtt, g["abl.noagg.Rtt"]) < 1e-5 and rel_err(rti, g["abl.noagg.Rti"]) < 1e-5


@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("s10", [True, False])
def test_detr_lrp_oracle_golden(golden_dir, norm, s10):
    """use_lrp=True: the restated relprop sweep (oracle/lrp.py) against the reference generator's own
    outputs.  fp32 through ~60 chained safe-divides: 1e-4 relative."""
    g = np.load(os.path.join(golden_dir, "detr_tiny.npz"))
    r = do.generate_ours_lrp(_sd(g), do.DETR_TINY, torch.from_numpy(g["src"]), torch.from_numpy(g["pos"]), torch.from_numpy(g["tq"]),
                             normalize_self_attention=norm, apply_self_in_rule_10=s10)
    assert rel_err(r, g[f"R.lrp.n{int(norm)}s{int(s10)}"]) < 1e-4


def test_detr_lrp_attention_relevance_golden(golden_dir):
    """The per-layer LRP relevance of A (get_attn_cam) of sample 0.  Structure: float64 restatement vs the reference run
    in float64 (1e-10).  fp32 vs fp32 only agrees to ~3e-4 - the sweep divides by many small numbers."""
    from oracle import lrp
    g = np.load(os.path.join(golden_dir, "detr_tiny.npz"))
    src, pos, tq = torch.from_numpy(g["src"])[:1], torch.from_numpy(g["pos"])[:1], int(g["tq"][0])
    for dtype, key, tol in ((torch.float64, "lrp.cam64.", 1e-10), (torch.float32, "lrp.cam.", 2e-3)):
        sd = {k: v.to(dtype) for k, v in _sd(g).items()}
        with torch.no_grad():
            logits = do.detr_forward(sd, do.DETR_TINY, src.to(dtype), pos.to(dtype))[0]
            cls = int(logits[0, tq, :-1].argmax())                      # the generator's class choice (EG:151-152)
            _, enc, dec = lrp.detr_lrp_sweep(sd, do.DETR_TINY, src.to(dtype), pos.to(dtype), tq, cls)
        for i, layer in enumerate(enc):
            assert rel_err(layer.attn.attn_cam, g[f"{key}enc{i}"]) < tol
        for i, layer in enumerate(dec):
            assert rel_err(layer.self_attn.attn_cam, g[f"{key}dec{i}.self"]) < tol
            assert rel_err(layer.cross.attn_cam, g[f"{key}dec{i}.cross"]) < tol


@pytest.mark.parametrize("norm", [True, False])
@pytest.mark.parametrize("s10", [True, False])
def test_lxmert_lrp_oracle_golden(golden_dir, norm, s10):
    """use_lrp=True for LXMERT (the reference default of GeneratorOurs.generate_ours): oracle/lrp.py vs the reference."""
    g = np.load(os.path.join(golden_dir, "lxmert_tiny.npz"))
    rtt, rti = lo.generate_ours_lrp(_sd(g), lo.LXMERT_TINY, torch.from_numpy(g["ids"]), torch.from_numpy(g["feats"]),
                                    torch.from_numpy(g["boxes"]), normalize_self_attention=norm, apply_self_in_rule_10=s10)
    key = f"lrp.n{int(norm)}s{int(s10)}"
    assert rel_err(rtt, g["Rtt." + key]) < 1e-5 and rel_err(rti, g["Rti." + key]) < 1e-5


@pytest.mark.parametrize("method", ["transformer_att", "transformer_att.sl1", "partial_lrp"])
def test_visualbert_lrp_oracle_golden(golden_dir, method):
    """The LRP-based VisualBERT methods (generate_transformer_att, generate_partial_lrp) vs the reference generator."""
    g = np.load(os.path.join(golden_dir, "visualbert_tiny.npz"))
    sd, inp, cfg = _sd(g), _vb_inputs(g), vo.VISUALBERT_TINY
    if method.startswith("transformer_att"):
        r = vo.generate_transformer_att(sd, cfg, inp, start_layer=1 if method.endswith("sl1") else 0)
    else:
        r = vo.generate_partial_lrp(sd, cfg, inp)
    assert rel_err(r, g["R." + method]) < 1e-5


@pytest.mark.parametrize("method", ["transformer_att", "partial_lrp"])
def test_detr_lrp_baselines_oracle_golden(golden_dir, method):
    g = np.load(os.path.join(golden_dir, "detr_tiny.npz"))
    r = do.generate_lrp_baseline(_sd(g), do.DETR_TINY, torch.from_numpy(g["src"]), torch.from_numpy(g["pos"]),
                                 torch.from_numpy(g["tq"]), method)
    assert rel_err(r, g["base." + method]) < 2e-4          # fp32 conditioning of the relprop sweep, see oracle/lrp.py


@pytest.mark.parametrize("method", ["transformer_attr", "partial_lrp"])
def test_lxmert_lrp_baselines_oracle_golden(golden_dir, method):
    g = np.load(os.path.join(golden_dir, "lxmert_tiny.npz"))
    rtt, rti = lo.generate_lrp_baseline(_sd(g), lo.LXMERT_TINY, torch.from_numpy(g["ids"]), torch.from_numpy(g["feats"]),
                                        torch.from_numpy(g["boxes"]), method)
    assert rel_err(rtt, g[f"base.{method}.Rtt"]) < 1e-5 and rel_err(rti, g[f"base.{method}.Rti"]) < 1e-5


@pytest.mark.parametrize("positive", [False, True])
@pytest.mark.parametrize("modality", ["image", "text"])
def test_lxmert_perturbation_oracle_golden(golden_dir, modality, positive):
    """The perturbation loops: the oracle restatement against the per-step answer scores produced by the
    UNMODIFIED loops of lxmert/lxmert/perturbation.py:85-194 driving the unmodified reference LXMERT
    (oracle/ref_perturbation.py; golden written by oracle/make_golden.py --lxmert-perturbation)."""
    g = np.load(os.path.join(golden_dir, "lxmert_perturbation.npz"))
    sd = _sd(g)
    ids, feats, boxes = (torch.from_numpy(g[k]) for k in ("ids", "feats", "boxes"))
    assert list(g["pert_steps"]) == lo.PERT_STEPS
    fn = lo.perturbation_image if modality == "image" else lo.perturbation_text
    got = fn(sd, lo.LXMERT_TINY, ids, feats, boxes, torch.from_numpy(g["cam_" + modality]), positive)
    assert rel_err(got, g[f"scores.{modality}.{int(positive)}"]) < 1e-5
