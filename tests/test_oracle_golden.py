"""CPU: the oracle restatement reproduces the committed outputs of the UNMODIFIED reference (tests/golden/, made by
oracle/make_golden.py).  This is what pins the oracle - the reference has no tests of its own for this path."""
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle as co
from oracle import rules as orules
from util import rel_err, text_rel_err


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("tag", ["tiny", "small"])
@pytest.mark.parametrize("sl", [-1, 0, 1])
def test_clip_oracle_matches_reference_golden(golden_dir, tag, sl):
    g = _load(golden_dir, f"clip_{tag}.npz")
    cfg = co.ClipConfig(*[int(v) for v in g["cfg"]])
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    images, tokens = torch.from_numpy(g["images"]), torch.from_numpy(g["tokens"])
    rt, ri = co.clip_interpret(sd, cfg, images, tokens, sl, sl)
    assert text_rel_err(rt, g[f"distinct.sl{sl}.R_text"]) < 1e-5
    assert rel_err(ri, g[f"distinct.sl{sl}.R_image"]) < 1e-5
    rt, ri = co.clip_interpret(sd, cfg, images[:1], tokens, sl, sl)      # the notebook's image.repeat mode
    assert text_rel_err(rt, g[f"repeat.sl{sl}.R_text"]) < 1e-5
    assert rel_err(ri, g[f"repeat.sl{sl}.R_image"]) < 1e-5


def test_clip_oracle_logits_golden(golden_dir):
    g = _load(golden_dir, "clip_small.npz")
    cfg = co.ClipConfig(*[int(v) for v in g["cfg"]])
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    with torch.no_grad():
        logits = co.clip_forward(sd, cfg, torch.from_numpy(g["images"]), torch.from_numpy(g["tokens"]))[0]
    assert rel_err(logits, g["logits_per_image"]) < 1e-5


def test_rules_oracle_matches_reference_golden(golden_dir):
    g = {k: torch.from_numpy(v) for k, v in _load(golden_dir, "rules.npz").items()}
    assert rel_err(orules.avg_heads(g["cam_ss"], g["grad_ss"]), g["abar_ss"]) < 1e-6
    assert rel_err(orules.avg_heads(g["cam_sq"], g["grad_sq"]), g["abar_sq"]) < 1e-6
    a, b = orules.apply_self_attention_rules(g["R_ss"], g["R_sq"], g["abar_ss"])
    assert rel_err(a, g["self.R_ss_add"]) < 1e-6 and rel_err(b, g["self.R_sq_add"]) < 1e-6
    assert rel_err(orules.handle_residual(g["R_ss"]), g["hr.R_ss"]) < 1e-6
    assert rel_err(orules.handle_residual(g["R_qq"]), g["hr.R_qq"]) < 1e-6
    for norm in (True, False):
        for s10 in (True, False):
            key = f"n{int(norm)}s{int(s10)}"
            d = orules.apply_mm_attention_rules_detr(g["R_ss"], g["R_qq"], g["abar_sq"], norm, s10)
            assert rel_err(d, g[f"mm_detr.{key}"]) < 1e-6
            x, y = orules.apply_mm_attention_rules_lxmert(g["R_ss"], g["R_qq"], g["R_qs"], g["abar_sq"], norm, s10)
            assert rel_err(x, g[f"mm_lx.{key}.sq"]) < 1e-6 and rel_err(y, g[f"mm_lx.{key}.ss"]) < 1e-6
    T, S = g["R_ss"].shape[0], g["R_qq"].shape[0]
    ident = orules.apply_mm_attention_rules_detr(torch.eye(T), torch.eye(S), g["abar_sq"])
    assert torch.equal(ident, g["mm_detr.identity_state"])          # 0/0 -> NaN -> 0 (DETR guard)
    mats = list(g["rollout.mats"])
    for sl in (0, 2):
        assert rel_err(orules.compute_rollout_attention(mats, sl), g[f"rollout.sl{sl}"]) < 1e-6


def test_handle_residual_asserts_like_reference():
    R = torch.eye(4)
    R[1, 1] = 0.5           # diag(R - I) < 0
    with pytest.raises(AssertionError):
        orules.handle_residual(R)


def test_otsu_oracle_matches_cv2_golden(golden_dir):
    """The restated OpenCV Otsu + the mask expression of DETR/mask_generator.py:115-121 against cv2.threshold's own
    outputs (tests/golden/otsu.npz, made with the cv2 of the build image), bit for bit."""
    import os
    import numpy as np
    import torch
    from oracle import rules as R
    g = np.load(os.path.join(golden_dir, "otsu.npz"))
    masks, ths = R.otsu_masks(torch.from_numpy(g["cams"]))
    assert np.array_equal(ths.numpy(), g["thresholds"])
    assert np.array_equal(masks.numpy(), g["masks"])


@pytest.mark.parametrize("index", [None, 0, 1, 2])
def test_clip_oracle_matches_example_py_golden(golden_dir, index):
    """The older unbatched interpret of CLIP/example.py:8-53 (one image, N texts, ``index`` picks the text, all image
    blocks): its relevance for text ``index`` is the notebook rule with start_layer = 0 on the pair (image, text[index])."""
    g = _load(golden_dir, "clip_example.npz")
    cfg = co.ClipConfig(*[int(v) for v in g["cfg"]])
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    image, tokens = torch.from_numpy(g["image"]), torch.from_numpy(g["tokens"])
    n = int(np.argmax(g["logits_per_image"][0])) if index is None else index
    _, ri = co.clip_interpret(sd, cfg, image, tokens[n:n + 1], 0, cfg.transformer_layers - 1)
    assert rel_err(ri[0], g[f"R.index{index}"]) < 1e-5
    with torch.no_grad():
        logits = co.clip_forward(sd, cfg, image.expand(tokens.shape[0], -1, -1, -1), tokens)[0]
    assert rel_err(logits[:1], g["logits_per_image"]) < 1e-5
