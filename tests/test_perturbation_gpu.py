"""GPU: the LXMERT perturbation driver - all steps of an item in one masked / compacted batch must give
the scores of the reference's per-step gathered forwards (oracle restatement of lxmert/lxmert/perturbation.py:85-194)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import lxmert_oracle as lo
from util import rel_err, TOL

pytestmark = pytest.mark.gpu


def test_topk_select_matches_topk():
    import mmx_b200
    gen = torch.Generator().manual_seed(0)
    for n in (1, 5, 36, 100, 1000):
        s = torch.randn(n, generator=gen)
        ks = [0, 1, n // 3, n - 1, n] if n > 1 else [0, 1]
        keep, pos = mmx_b200.topk_select(s.cuda(), ks)
        for r, k in enumerate(ks):
            want = torch.zeros(n, dtype=torch.int32)
            want[s.topk(k).indices] = 1
            assert torch.equal(keep[r].cpu(), want), (n, k)
            kept = want.bool()
            assert torch.equal(pos[r].cpu()[kept], torch.arange(int(kept.sum()), dtype=torch.int32))
            assert (pos[r].cpu()[~kept] == -1).all()
    # ties: lower index first
    keep, _ = mmx_b200.topk_select(torch.tensor([1.0, 2.0, 2.0, 2.0, 0.5]).cuda(), [2])
    assert keep[0].tolist() == [0, 1, 1, 0, 0]


def test_attention_removed_keys():
    """-inf key bias == the key is not there; a row with every key removed is the empty softmax: A = 0, O = 0."""
    import mmx_b200
    from mmx_b200._lib import lib, check, ptr, current_stream
    B, H, T, S, hd = 2, 2, 5, 7, 16
    D = H * hd
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(B, n, D, generator=g).cuda() for n in (T, S, S))
    bias = torch.zeros(B, S)
    bias[0, [1, 4]] = float("-inf")
    bias[1, :] = float("-inf")
    A = torch.empty(B, H, T, 8, device="cuda"); O = torch.empty(B, T, D, device="cuda")
    check(lib().mmx_attention_fwd(ptr(q), D, ptr(k), D, ptr(v), D, ptr(bias.cuda()), ptr(A), 8, ptr(O), D, B, H, T, S, hd,
                                  C.c_float(1 / math.sqrt(hd)), 2, current_stream()))
    keep = [0, 2, 3, 5, 6]
    qh = q[0].view(T, H, hd).transpose(0, 1).double().cpu(); kh = k[0, keep].view(5, H, hd).transpose(0, 1).double().cpu()
    vh = v[0, keep].view(5, H, hd).transpose(0, 1).double().cpu()
    P = (qh @ kh.transpose(1, 2) / math.sqrt(hd)).softmax(-1)
    assert rel_err(A[0][..., keep], P) < 5e-6 and (A[0][..., [1, 4]] == 0).all()
    assert rel_err(O[0], (P @ vh).transpose(0, 1).reshape(T, D)) < 5e-6
    assert (A[1] == 0).all() and (O[1] == 0).all()


@pytest.mark.parametrize("positive", [False, True])
@pytest.mark.parametrize("modality", ["image", "text"])
def test_lxmert_perturbation_vs_oracle(modality, positive):
    import mmx_b200
    cfg = lo.LXMERT_TINY
    sd = lo.init_state_dict(cfg, 3)
    ids, feats, boxes = lo.synthetic_inputs(cfg, 1, 9, 11, seed=6)
    eng = mmx_b200.LxmertEngine(sd, num_heads=cfg.heads, device="cuda:0")
    item = (ids.cuda(), feats.cuda(), boxes.cuda())
    R_t_t, R_t_i = mmx_b200.GeneratorOurs(eng).generate_ours(item, use_lrp=False)
    cam_image, cam_text = R_t_i[0], R_t_t[0]                                   # perturbation.py:241-244
    cam_image = mmx_b200.minmax_normalize(cam_image.reshape(1, -1))[0]
    cam_text = mmx_b200.minmax_normalize(cam_text.reshape(1, -1))[0]
    pert = mmx_b200.LxmertPerturbation(eng)
    assert pert.pert_steps == lo.PERT_STEPS
    if modality == "image":
        ans = pert.perturbation_image(item, cam_image, cam_text, positive)
        ref = lo.perturbation_image(sd, cfg, ids, feats, boxes, cam_image.cpu(), positive)
    else:
        ans = pert.perturbation_text(item, cam_image, cam_text, positive)
        ref = lo.perturbation_text(sd, cfg, ids, feats, boxes, cam_text.cpu(), positive)
    assert pert.scores.shape == ref.shape == (len(lo.PERT_STEPS), cfg.num_labels)
    assert torch.isfinite(pert.scores).all()
    assert rel_err(pert.scores, ref) < TOL
    assert torch.equal(ans.cpu(), ref.argmax(-1))


@pytest.mark.parametrize("positive", [False, True])
@pytest.mark.parametrize("modality", ["image", "text"])
def test_lxmert_perturbation_reference_golden(golden_dir, modality, positive):
    """Against the per-step answer scores of the UNMODIFIED reference loops (lxmert/lxmert/perturbation.py:85-194 driving the
    unmodified reference LXMERT; tests/golden/lxmert_perturbation.npz, oracle/ref_perturbation.py)."""
    import os
    import mmx_b200
    g = np.load(os.path.join(golden_dir, "lxmert_perturbation.npz"))
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    eng = mmx_b200.LxmertEngine(sd, num_heads=lo.LXMERT_TINY.heads, device="cuda:0")
    item = tuple(torch.from_numpy(g[k]).cuda() for k in ("ids", "feats", "boxes"))
    cam_image, cam_text = torch.from_numpy(g["cam_image"]).cuda(), torch.from_numpy(g["cam_text"]).cuda()
    pert = mmx_b200.LxmertPerturbation(eng)
    assert pert.pert_steps == list(g["pert_steps"])
    ans = getattr(pert, "perturbation_" + modality)(item, cam_image, cam_text, positive)
    ref = g[f"scores.{modality}.{int(positive)}"]
    assert rel_err(pert.scores, ref) < TOL
    assert torch.equal(ans.cpu(), torch.from_numpy(ref).argmax(-1))


@pytest.mark.parametrize("positive", [False, True])
@pytest.mark.parametrize("modality", ["image", "text"])
def test_visualbert_perturbation_vs_oracle(modality, positive):
    import mmx_b200
    from oracle import visualbert_oracle as vo
    cfg = vo.VISUALBERT_TINY
    sd = vo.init_state_dict(cfg, 3)
    inp = vo.synthetic_inputs(cfg, 1, 10, 9, seed=8)
    inp["attention_mask"][0, -1] = 0                                        # one padded box (image_mask)
    eng = mmx_b200.VisualBertEngine(sd, num_heads=cfg.heads, device="cuda:0")
    dinp = {k: v.cuda() for k, v in inp.items()}
    cam = mmx_b200.SelfAttentionGenerator(eng).generate_ours(dinp)           # [1, T+V]
    pert = mmx_b200.VisualBertPerturbation(eng)
    fn = "perturbation_" + modality
    ans = getattr(pert, fn)(dinp, cam, positive)
    ref = getattr(vo, fn)(sd, cfg, inp, cam.cpu(), positive)
    steps = vo.PERT_STEPS_IMAGE if modality == "image" else vo.PERT_STEPS          # evaluation_loop.py:93-96
    assert pert.steps_for(modality) == steps
    assert pert.scores.shape == ref.shape == (len(steps), cfg.num_labels)
    assert rel_err(pert.scores, ref) < TOL
    assert torch.equal(ans.cpu(), ref.argmax(-1))
