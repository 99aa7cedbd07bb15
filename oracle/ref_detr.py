"""Runs the UNMODIFIED reference DETR transformer + Generator on CPU (build container only; TEST INFRASTRUCTURE).
Shims: torchvision version guard at DETR/util/misc.py:21, ``.cuda()`` identity; the ResNet backbone is
replaced by given features (it sits below every attention layer)."""
from __future__ import annotations

import sys

import torch

from . import ref_shims as rs


def build(cfg, sd):
    rs._ensure_path()
    import torchvision
    tv = torchvision.__version__
    torchvision.__version__ = "0.9.9"
    try:
        from DETR.models.transformer import Transformer
        from DETR.modules import layers as L
        from DETR.modules.ExplanationGenerator import Generator
    finally:
        torchvision.__version__ = tv
    from .detr_oracle import to_checkpoint_format

    class Wrap(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.transformer = Transformer(cfg.d_model, cfg.nhead, cfg.enc_layers, cfg.dec_layers, cfg.dim_ff, 0.0,
                                           return_intermediate_dec=True)       # as build() does: DETR/models/transformer.py:466-476
            self.class_embed = L.Linear(cfg.d_model, cfg.classes + 1)
            self.query_embed = torch.nn.Embedding(cfg.queries, cfg.d_model)
            self.index_select = L.IndexSelect()

        def forward(self, inp):
            src, pos = inp
            mask = torch.zeros(src.shape[0], src.shape[2], src.shape[3], dtype=torch.bool)
            hs, memory = self.transformer(src, mask, self.query_embed.weight, pos)
            self.memory_shape = memory.shape
            # DETR.forward (DETR/models/detr.py:71-72): the last decoder layer's logits through IndexSelect (the
            # reference hard-codes index 5 = the last of its 6 intermediate outputs)
            a = self.index_select(self.class_embed(hs), 0, torch.tensor([hs.shape[0] - 1])).squeeze(0)
            return {"pred_logits": a}

        def relprop(self, cam=None, alpha=1, **kwargs):
            # DETR.relprop, DETR/models/detr.py:79-92, verbatim in structure (the full DETR class needs the backbone)
            model_predictions = self.index_select.Y
            cam = torch.zeros_like(model_predictions)
            if kwargs["target_class"] is None:
                probs = model_predictions.max(dim=-1)[1]
                kwargs["target_class"] = probs[0, kwargs["target_index"]]
            cam[0, 0, kwargs["target_index"], kwargs["target_class"]] = 1
            cam = self.index_select.relprop(cam, alpha)
            cam = self.class_embed.relprop(cam, alpha)
            mem_zero = torch.zeros(self.memory_shape).to(cam.device)
            return self.transformer.relprop([cam, mem_zero], alpha)

    m = Wrap().eval()
    missing, unexpected = m.load_state_dict(to_checkpoint_format(sd), strict=False)
    assert not missing, missing
    return m, Generator


def generate_ours(cfg, sd, src, pos, tq, use_lrp=False, **kw):
    m, Generator = build(cfg, sd)
    gen = Generator(m)
    outs = []
    with rs.cuda_is_identity():
        for b in range(src.shape[0]):
            r = gen.generate_ours((src[b:b + 1], pos[b:b + 1]), torch.tensor([int(tq[b])]), use_lrp=use_lrp, **kw)
            outs.append(r.reshape(-1).detach())
    return torch.stack(outs)


def generate_baseline(cfg, sd, src, pos, tq, method):
    """method in {"raw_attn", "rollout", "attn_gradcam"}: the reference Generator's baseline of that name, per sample."""
    m, Generator = build(cfg, sd)
    gen = Generator(m)
    outs = []
    with rs.cuda_is_identity():
        for b in range(src.shape[0]):
            fn = getattr(gen, "generate_" + method)
            r = fn((src[b:b + 1], pos[b:b + 1]), torch.tensor([int(tq[b])]))
            outs.append(r.reshape(-1).detach())
    return torch.stack(outs)


def generate_ours_abl(cfg, sd, src, pos, tq, **kw):
    """The reference's GeneratorAlbationNoAgg.generate_ours_abl (DETR/modules/ExplanationGenerator.py:306-403), per sample."""
    m, _ = build(cfg, sd)
    from DETR.modules.ExplanationGenerator import GeneratorAlbationNoAgg
    gen = GeneratorAlbationNoAgg(m)
    outs = []
    with rs.cuda_is_identity():
        for b in range(src.shape[0]):
            r = gen.generate_ours_abl((src[b:b + 1], pos[b:b + 1]), torch.tensor([int(tq[b])]), use_lrp=False, **kw)
            outs.append(r.reshape(-1).detach())
    return torch.stack(outs)


def lrp_attn_cams(cfg, sd, src, pos, tq, double=False):
    """The ``get_attn_cam()`` tensors the reference holds after ``generate_ours(use_lrp=True)`` on one sample: dict
    enc{i} / dec{i}.self / dec{i}.cross -> [H,T,S].  ``double=True`` runs the reference model in float64 (the relprop
    sweep divides by many small numbers, so fp32 results carry ~1e-4 of conditioning noise); its fp32 ``R = eye`` then
    stops the rule stage with a dtype error AFTER the sweep, which is all that is needed here."""
    m, Generator = build(cfg, sd)
    if double:
        m, src, pos = m.double(), src.double(), pos.double()
    gen = Generator(m)
    with rs.cuda_is_identity():
        try:
            gen.generate_ours((src, pos), torch.tensor([tq]), use_lrp=True, normalize_self_attention=False)
        except RuntimeError:
            if not double:
                raise
    out = {}
    for i, blk in enumerate(m.transformer.encoder.layers):
        out[f"enc{i}"] = blk.self_attn.get_attn_cam().detach().clone()
    for i, blk in enumerate(m.transformer.decoder.layers):
        out[f"dec{i}.self"] = blk.self_attn.get_attn_cam().detach().clone()
        out[f"dec{i}.cross"] = blk.multihead_attn.get_attn_cam().detach().clone()
    return out
