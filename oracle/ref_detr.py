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
            self.transformer = Transformer(cfg.d_model, cfg.nhead, cfg.enc_layers, cfg.dec_layers, cfg.dim_ff, 0.0)
            self.class_embed = L.Linear(cfg.d_model, cfg.classes + 1)
            self.query_embed = torch.nn.Embedding(cfg.queries, cfg.d_model)

        def forward(self, inp):
            src, pos = inp
            mask = torch.zeros(src.shape[0], src.shape[2], src.shape[3], dtype=torch.bool)
            hs, _ = self.transformer(src, mask, self.query_embed.weight, pos)
            return {"pred_logits": self.class_embed(hs)[-1]}

    m = Wrap().eval()
    missing, unexpected = m.load_state_dict(to_checkpoint_format(sd), strict=False)
    assert not missing, missing
    return m, Generator


def generate_ours(cfg, sd, src, pos, tq, **kw):
    m, Generator = build(cfg, sd)
    gen = Generator(m)
    outs = []
    with rs.cuda_is_identity():
        for b in range(src.shape[0]):
            r = gen.generate_ours((src[b:b + 1], pos[b:b + 1]), torch.tensor([int(tq[b])]), use_lrp=False, **kw)
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
