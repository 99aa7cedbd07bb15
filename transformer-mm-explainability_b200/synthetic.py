"""Synthetic workloads of the benchmark: seeded random-init weights with the reference constructors'
distributions and seeded input batches for CLIP (configs 2, 5), DETR (config 3) and LXMERT (config 4).  There is no
network for checkpoints or data sets, so ``bench.py``, ``__graft_entry__.smoke()`` and the tests all draw from here (the
CPU oracles re-export these functions so that both sides of every comparison see the same tensors)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict

import torch

def clip_init_state_dict(cfg, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init weights with the distributions of the reference constructor + ``initialize_parameters``
    (CLIP/clip/model.py:211-227 vision scale init, :305-332 text init; vision blocks keep
    nn.MultiheadAttention / nn.Linear defaults).  Not bit-identical to the reference RNG order - weights are
    always passed explicitly, so that does not matter."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def normal(*shape, std=1.0):
        return torch.randn(*shape, generator=g, dtype=torch.float32) * std

    def uniform(*shape, bound):
        return (torch.rand(*shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    W, Wt = cfg.vision_width, cfg.transformer_width
    sc = W ** -0.5
    sd["visual.conv1.weight"] = uniform(W, 3, cfg.vision_patch_size, cfg.vision_patch_size,
                                        bound=1.0 / math.sqrt(3 * cfg.vision_patch_size ** 2))
    sd["visual.class_embedding"] = normal(W, std=sc)
    sd["visual.positional_embedding"] = normal(cfg.vision_tokens, W, std=sc)
    sd["visual.proj"] = normal(W, cfg.embed_dim, std=sc)
    for nm in ("visual.ln_pre", "visual.ln_post", "ln_final"):
        d = Wt if nm == "ln_final" else W
        # non-trivial affine so the LayerNorm gamma/beta paths are exercised by parity tests
        sd[nm + ".weight"] = 1.0 + 0.1 * normal(d)
        sd[nm + ".bias"] = 0.05 * normal(d)

    def block(prefix, d, attn_std=None, proj_std=None, fc_std=None):
        if attn_std is None:      # vision tower: torch defaults (xavier_uniform in_proj, kaiming-uniform Linear)
            sd[prefix + "attn.in_proj_weight"] = uniform(3 * d, d, bound=math.sqrt(6.0 / (3 * d + d)))
            sd[prefix + "attn.out_proj.weight"] = uniform(d, d, bound=1 / math.sqrt(d))
            sd[prefix + "mlp.c_fc.weight"] = uniform(4 * d, d, bound=1 / math.sqrt(d))
            sd[prefix + "mlp.c_proj.weight"] = uniform(d, 4 * d, bound=1 / math.sqrt(4 * d))
        else:
            sd[prefix + "attn.in_proj_weight"] = normal(3 * d, d, std=attn_std)
            sd[prefix + "attn.out_proj.weight"] = normal(d, d, std=proj_std)
            sd[prefix + "mlp.c_fc.weight"] = normal(4 * d, d, std=fc_std)
            sd[prefix + "mlp.c_proj.weight"] = normal(d, 4 * d, std=proj_std)
        sd[prefix + "attn.in_proj_bias"] = 0.02 * normal(3 * d)
        sd[prefix + "attn.out_proj.bias"] = 0.02 * normal(d)
        sd[prefix + "mlp.c_fc.bias"] = uniform(4 * d, bound=1 / math.sqrt(d))
        sd[prefix + "mlp.c_proj.bias"] = uniform(d, bound=1 / math.sqrt(4 * d))
        for ln in ("ln_1", "ln_2"):
            sd[prefix + ln + ".weight"] = 1.0 + 0.1 * normal(d)
            sd[prefix + ln + ".bias"] = 0.05 * normal(d)

    for i in range(cfg.vision_layers):
        block(f"visual.transformer.resblocks.{i}.", W)
    proj_std = (Wt ** -0.5) * ((2 * cfg.transformer_layers) ** -0.5)
    for i in range(cfg.transformer_layers):
        block(f"transformer.resblocks.{i}.", Wt, attn_std=Wt ** -0.5, proj_std=proj_std,
              fc_std=(2 * Wt) ** -0.5)
    sd["token_embedding.weight"] = normal(cfg.vocab_size, Wt, std=0.02)
    sd["positional_embedding"] = normal(cfg.context_length, Wt, std=0.01)
    sd["text_projection"] = normal(Wt, cfg.embed_dim, std=Wt ** -0.5)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07), dtype=torch.float32)
    return {k: v.to(dtype) for k, v in sd.items()}


def clip_synthetic_inputs(cfg, batch: int, seed: int = 1234, length_seed=None):
    """Seeded synthetic batch: N(0,1) pixels; token rows ``[SOT, U{1..V-3}.., EOT, 0..]`` with
    EOT the unique row maximum so ``argmax`` finds it (CLIP/clip/model.py:360).  ``length_seed``: draw the prompt
    lengths from their own generator, so that ranks with different ``seed`` (different pixels / ids) carry the same length
    multiset - equal ragged text work on every rank of a weak-scaling run."""
    g = torch.Generator().manual_seed(seed)
    gl = g if length_seed is None else torch.Generator().manual_seed(length_seed)
    images = torch.randn(batch, 3, cfg.image_resolution, cfg.image_resolution, generator=g)
    sot, eot = cfg.vocab_size - 2, cfg.vocab_size - 1
    tokens = torch.zeros(batch, cfg.context_length, dtype=torch.int64)
    lo, hi = 1, max(2, cfg.context_length - 2)
    for b in range(batch):
        n = int(torch.randint(lo, hi + 1, (1,), generator=gl))
        tokens[b, 0] = sot
        tokens[b, 1:1 + n] = torch.randint(1, cfg.vocab_size - 2, (n,), generator=g)
        tokens[b, 1 + n] = eot
    return images, tokens


# ======================================================================================================================
# DETR (BASELINE config 3): transformer + class head on backbone features; xavier-uniform like DETR/models/transformer.py:46-49
# ======================================================================================================================
@dataclass(frozen=True)
class DetrConfig:
    d_model: int = 256
    nhead: int = 8
    enc_layers: int = 6
    dec_layers: int = 6
    dim_ff: int = 2048
    queries: int = 100
    classes: int = 91          # logits have classes + 1 entries (last = no-object)


DETR_R50 = DetrConfig()
DETR_TINY = DetrConfig(64, 2, 2, 2, 96, 7, 5)


def detr_init_state_dict(cfg: DetrConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def xavier(o, i):
        b = (6.0 / (o + i)) ** 0.5
        return (torch.rand(o, i, generator=g) * 2 - 1) * b

    def vec(n, s=0.05):
        return torch.randn(n, generator=g) * s

    def mha(p):
        for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
            sd[p + nm + ".weight"] = xavier(cfg.d_model, cfg.d_model)
            sd[p + nm + ".bias"] = vec(cfg.d_model)

    def ffn_norms(p, norms):
        sd[p + "linear1.weight"] = xavier(cfg.dim_ff, cfg.d_model); sd[p + "linear1.bias"] = vec(cfg.dim_ff)
        sd[p + "linear2.weight"] = xavier(cfg.d_model, cfg.dim_ff); sd[p + "linear2.bias"] = vec(cfg.d_model)
        for n in norms:
            sd[p + n + ".weight"] = 1 + vec(cfg.d_model, 0.1); sd[p + n + ".bias"] = vec(cfg.d_model)

    for i in range(cfg.enc_layers):
        p = f"transformer.encoder.layers.{i}."
        mha(p + "self_attn."); ffn_norms(p, ("norm1", "norm2"))
    for i in range(cfg.dec_layers):
        p = f"transformer.decoder.layers.{i}."
        mha(p + "self_attn."); mha(p + "multihead_attn."); ffn_norms(p, ("norm1", "norm2", "norm3"))
    sd["transformer.decoder.norm.weight"] = 1 + vec(cfg.d_model, 0.1)
    sd["transformer.decoder.norm.bias"] = vec(cfg.d_model)
    sd["query_embed.weight"] = torch.randn(cfg.queries, cfg.d_model, generator=g)
    sd["class_embed.weight"] = xavier(cfg.classes + 1, cfg.d_model)
    sd["class_embed.bias"] = vec(cfg.classes + 1)
    return sd


def detr_sine_position_embedding(B: int, h: int, w: int, d_model: int, temperature: float = 10000.0) -> torch.Tensor:
    """PositionEmbeddingSine(normalize=True) for an all-valid mask (DETR/models/position_encoding.py:28-50)."""
    npf = d_model // 2
    ones = torch.ones(B, h, w)
    y_embed, x_embed = ones.cumsum(1), ones.cumsum(2)
    eps, scale = 1e-6, 2 * torch.pi
    y_embed = y_embed / (y_embed[:, -1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, :, -1:] + eps) * scale
    dim_t = torch.arange(npf, dtype=torch.float32)
    dim_t = temperature ** (2 * (dim_t // 2) / npf)
    pos_x, pos_y = x_embed[:, :, :, None] / dim_t, y_embed[:, :, :, None] / dim_t
    pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
    pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
    return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


def detr_synthetic_inputs(cfg: DetrConfig, B: int, h: int, w: int, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(B, cfg.d_model, h, w, generator=g)
    pos = detr_sine_position_embedding(B, h, w, cfg.d_model)
    tq = torch.randint(0, cfg.queries, (B,), generator=g)
    return src, pos, tq


# ======================================================================================================================
# LXMERT (BASELINE config 4): LxmertForQuestionAnswering key names (lxmert/lxmert/src/lxmert_lrp.py:1532)
# ======================================================================================================================
@dataclass(frozen=True)
class LxmertConfig:
    hidden: int = 768
    heads: int = 12
    intermediate: int = 3072
    l_layers: int = 9
    x_layers: int = 5
    r_layers: int = 5
    vocab: int = 30522
    max_pos: int = 512
    feat_dim: int = 2048
    pos_dim: int = 4
    num_labels: int = 3129


LXMERT_BASE = LxmertConfig()
LXMERT_TINY = LxmertConfig(hidden=64, heads=2, intermediate=96, l_layers=2, x_layers=2, r_layers=2, vocab=50, max_pos=16,
                           feat_dim=24, pos_dim=4, num_labels=11)


def lxmert_init_state_dict(cfg: LxmertConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    Hd = cfg.hidden

    def lin(p, o, i, std=None):
        sd[p + ".weight"] = torch.randn(o, i, generator=g) * (std if std else i ** -0.5)
        sd[p + ".bias"] = torch.randn(o, generator=g) * 0.02

    def ln(p, d):
        sd[p + ".weight"] = 1 + 0.1 * torch.randn(d, generator=g)
        sd[p + ".bias"] = 0.05 * torch.randn(d, generator=g)

    def att(p):
        for n in ("query", "key", "value"):
            lin(p + n, Hd, Hd)

    def att_out(p):
        lin(p + "dense", Hd, Hd); ln(p + "LayerNorm", Hd)

    def ffn(pi, po):
        lin(pi + "dense", cfg.intermediate, Hd); lin(po + "dense", Hd, cfg.intermediate); ln(po + "LayerNorm", Hd)

    e = "lxmert.embeddings."
    sd[e + "word_embeddings.weight"] = torch.randn(cfg.vocab, Hd, generator=g) * 0.5
    sd[e + "position_embeddings.weight"] = torch.randn(cfg.max_pos, Hd, generator=g) * 0.5
    sd[e + "token_type_embeddings.weight"] = torch.randn(2, Hd, generator=g) * 0.5
    ln(e + "LayerNorm", Hd)
    v = "lxmert.encoder.visn_fc."
    lin(v + "visn_fc", Hd, cfg.feat_dim); ln(v + "visn_layer_norm", Hd); lin(v + "box_fc", Hd, cfg.pos_dim); ln(v + "box_layer_norm", Hd)
    for name, n in (("layer", cfg.l_layers), ("r_layers", cfg.r_layers)):
        for i in range(n):
            p = f"lxmert.encoder.{name}.{i}."
            att(p + "attention.self."); att_out(p + "attention.output."); ffn(p + "intermediate.", p + "output.")
    for i in range(cfg.x_layers):
        p = f"lxmert.encoder.x_layers.{i}."
        att(p + "visual_attention.att."); att_out(p + "visual_attention.output.")
        for s in ("lang_self_att", "visn_self_att"):
            att(p + s + ".self."); att_out(p + s + ".output.")
        ffn(p + "lang_inter.", p + "lang_output."); ffn(p + "visn_inter.", p + "visn_output.")
    lin("lxmert.pooler.dense", Hd, Hd)
    lin("answer_head.logit_fc.0", 2 * Hd, Hd); ln("answer_head.logit_fc.2", 2 * Hd); lin("answer_head.logit_fc.3", cfg.num_labels, 2 * Hd)
    return sd


def lxmert_synthetic_inputs(cfg: LxmertConfig, B: int, T: int, I: int, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, cfg.vocab, (B, T), generator=g)
    ids[:, 0] = 1
    ids[:, -1] = 2
    feats = torch.randn(B, I, cfg.feat_dim, generator=g)
    boxes = torch.rand(B, I, cfg.pos_dim, generator=g)
    return ids, feats, boxes
