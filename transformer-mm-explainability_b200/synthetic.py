"""Synthetic CLIP workload of the benchmark: seeded random-init weights with the reference
constructor's distributions and a seeded batch of images / token rows.  There is no network for checkpoints or data
sets, so ``bench.py``, ``__graft_entry__.smoke()`` and the tests all draw from here (the CPU oracle re-exports these
two functions so that both sides of every comparison see the same tensors)."""
from __future__ import annotations

import math
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


def clip_synthetic_inputs(cfg, batch: int, seed: int = 1234):
    """Seeded synthetic batch: N(0,1) pixels; token rows ``[SOT, U{1..V-3}.., EOT, 0..]`` with
    EOT the unique row maximum so ``argmax`` finds it (CLIP/clip/model.py:360)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, cfg.image_resolution, cfg.image_resolution, generator=g)
    sot, eot = cfg.vocab_size - 2, cfg.vocab_size - 1
    tokens = torch.zeros(batch, cfg.context_length, dtype=torch.int64)
    lo, hi = 1, max(2, cfg.context_length - 2)
    for b in range(batch):
        n = int(torch.randint(lo, hi + 1, (1,), generator=g))
        tokens[b, 0] = sot
        tokens[b, 1:1 + n] = torch.randint(1, cfg.vocab_size - 2, (n,), generator=g)
        tokens[b, 1 + n] = eot
    return images, tokens
