"""Oracle restatement of the CLIP (ViT image tower + text tower) relevancy path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Plain PyTorch + autograd on CPU.  It follows the
reference line by line but is written functionally over a ``state_dict`` (same key names as the reference
``CLIP`` module, CLIP/clip/model.py:248-303) so the same weights can be fed to the reference, to this oracle
and to the CUDA engine.

Reference lines followed:
  * model forward ............ CLIP/clip/model.py:195-198 (block), :229-246 (vision), :349-362 (text), :364-378
  * attention + hook site .... CLIP/clip/auxilary.py:72 (scaling), :153 (q*scaling), :194-198 (head split,
                               index b*H+h), :225-244 (bmm, additive mask, softmax), :247-255
  * causal mask .............. CLIP/clip/model.py:334-340
  * interpret() .............. CLIP_explainability.ipynb:151-208 (cell 6)
The only deliberate difference: the notebook calls ``torch.autograd.grad`` once per block; here all dA_l come
from ONE ``autograd.grad`` call (same values).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from .rules import avg_heads_batched


@dataclass(frozen=True)
class ClipConfig:
    """Constructor arguments of the reference ``CLIP`` (CLIP/clip/model.py:249-262), ViT image tower only."""
    embed_dim: int = 512
    image_resolution: int = 224
    vision_layers: int = 12
    vision_width: int = 768
    vision_patch_size: int = 32
    context_length: int = 77
    vocab_size: int = 49408
    transformer_width: int = 512
    transformer_heads: int = 8
    transformer_layers: int = 12

    @property
    def vision_heads(self) -> int:  # CLIP/clip/model.py:276
        return self.vision_width // 64

    @property
    def grid(self) -> int:
        return self.image_resolution // self.vision_patch_size

    @property
    def vision_tokens(self) -> int:
        return self.grid * self.grid + 1

    def ref_args(self) -> Tuple[int, ...]:
        return (self.embed_dim, self.image_resolution, self.vision_layers, self.vision_width,
                self.vision_patch_size, self.context_length, self.vocab_size, self.transformer_width,
                self.transformer_heads, self.transformer_layers)

    def to_dict(self):
        return asdict(self)


VIT_B32 = ClipConfig()
VIT_L14_336 = ClipConfig(768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12)
TINY = ClipConfig(32, 32, 2, 64, 16, 8, 64, 32, 2, 2)           # golden-fixture size
SMALL = ClipConfig(64, 64, 3, 128, 16, 16, 512, 64, 2, 3)        # 17 vision tokens (odd), 16 text tokens
EXAMPLE = ClipConfig(32, 224, 3, 64, 32, 16, 100, 64, 2, 2)      # 7x7 patches at 224 px: the grid CLIP/example.py:42 hard-codes


def init_state_dict(cfg: ClipConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Random-init weights of the synthetic workload (shared with bench.py: mmx_b200/synthetic.py)."""
    from mmx_b200.synthetic import clip_init_state_dict
    return clip_init_state_dict(cfg, seed, dtype)


def synthetic_inputs(cfg: ClipConfig, batch: int, seed: int = 1234):
    """Seeded synthetic batch of the workload (shared with bench.py: mmx_b200/synthetic.py)."""
    from mmx_b200.synthetic import clip_synthetic_inputs
    return clip_synthetic_inputs(cfg, batch, seed)


def _mha(x, sd, prefix, heads, mask, stage: List[torch.Tensor]):
    """x: [S,B,D] (LND like the reference).  CLIP/clip/auxilary.py:26-262, self-attention branch."""
    S, B, D = x.shape
    hd = D // heads
    qkv = F.linear(x, sd[prefix + "in_proj_weight"], sd[prefix + "in_proj_bias"])      # :77
    q, k, v = qkv.chunk(3, dim=-1)
    q = q * (float(hd) ** -0.5)                                                           # :72,:153
    q = q.contiguous().view(S, B * heads, hd).transpose(0, 1)                             # :194
    k = k.contiguous().view(S, B * heads, hd).transpose(0, 1)
    v = v.contiguous().view(S, B * heads, hd).transpose(0, 1)
    w = torch.bmm(q, k.transpose(1, 2))                                                   # :225
    if mask is not None:
        w = w + mask.unsqueeze(0)                                                         # :228-232
    w = F.softmax(w, dim=-1)                                                              # :243
    stage.append(w)                                                                       # hook site :247-250
    o = torch.bmm(w, v)                                                                   # :252
    o = o.transpose(0, 1).contiguous().view(S, B, D)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])     # :254-255


def _tower(x, sd, prefix, layers, heads, mask, stage):
    for i in range(layers):
        p = f"{prefix}resblocks.{i}."
        d = x.shape[-1]
        h = F.layer_norm(x, (d,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        x = x + _mha(h, sd, p + "attn.", heads, mask, stage)                              # model.py:196
        h = F.layer_norm(x, (d,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
        f = F.linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
        f = f * torch.sigmoid(1.702 * f)                                                  # QuickGELU model.py:162-164
        x = x + F.linear(f, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])      # model.py:197
    return x


def clip_forward(sd, cfg: ClipConfig, images, tokens):
    """Returns (logits_per_image [B,B], A_vision list of [B*H,S,S], A_text list of [B*H,77,77],
    image_features, text_features)."""
    W = cfg.vision_width
    x = F.conv2d(images, sd["visual.conv1.weight"], stride=cfg.vision_patch_size)        # model.py:230
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    cls = sd["visual.class_embedding"] + torch.zeros(x.shape[0], 1, W, dtype=x.dtype, device=x.device)
    x = torch.cat([cls, x], dim=1) + sd["visual.positional_embedding"]
    x = F.layer_norm(x, (W,), sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    A_v: List[torch.Tensor] = []
    x = _tower(x.permute(1, 0, 2), sd, "visual.transformer.", cfg.vision_layers, cfg.vision_heads, None, A_v)
    x = x.permute(1, 0, 2)
    x = F.layer_norm(x[:, 0, :], (W,), sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    img_f = x @ sd["visual.proj"]                                                         # model.py:243-244

    Wt = cfg.transformer_width
    mask = torch.full((cfg.context_length, cfg.context_length), float("-inf"), dtype=images.dtype, device=images.device).triu_(1)
    t = F.embedding(tokens, sd["token_embedding.weight"]) + sd["positional_embedding"]   # model.py:350-352
    A_t: List[torch.Tensor] = []
    t = _tower(t.permute(1, 0, 2), sd, "transformer.", cfg.transformer_layers, cfg.transformer_heads, mask, A_t)
    t = t.permute(1, 0, 2)
    t = F.layer_norm(t, (Wt,), sd["ln_final.weight"], sd["ln_final.bias"])
    txt_f = t[torch.arange(t.shape[0], device=t.device), tokens.argmax(dim=-1)] @ sd["text_projection"]   # model.py:360
    img_n = img_f / img_f.norm(dim=-1, keepdim=True)
    txt_n = txt_f / txt_f.norm(dim=-1, keepdim=True)
    logits = sd["logit_scale"].exp() * img_n @ txt_n.t()                                  # model.py:373-374
    return logits, A_v, A_t, img_f, txt_f


def clip_interpret(sd, cfg: ClipConfig, images, tokens, start_layer: int = -1, start_layer_text: int = -1,
                   dtype=torch.float32, return_stages: bool = False, per_layer_grad: bool = False):
    """Restatement of ``interpret()`` (CLIP_explainability.ipynb:151-208) for B distinct images (or one image
    repeated when ``images.shape[0] == 1``, :153).  Returns ``(text_relevance [B,77,77], image_relevance
    [B,S-1])`` and, with ``return_stages``, a dict with logits, A_l, dA_l, Abar_l per tower."""
    sd = {k: v.detach().to(dtype) for k, v in sd.items()}
    B = tokens.shape[0]
    images = images.to(dtype)
    if images.shape[0] == 1 and B > 1:
        images = images.repeat(B, 1, 1, 1)
    logits, A_v, A_t, _, _ = _with_grad(sd, cfg, images, tokens)
    one_hot = logits.diagonal().sum()                                                     # ipynb:156-160
    if per_layer_grad:
        # cost-faithful mode for the CPU baseline: one autograd.grad per relevant block, exactly the notebook's
        # access pattern (ipynb:175,198); same values as the single call below
        sv = cfg.vision_layers - 1 if start_layer == -1 else start_layer
        stx = cfg.transformer_layers - 1 if start_layer_text == -1 else start_layer_text
        G_v = [torch.autograd.grad(one_hot, [a], retain_graph=True)[0] if i >= sv else None for i, a in enumerate(A_v)]
        G_t = [torch.autograd.grad(one_hot, [a], retain_graph=True)[0] if i >= stx else None for i, a in enumerate(A_t)]
    else:
        grads = torch.autograd.grad(one_hot, A_v + A_t)
        G_v, G_t = grads[:len(A_v)], grads[len(A_v):]

    def rule(A, G, start):
        L = len(A)
        if start == -1:
            start = L - 1                                                                 # ipynb:165-167
        S = A[0].shape[-1]
        R = torch.eye(S, dtype=dtype, device=A[0].device).unsqueeze(0).expand(B, S, S)
        bars = {}
        for i in range(L):
            if i < start:
                continue
            cam = avg_heads_batched(A[i].detach(), G[i], B)                               # ipynb:176-181
            bars[i] = cam
            R = R + torch.bmm(cam, R)                                                     # ipynb:182
        return R, bars

    R_img, bars_v = rule(A_v, G_v, start_layer)
    R_txt, bars_t = rule(A_t, G_t, start_layer_text)
    out = (R_txt, R_img[:, 0, 1:])                                                        # ipynb:183,206-208
    if return_stages:
        stages = dict(logits=logits.detach(), A_v=[a.detach() for a in A_v], A_t=[a.detach() for a in A_t],
                      G_v=list(G_v), G_t=list(G_t), bar_v=bars_v, bar_t=bars_t, R_img=R_img)
        return out + (stages,)
    return out


def _with_grad(sd, cfg, images, tokens):
    with torch.enable_grad():
        # make the A_l differentiable leaves-of-interest: the graph needs at least one grad-requiring input
        images = images.detach().requires_grad_(True)
        sd = dict(sd)
        sd["token_embedding.weight"] = sd["token_embedding.weight"].detach().requires_grad_(True)
        return clip_forward(sd, cfg, images, tokens)
