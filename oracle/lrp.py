"""Oracle restatement of the LRP (``relprop``) sweep behind ``use_lrp=True``, DETR flavour.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); the product does not implement ``use_lrp=True`` yet (it raises) -
this file and the ``lrp`` goldens pin the target for it.  With ``use_lrp=True`` the generators feed rule 5 with the LRP
relevance of the attention probabilities (``get_attn_cam()``) instead of the probabilities themselves
(DETR/modules/ExplanationGenerator.py:112-115); that relevance comes from one relevance-propagation sweep from the
one-hot logit back through every layer.  Reference lines followed (DETR/modules/layers.py unless noted):
  * safe_divide :11-14;  RelProp / RelPropSimple (gradient x input with S = R / Z) :38-66
  * Linear.relprop :409-432 (alpha-beta rule on the positive / negative parts of x and W, bias ignored, renormalised to R.sum())
  * Add.relprop :194-221 (z-rule, then each branch rescaled to its share |sum| of R.sum())
  * Clone.relprop :252-270,  IndexSelect.relprop :230-249,  einsum (RelPropSimple) :223-228
  * LayerNorm / ReLU / Softmax / Dropout / WithPosEmbd: relevance passes through unchanged (:48-49, :107-108)
  * MultiheadAttention.relprop :770-801 (attn_cam saved after the /2, the zero-value special case :790-799)
  * TransformerEncoderLayer.forward_post_relprop DETR/models/transformer.py:256-275,  TransformerDecoderLayer ... :410-436
    (``sum`` instead of clone1.relprop :434),  TransformerDecoder.relprop :166-199 (return_intermediate=True, as
    build_transformer constructs it),  Transformer.relprop :68-79,  DETR.relprop DETR/models/detr.py:79-92
One sample per sweep, like the reference (several steps normalise by sums over the whole tensor).
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn.functional as F


def safe_divide(a, b):
    den = b.clamp(min=1e-9) + b.clamp(max=1e-9)
    den = den + den.eq(0).type(den.type()) * 1e-9
    return a / den * b.ne(0).type(b.type())


def linear_relprop(R, X, W, alpha: float = 1.0):
    beta = alpha - 1
    pw, nw = W.clamp(min=0), W.clamp(max=0)
    px, nx = X.clamp(min=0), X.clamp(max=0)

    def f(w1, w2, x1, x2):
        S = safe_divide(R, F.linear(x1, w1) + F.linear(x2, w2))
        return x1 * (S @ w1) + x2 * (S @ w2)

    out = alpha * f(pw, nw, px, nx) - beta * f(nw, pw, px, nx)
    return out * safe_divide(R.sum(), out.sum())


def add_relprop(R, x0, x1):
    S = safe_divide(R, x0 + x1)
    a, b = x0 * S, x1 * S
    a_sum, b_sum = a.sum(), b.sum()
    a_fact = safe_divide(a_sum.abs(), a_sum.abs() + b_sum.abs()) * R.sum()
    b_fact = safe_divide(b_sum.abs(), a_sum.abs() + b_sum.abs()) * R.sum()
    return a * safe_divide(a_fact, a.sum()), b * safe_divide(b_fact, b.sum())


def clone_relprop(R_list, X):
    C = sum(safe_divide(r, X) for r in R_list)
    return X * C


def index_select_relprop(R, X, dim, indices):
    S = safe_divide(R, X.index_select(dim, indices))
    return X * torch.zeros_like(X).index_add_(dim, indices, S)


def scores_relprop(R, q, k):
    """einsum 'bid,bjd->bij' as RelPropSimple: returns (R_q, R_k)."""
    S = safe_divide(R, torch.einsum('bid,bjd->bij', q, k))
    return q * torch.einsum('bij,bjd->bid', S, k), k * torch.einsum('bij,bid->bjd', S, q)


def pv_relprop(R, A, v):
    """einsum 'bij,bjd->bid' as RelPropSimple: returns (R_A, R_v)."""
    S = safe_divide(R, torch.einsum('bij,bjd->bid', A, v))
    return A * torch.einsum('bid,bjd->bij', S, v), v * torch.einsum('bij,bid->bjd', A, S)


class _MHA:
    """Forward of DETR's MultiheadAttention keeping what its relprop needs (layers.py:728-801)."""

    def __init__(self, sd, p, H):
        self.w = {n: sd[p + n + "_proj.weight"] for n in ("q", "k", "v", "out")}
        self.b = {n: sd[p + n + "_proj.bias"] for n in ("q", "k", "v", "out")}
        self.H = H

    def forward(self, query, key, value):
        T, B, D = query.shape
        self.T, self.S, self.B, self.hd = T, key.shape[0], B, D // self.H
        self.Xq, self.Xk, self.Xv = query, key, value
        q = F.linear(query, self.w["q"], self.b["q"]) * (float(self.hd) ** -0.5)
        k = F.linear(key, self.w["k"], self.b["k"])
        v = F.linear(value, self.w["v"], self.b["v"])
        self.q = q.contiguous().view(T, B * self.H, self.hd).transpose(0, 1)
        self.k = k.contiguous().view(-1, B * self.H, self.hd).transpose(0, 1)
        self.v = v.contiguous().view(-1, B * self.H, self.hd).transpose(0, 1)
        self.attn = torch.einsum('bid,bjd->bij', self.q, self.k).softmax(dim=-1)
        self.Xo = torch.einsum('bij,bjd->bid', self.attn, self.v).transpose(0, 1).contiguous().view(T, B, D)
        return F.linear(self.Xo, self.w["out"], self.b["out"])

    def relprop(self, cam):
        cam = linear_relprop(cam, self.Xo, self.w["out"])
        cam = cam.view(self.T, self.B * self.H, self.hd).transpose(0, 1)
        cam_A, cam_v = pv_relprop(cam, self.attn, self.v)
        cam_A, cam_v = cam_A / 2, cam_v / 2
        self.attn_cam = cam_A
        cam_q, cam_k = scores_relprop(cam_A, self.q, self.k)           # dropout / softmax pass relevance through
        cam_q, cam_k = cam_q / 2, cam_k / 2
        D = self.H * self.hd
        cam_v = cam_v.transpose(0, 1).reshape(self.S, self.B, D)
        cam_k = cam_k.transpose(0, 1).reshape(self.S, self.B, D)
        cam_q = cam_q.transpose(0, 1).reshape(self.T, self.B, D)
        pre_zero = bool(cam_v.min() == cam_v.max() == 0)
        cam_v = linear_relprop(cam_v, self.Xv, self.w["v"])
        cam_k = linear_relprop(cam_k, self.Xk, self.w["k"])
        cam_q = linear_relprop(cam_q, self.Xq, self.w["q"])
        if bool(cam_v.min() == cam_v.max() == 0) and not pre_zero:      # all-zero value input (first decoder layer)
            ks, qs = cam_k.sum(), cam_q.sum()
            k_fact = safe_divide(ks.abs(), ks.abs() + qs.abs()) * cam.sum()
            q_fact = safe_divide(qs.abs(), ks.abs() + qs.abs()) * cam.sum()
            cam_k = cam_k * safe_divide(k_fact, cam_k.sum())
            cam_q = cam_q * safe_divide(q_fact, cam_q.sum())
        return cam_q, cam_k, cam_v


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"])


class _EncLayer:
    def __init__(self, sd, p, H):
        self.sd, self.p = sd, p
        self.attn = _MHA(sd, p + "self_attn.", H)

    def forward(self, src, pos):
        sd, p = self.sd, self.p
        self.src = src
        self.webmd = src + pos
        self.drop = self.attn.forward(self.webmd, self.webmd, src)
        self.x1 = _ln(sd, p + "norm1", src + self.drop)
        self.h = F.relu(F.linear(self.x1, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))
        self.ff = F.linear(self.h, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        return _ln(sd, p + "norm2", self.x1 + self.ff)

    def relprop(self, cam):                                             # transformer.py:256-275
        sd, p = self.sd, self.p
        cam_2, cam2 = add_relprop(cam, self.x1, self.ff)
        cam_1 = linear_relprop(cam2, self.h, sd[p + "linear2.weight"])
        cam_1 = linear_relprop(cam_1, self.x1, sd[p + "linear1.weight"])
        cam = clone_relprop([cam_1, cam_2], self.x1)
        cam_3, cam_drop = add_relprop(cam, self.src, self.drop)
        cam_q, cam_k, cam_v = self.attn.relprop(cam_drop)
        cam_w = clone_relprop([cam_q, cam_k], self.webmd)
        return clone_relprop([cam_w, cam_v, cam_3], self.src)


class _DecLayer:
    def __init__(self, sd, p, H):
        self.sd, self.p = sd, p
        self.self_attn = _MHA(sd, p + "self_attn.", H)
        self.cross = _MHA(sd, p + "multihead_attn.", H)

    def forward(self, tgt, memory, pos, qpos):
        sd, p = self.sd, self.p
        self.tgt, self.memory = tgt, memory
        self.webmd = tgt + qpos
        self.drop1 = self.self_attn.forward(self.webmd, self.webmd, tgt)
        self.x1 = _ln(sd, p + "norm1", tgt + self.drop1)
        self.drop2 = self.cross.forward(self.x1 + qpos, memory + pos, memory)
        self.x2 = _ln(sd, p + "norm2", self.x1 + self.drop2)
        self.h = F.relu(F.linear(self.x2, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))
        self.ff = F.linear(self.h, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        return _ln(sd, p + "norm3", self.x2 + self.ff)

    def relprop(self, cam):                                             # transformer.py:410-436
        sd, p = self.sd, self.p
        cam_2, cam2 = add_relprop(cam, self.x2, self.ff)
        cam2 = linear_relprop(cam2, self.h, sd[p + "linear2.weight"])
        cam_1 = linear_relprop(cam2, self.x2, sd[p + "linear1.weight"])
        cam = clone_relprop([cam_1, cam_2], self.x2)
        cam_2, cam_drop = add_relprop(cam, self.x1, self.drop2)
        cam_q, cam_k, cam_mem_2 = self.cross.relprop(cam_drop)
        cam_mem = clone_relprop([cam_k, cam_mem_2], self.memory)
        cam = clone_relprop([cam_q, cam_2], self.x1)
        cam_3, cam_drop = add_relprop(cam, self.tgt, self.drop1)
        cam_q, cam_k, cam_v = self.self_attn.relprop(cam_drop)
        cam_w = clone_relprop([cam_q, cam_k], self.webmd)
        return cam_w + cam_v + cam_3, cam_mem                           # the reference sums here (:434)


def detr_lrp_sweep(sd: Dict[str, torch.Tensor], cfg, src, pos, target_index: int, target_class=None):
    """One sample (src, pos: [1,d,h,w]).  Runs the forward and the relprop sweep; returns
    (pred_logits [1,Q,C+1], encoder layers, decoder layers) whose ``attn`` / ``attn_cam`` ([H,T,S]) feed rule 5."""
    assert src.shape[0] == 1
    H = cfg.nhead
    x = src.flatten(2).permute(2, 0, 1)
    pe = pos.flatten(2).permute(2, 0, 1)
    qe = sd["query_embed.weight"].unsqueeze(1)
    enc = [_EncLayer(sd, f"transformer.encoder.layers.{i}.", H) for i in range(cfg.enc_layers)]
    dec = [_DecLayer(sd, f"transformer.decoder.layers.{i}.", H) for i in range(cfg.dec_layers)]
    for layer in enc:
        x = layer.forward(x, pe)
    memory = x
    t = torch.zeros_like(qe)
    outs: List[torch.Tensor] = []                                       # decoder output after each layer (clone_list inputs)
    for layer in dec:
        t = layer.forward(t, memory, pe, qe)
        outs.append(t)
    hs = torch.stack([_ln(sd, "transformer.decoder.norm", o) for o in outs]).transpose(1, 2)     # [L,1,Q,d]
    logits_all = F.linear(hs, sd["class_embed.weight"], sd["class_embed.bias"])                # [L,1,Q,C+1]
    sel = torch.tensor([cfg.dec_layers - 1])
    logits = logits_all.index_select(0, sel).squeeze(0)
    # ---- relprop (detr.py:79-92)
    cam = torch.zeros_like(logits_all.index_select(0, sel))
    if target_class is None:
        target_class = logits_all.index_select(0, sel).max(dim=-1)[1][0, 0, target_index]
    cam[0, 0, target_index, target_class] = 1
    cam = index_select_relprop(cam, logits_all, 0, sel)
    cam = linear_relprop(cam, hs, sd["class_embed.weight"])
    cam_list = cam.transpose(1, 2)                                      # [L,Q,1,d]   (transformer.py:69)
    cam_mem_list = []
    cam = None
    for j in reversed(range(cfg.dec_layers)):                          # transformer.py:179-195
        if j == cfg.dec_layers - 1:
            cam = cam_list[j]
        else:
            cam = clone_relprop([cam, cam_list[j]], outs[j])
        cam, cam_mem_j = dec[j].relprop(cam)
        cam_mem_list.append(cam_mem_j)
    cam_mem = clone_relprop(cam_mem_list, memory)
    cam = clone_relprop([torch.zeros_like(memory), cam_mem], memory)    # transformer.py:70-74 (mem_zero branch)
    for layer in reversed(enc):
        cam = layer.relprop(cam)
    return logits, enc, dec
