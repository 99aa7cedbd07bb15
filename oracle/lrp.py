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


# ======================================================================================================================
# LXMERT flavour (lxmert/lxmert/src/layers.py + lxmert_lrp.py).  Differences from DETR's layer library: Linear.relprop
# does NOT renormalise to R.sum() (layers.py:219-242), the attention products are RelPropSimple MatMuls on [B,H,T,S]
# tensors (layers.py:89-91), and the scores are divided by sqrt(d) after the product (a plain tensor op: relevance
# passes through).  Reference relprop lines: LxmertAttention :422-461, LxmertAttentionOutput :479-484,
# LxmertCrossAttentionLayer :505-509, LxmertSelfAttentionLayer :535-539, LxmertIntermediate :554-557, LxmertOutput
# :575-580, LxmertLayer :601-606, LxmertXLayer :657-733, LxmertEncoder :855-863, LxmertPooler :886-892,
# LxmertVisualAnswerHead :955-958, LxmertModel :1253-1257, LxmertForQuestionAnswering :1689-1693.
# ======================================================================================================================
def _lx_linear_relprop(R, X, W):
    pw, nw = W.clamp(min=0), W.clamp(max=0)
    px, nx = X.clamp(min=0), X.clamp(max=0)
    S = safe_divide(R, F.linear(px, pw) + F.linear(nx, nw))             # alpha = 1, beta = 0: activator term only
    S2 = safe_divide(R, F.linear(px, nw) + F.linear(nx, pw))
    inhib = px * (S2 @ nw) + nx * (S2 @ pw)
    return 1.0 * (px * (S @ pw) + nx * (S @ nw)) - 0.0 * inhib


def _matmul_relprop(R, a, b):
    """RelPropSimple on torch.matmul(a, b): returns (R_a, R_b)."""
    S = safe_divide(R, a @ b)
    return a * (S @ b.transpose(-1, -2)), b * (a.transpose(-1, -2) @ S)


class _LxAttention:
    def __init__(self, sd, p, H):
        self.sd, self.p, self.H = sd, p, H

    def _heads(self, x):
        B, T, D = x.shape
        return x.view(B, T, self.H, D // self.H).permute(0, 2, 1, 3)

    def forward(self, hidden, context):
        sd, p = self.sd, self.p
        self.hidden, self.context = hidden, context
        lin = lambda n, x: F.linear(x, sd[p + n + ".weight"], sd[p + n + ".bias"])
        self.q, self.k, self.v = self._heads(lin("query", hidden)), self._heads(lin("key", context)), self._heads(lin("value", context))
        hd = self.q.shape[-1]
        self.probs = (self.q @ self.k.transpose(-1, -2) / (hd ** 0.5)).softmax(dim=-1)
        ctx = (self.probs @ self.v).permute(0, 2, 1, 3).contiguous()
        return ctx.view(ctx.shape[0], ctx.shape[1], -1)

    def relprop(self, cam):
        sd, p = self.sd, self.p
        merge = lambda x: x.permute(0, 2, 1, 3).flatten(2)
        cam = self._heads(cam)
        cam1, cam2 = _matmul_relprop(cam, self.probs, self.v)
        cam1, cam2 = cam1 / 2, cam2 / 2
        self.attn_cam = cam1
        kT = self.k.transpose(-1, -2)
        cam_q, cam_kT = _matmul_relprop(cam1, self.q, kT)
        cam_q, cam_kT = cam_q / 2, cam_kT / 2
        cam_q = _lx_linear_relprop(merge(cam_q), self.hidden, sd[p + "query.weight"])
        cam_k = _lx_linear_relprop(merge(cam_kT.transpose(-1, -2)), self.context, sd[p + "key.weight"])
        cam_v = _lx_linear_relprop(merge(cam2), self.context, sd[p + "value.weight"])
        return cam_q, clone_relprop((cam_k, cam_v), self.context)


class _LxAttBlock:
    """LxmertCrossAttentionLayer / LxmertSelfAttentionLayer: attention + (dense, +input, LayerNorm)."""

    def __init__(self, sd, p_att, p_out, H):
        self.sd, self.p_out = sd, p_out
        self.att = _LxAttention(sd, p_att, H)

    def forward(self, x, ctx):
        sd, p = self.sd, self.p_out
        self.x = x
        self.ctx_out = self.att.forward(x, ctx)
        self.dense = F.linear(self.ctx_out, sd[p + "dense.weight"], sd[p + "dense.bias"])
        return F.layer_norm(self.dense + x, (x.shape[-1],), sd[p + "LayerNorm.weight"], sd[p + "LayerNorm.bias"], 1e-12)

    def relprop(self, cam, self_attention: bool):
        cam_dense, cam_res = add_relprop(cam, self.dense, self.x)
        cam_out = _lx_linear_relprop(cam_dense, self.ctx_out, self.sd[self.p_out + "dense.weight"])
        cam_hidden, cam_ctx = self.att.relprop(cam_out)
        if self_attention:                                               # clone(input, 3): query, context, residual
            return clone_relprop((cam_hidden, cam_ctx, cam_res), self.x)
        return clone_relprop((cam_hidden, cam_res), self.x), cam_ctx       # clone(input, 2) + the context relevance


class _LxFfn:
    """LxmertIntermediate + LxmertOutput around a Clone (LxmertLayer :592-606, LxmertXLayer.output_fc :680-697)."""

    def __init__(self, sd, pi, po):
        self.sd, self.pi, self.po = sd, pi, po

    def forward(self, x):
        sd = self.sd
        self.x = x
        self.inter = F.gelu(F.linear(x, sd[self.pi + "dense.weight"], sd[self.pi + "dense.bias"]))
        self.dense = F.linear(self.inter, sd[self.po + "dense.weight"], sd[self.po + "dense.bias"])
        return F.layer_norm(self.dense + x, (x.shape[-1],), sd[self.po + "LayerNorm.weight"], sd[self.po + "LayerNorm.bias"], 1e-12)

    def relprop(self, cam):
        cam1, cam2 = add_relprop(cam, self.dense, self.x)
        cam1 = _lx_linear_relprop(cam1, self.inter, self.sd[self.po + "dense.weight"])
        cam1 = _lx_linear_relprop(cam1, self.x, self.sd[self.pi + "dense.weight"])
        return clone_relprop((cam1, cam2), self.x)


class _LxLayer:
    def __init__(self, sd, p, H):
        self.att = _LxAttBlock(sd, p + "attention.self.", p + "attention.output.", H)
        self.ffn = _LxFfn(sd, p + "intermediate.", p + "output.")

    def forward(self, x):
        return self.ffn.forward(self.att.forward(x, x))

    def relprop(self, cam):
        return self.att.relprop(self.ffn.relprop(cam), True)


class _LxXLayer:
    def __init__(self, sd, p, H):
        mk = lambda: _LxAttBlock(sd, p + "visual_attention.att.", p + "visual_attention.output.", H)
        self.cross, self.cross_copy = mk(), mk()                       # the deepcopy: same weights, own activations
        self.lang_self = _LxAttBlock(sd, p + "lang_self_att.self.", p + "lang_self_att.output.", H)
        self.visn_self = _LxAttBlock(sd, p + "visn_self_att.self.", p + "visn_self_att.output.", H)
        self.lang_ffn = _LxFfn(sd, p + "lang_inter.", p + "lang_output.")
        self.visn_ffn = _LxFfn(sd, p + "visn_inter.", p + "visn_output.")

    def forward(self, lang, vis):
        self.lang_in, self.vis_in = lang, vis
        l2, v2 = self.cross.forward(lang, vis), self.cross_copy.forward(vis, lang)
        l3, v3 = self.lang_self.forward(l2, l2), self.visn_self.forward(v2, v2)
        return self.lang_ffn.forward(l3), self.visn_ffn.forward(v3)

    def relprop(self, cam_lang, cam_vis):
        cam_vis, cam_lang = self.visn_ffn.relprop(cam_vis), self.lang_ffn.relprop(cam_lang)            # relprop_output
        cam_vis, cam_lang = self.visn_self.relprop(cam_vis, True), self.lang_self.relprop(cam_lang, True)  # relprop_self
        cam_vis2, cam_lang2 = self.cross_copy.relprop(cam_vis, False)                                   # relprop_cross
        cam_lang1, cam_vis1 = self.cross.relprop(cam_lang, False)
        return clone_relprop((cam_lang1, cam_lang2), self.lang_in), clone_relprop((cam_vis1, cam_vis2), self.vis_in)


def lxmert_lrp_sweep(sd: Dict[str, torch.Tensor], cfg, ids, feats, boxes, index=None):
    """One sample (ids [1,T], feats [1,I,F], boxes [1,I,4]).  Forward + relprop; returns (logits, dict of layer lists
    lang / vis / x) whose attention objects carry ``probs`` and ``attn_cam`` ([1,H,T,S])."""
    assert ids.shape[0] == 1
    H = cfg.heads
    e = "lxmert.embeddings."
    T = ids.shape[1]
    ln = lambda p, x: F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-12)
    lin = lambda p, x: F.linear(x, sd[p + ".weight"], sd[p + ".bias"])
    lang = ln(e + "LayerNorm", sd[e + "token_type_embeddings.weight"][torch.zeros_like(ids)]
              + sd[e + "position_embeddings.weight"][torch.arange(T)] + sd[e + "word_embeddings.weight"][ids])
    v = "lxmert.encoder.visn_fc."
    vis = (ln(v + "visn_layer_norm", lin(v + "visn_fc", feats)) + ln(v + "box_layer_norm", lin(v + "box_fc", boxes))) / 2
    L = [_LxLayer(sd, f"lxmert.encoder.layer.{i}.", H) for i in range(cfg.l_layers)]
    R = [_LxLayer(sd, f"lxmert.encoder.r_layers.{i}.", H) for i in range(cfg.r_layers)]
    X = [_LxXLayer(sd, f"lxmert.encoder.x_layers.{i}.", H) for i in range(cfg.x_layers)]
    for layer in L:
        lang = layer.forward(lang)
    for layer in R:
        vis = layer.forward(vis)
    for layer in X:
        lang, vis = layer.forward(lang, vis)
    first = lang[:, 0]
    pooled = torch.tanh(lin("lxmert.pooler.dense", first))
    h_pre = lin("answer_head.logit_fc.0", pooled)
    h = F.layer_norm(F.gelu(h_pre), (h_pre.shape[-1],), sd["answer_head.logit_fc.2.weight"], sd["answer_head.logit_fc.2.bias"], 1e-12)
    logits = lin("answer_head.logit_fc.3", h)
    # ---- relprop from the one-hot answer (ExplanationGenerator.py:152-165)
    idx = int(logits.argmax(-1)) if index is None else int(index)
    cam = torch.zeros_like(logits)
    cam[0, idx] = 1
    cam = _lx_linear_relprop(cam, h, sd["answer_head.logit_fc.3.weight"])
    cam = _lx_linear_relprop(cam, pooled, sd["answer_head.logit_fc.0.weight"])      # LayerNorm / GELU pass through
    cam = _lx_linear_relprop(cam, first, sd["lxmert.pooler.dense.weight"])          # Tanh passes through
    cam_lang = index_select_relprop(cam.unsqueeze(1), lang, 1, torch.tensor([0]))
    cam_vis = torch.zeros_like(vis)
    for layer in reversed(X):
        cam_lang, cam_vis = layer.relprop(cam_lang, cam_vis)
    for layer in reversed(R):
        cam_vis = layer.relprop(cam_vis)
    for layer in reversed(L):
        cam_lang = layer.relprop(cam_lang)
    return logits, {"lang": L, "vis": R, "x": X}


# ======================================================================================================================
# VisualBERT flavour (VisualBERT/mmf/models/transformers/backends/BERT_ours.py + layers_ours.py: the same primitive
# library as LXMERT).  BertSelfAttention.relprop :352-395 (clone of the hidden state into q / k / v, the additive mask goes
# through Add.relprop because ``self.attention_mask`` is set), BertAttention :227-232, BertSelfOutput :412-419,
# BertIntermediate :436-441, BertOutput :459-472, BertLayer :506-514, BertEncoder :152-156, BertPredictionHeadTransform
# :533-537; VisualBERTForClassification.relprop VisualBERT/mmf/models/visual_bert.py:398-403.
# ======================================================================================================================
class _BertLayer:
    def __init__(self, sd, p, H):
        self.sd, self.p, self.H = sd, p, H
        self.ffn = _LxFfn(sd, p + "intermediate.", p + "output.")

    def _heads(self, x):
        B, T, D = x.shape
        return x.view(B, T, self.H, D // self.H).permute(0, 2, 1, 3)

    def forward(self, x, mask):
        sd, p = self.sd, self.p
        self.x, self.mask = x, mask
        lin = lambda n, t: F.linear(t, sd[p + n + ".weight"], sd[p + n + ".bias"])
        self.q, self.k, self.v = (self._heads(lin("attention.self." + n, x)) for n in ("query", "key", "value"))
        hd = self.q.shape[-1]
        self.scores = self.q @ self.k.transpose(-1, -2) / (hd ** 0.5)
        self.probs = (self.scores + mask).softmax(dim=-1) if mask is not None else self.scores.softmax(dim=-1)
        ctx = (self.probs @ self.v).permute(0, 2, 1, 3).contiguous()
        self.ctx = ctx.view(ctx.shape[0], ctx.shape[1], -1)
        self.dense = lin("attention.output.dense", self.ctx)
        self.att_out = F.layer_norm(self.dense + x, (x.shape[-1],), sd[p + "attention.output.LayerNorm.weight"],
                                    sd[p + "attention.output.LayerNorm.bias"], 1e-12)
        return self.ffn.forward(self.att_out)

    def relprop(self, cam):
        sd, p = self.sd, self.p
        merge = lambda t: t.permute(0, 2, 1, 3).flatten(2)
        cam = self.ffn.relprop(cam)                                       # BertLayer: output, intermediate, clone
        cam_dense, cam_res = add_relprop(cam, self.dense, self.x)         # BertSelfOutput
        cam_ctx = _lx_linear_relprop(cam_dense, self.ctx, sd[p + "attention.output.dense.weight"])
        cam1, cam2 = _matmul_relprop(self._heads(cam_ctx), self.probs, self.v)
        cam1, cam2 = cam1 / 2, cam2 / 2
        self.attn_cam = cam1
        if self.mask is not None:
            cam1, _ = add_relprop(cam1, self.scores, self.mask.expand_as(self.scores))
        cam_q, cam_kT = _matmul_relprop(cam1, self.q, self.k.transpose(-1, -2))
        cam_q, cam_kT = cam_q / 2, cam_kT / 2
        cam_q = _lx_linear_relprop(merge(cam_q), self.x, sd[p + "attention.self.query.weight"])
        cam_k = _lx_linear_relprop(merge(cam_kT.transpose(-1, -2)), self.x, sd[p + "attention.self.key.weight"])
        cam_v = _lx_linear_relprop(merge(cam2), self.x, sd[p + "attention.self.value.weight"])
        cam_h1 = clone_relprop((cam_q, cam_k, cam_v), self.x)            # BertSelfAttention's clone(hidden, 3)
        return clone_relprop((cam_h1, cam_res), self.x)                   # BertAttention's clone(hidden, 2)


def visualbert_lrp_sweep(sd: Dict[str, torch.Tensor], cfg, inp, index=None):
    """One sample (the dict of oracle/visualbert_oracle.py).  Forward + relprop; returns (scores, layers) whose
    ``probs`` / ``attn_cam`` are [1,H,S,S]."""
    ids, vis = inp["input_ids"], inp["visual_embeddings"]
    assert ids.shape[0] == 1
    T, V, H = ids.shape[1], vis.shape[1], cfg.heads
    e = "bert.embeddings."
    ln = lambda p, x: F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-12)
    lin = lambda p, x: F.linear(x, sd[p + ".weight"], sd[p + ".bias"])
    tt = inp.get("token_type_ids")
    tt = torch.zeros_like(ids) if tt is None else tt
    vt = inp.get("visual_embeddings_type")
    vt = torch.zeros(1, V, dtype=torch.long) if vt is None else vt
    text = sd[e + "word_embeddings.weight"][ids] + sd[e + "position_embeddings.weight"][torch.arange(T)] \
        + sd[e + "token_type_embeddings.weight"][tt]
    v_emb = lin(e + "projection", vis.to(text.dtype)) + sd[e + "position_embeddings_visual.weight"][torch.zeros(1, V, dtype=torch.long)] \
        + sd[e + "token_type_embeddings_visual.weight"][vt]
    x = ln(e + "LayerNorm", torch.cat((text, v_emb), dim=1))
    am = inp.get("attention_mask")
    mask = None if am is None else ((1.0 - am.to(x.dtype)) * -10000.0)[:, None, None, :]
    layers = [_BertLayer(sd, f"bert.encoder.layer.{i}.", H) for i in range(cfg.layers)]
    for layer in layers:
        x = layer.forward(x, mask)
    cls_index = inp["input_mask"].sum(1) - 2
    pooled = x.index_select(1, cls_index)                                 # vqa_pooler (IndexSelect)
    t_pre = lin("classifier.0.dense", pooled)
    h = ln("classifier.0.LayerNorm", F.gelu(t_pre))
    scores = lin("classifier.1", h).contiguous().view(-1, cfg.num_labels)
    idx = int(scores.argmax(-1)) if index is None else int(index)
    cam = torch.zeros(1, cfg.num_labels, dtype=scores.dtype)
    cam[0, idx] = 1
    cam = _lx_linear_relprop(cam, h, sd["classifier.1.weight"])
    cam = _lx_linear_relprop(cam, pooled, sd["classifier.0.dense.weight"])      # LayerNorm / GELU pass through
    cam = index_select_relprop(cam, x, 1, cls_index)
    for layer in reversed(layers):
        cam = layer.relprop(cam)
    return scores, layers
