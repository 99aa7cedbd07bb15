"""The LRP (``relprop``) sweep behind ``use_lrp=True`` / ``generate_transformer_att`` / ``generate_partial_lrp``
 over the ``mmx_lrp_*`` kernels and the GEMM backends of libmmx.

The reference propagates relevance with one autograd call per module (gradient x input, DETR/modules/layers.py:38-66);
here every rule is a direct kernel sequence on the activations the forward tape already holds (nothing is recomputed, no
autograd).  The reference runs relprop with batch 1 and several rules normalise by sums over the whole tensor; a batch
here is B independent samples and those sums are taken per sample, so a sample gets the same relevance alone or inside a
batch.

Rules (DETR flavour DETR/modules/layers.py, LXMERT / VisualBERT flavour lxmert/lxmert/src/layers.py - they differ only in
the renormalisation that ends DETR's Linear.relprop):
  Linear :409-432   Add :194-221   Clone :252-270   IndexSelect :230-249   einsum / MatMul (RelPropSimple) :38-66,223-228
  MultiheadAttention.relprop :770-801, LxmertAttention.relprop lxmert_lrp.py:422-461, BertSelfAttention.relprop
  VisualBERT/mmf/models/transformers/backends/BERT_ours.py:352-395.
LayerNorm / activations / dropout / softmax / scalar scaling pass relevance through unchanged (:48-49, :107-108).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from ._lib import lib, check, ptr, current_stream
from .nn import Weight, AttnRecord, ACT_NONE

ACT_MUL = 5


class _LrpWeight:
    """Positive / negative parts of W [N,K] and of W^T [K,N] (the K-major operand of the `S W` product)."""

    def __init__(self, sweep: "Sweep", W: Weight):
        self.pw, self.nw = sweep.split(W.w)
        self.pwt, self.nwt = sweep.split(W.wt)
        self.N, self.K = W.w.shape


class Sweep:
    """Kernel-level relprop rules for a batch of B samples; matrices are [B * rows_per_sample, cols] fp32."""

    def __init__(self, device, B: int):
        self.device = torch.device(device)
        self.B = B
        self.lib = lib()
        self.nch = self.lib.mmx_lrp_partials_len()

    # ------------------------------------------------------------------ small helpers
    def new(self, rows, cols) -> torch.Tensor:
        return torch.empty(rows, cols, device=self.device, dtype=torch.float32)

    def split(self, X: torch.Tensor):
        P, N = torch.empty_like(X), torch.empty_like(X)
        check(self.lib.mmx_lrp_split(ptr(X), X.stride(0), ptr(P), ptr(N), P.stride(0), X.shape[0], X.shape[1], current_stream()))
        return P, N

    def weight(self, W: Weight) -> _LrpWeight:
        lw = getattr(W, "_lrp", None)
        if lw is None:
            lw = W._lrp = _LrpWeight(self, W)
        return lw

    def sums(self, X: torch.Tensor) -> torch.Tensor:
        """Per-sample (sum, abs-sum) partials of X [B*rows, cols]."""
        part = torch.empty(self.B, self.nch, 2, device=self.device, dtype=torch.float64)
        check(self.lib.mmx_lrp_sums(ptr(X), X.stride(0), X.shape[0] // self.B, X.shape[1], self.B, ptr(part), current_stream()))
        return part

    def _gemm(self, A, Bt, pre=None, residual=None, act=ACT_NONE):
        M, K = A.shape
        N = Bt.shape[0]
        out = self.new(M, N)
        check(self.lib.mmx_gemm_nt(ptr(A), A.stride(0), ptr(Bt), Bt.stride(0), None, None, ptr(pre), 0 if pre is None else pre.stride(0),
                                   ptr(residual), 0 if residual is None else residual.stride(0), ptr(out), N, None, act, M, N, K,
                                   current_stream()))
        return out

    # ------------------------------------------------------------------ rules
    def linear(self, R: torch.Tensor, X: torch.Tensor, W: Weight, renorm: bool) -> torch.Tensor:
        """Linear.relprop with alpha = 1, beta = 0 (DETR/modules/layers.py:409-432; lxmert/lxmert/src/layers.py:219-242):
        Z = x+ W+^T + x- W-^T, S = safe_divide(R, Z), out = x+ (S W+) + x- (S W-); ``renorm`` adds DETR's
        ``out * safe_divide(R.sum(), out.sum())``."""
        lw = self.weight(W)
        px, nx = self.split(X)
        Z = self._gemm(nx, lw.nw, residual=self._gemm(px, lw.pw))
        S = torch.empty_like(Z)
        check(self.lib.mmx_lrp_safe_divide(ptr(R), R.stride(0), ptr(Z), Z.stride(0), ptr(S), S.stride(0), Z.shape[0], Z.shape[1],
                                           current_stream()))
        out = self._gemm(S, lw.nwt, pre=nx, act=ACT_MUL, residual=self._gemm(S, lw.pwt, pre=px, act=ACT_MUL))
        if renorm:
            check(self.lib.mmx_lrp_renorm(ptr(out), out.stride(0), out.shape[0] // self.B, out.shape[1], self.B, ptr(self.sums(R)),
                                          ptr(self.sums(out)), current_stream()))
        return out

    def add(self, R, x0, x1, want_b: bool = True):
        """Add.relprop (layers.py:194-221): the relevances of the two summands."""
        a = torch.empty_like(x0)
        b = torch.empty_like(x1) if want_b else None
        ws = torch.empty(self.B, self.nch, 3, device=self.device, dtype=torch.float64)
        check(self.lib.mmx_lrp_add(ptr(R), R.stride(0), ptr(x0), x0.stride(0), ptr(x1), x1.stride(0), ptr(a), a.stride(0), ptr(b),
                                   0 if b is None else b.stride(0), x0.shape[0] // self.B, x0.shape[1], self.B, ptr(ws),
                                   current_stream()))
        return a, b

    def clone(self, Rs: Sequence[torch.Tensor], X: torch.Tensor) -> torch.Tensor:
        """Clone.relprop (layers.py:252-270): X * sum_i safe_divide(R_i, X)."""
        n = len(Rs)
        out = torch.empty_like(X)
        pa = (C.c_void_p * n)(*[r.data_ptr() for r in Rs])
        la = (C.c_int * n)(*[r.stride(0) for r in Rs])
        check(self.lib.mmx_lrp_clone(ptr(X), X.stride(0), pa, la, n, ptr(out), out.stride(0), X.shape[0], X.shape[1], current_stream()))
        return out

    def plus(self, a, b):
        out = torch.empty_like(a)
        check(self.lib.mmx_add(ptr(a), a.stride(0), ptr(b), b.stride(0), C.c_float(1.0), ptr(out), out.stride(0), a.shape[0],
                               a.shape[1], current_stream()))
        return out

    def attn_pv(self, R_o, rec: AttnRecord):
        """RelPropSimple through ``attn @ v``; stages the relevance of the probabilities (``attn_cam``) in the record."""
        sv = rec.saved
        B, H, T, S, hd = sv["B"], sv["H"], sv["T"], sv["S"], sv["hd"]
        A, v, o = rec.A, sv["v"].v, sv["o"].v
        cam_A = torch.empty_like(A)
        cam_v = torch.empty_like(v)
        check(self.lib.mmx_lrp_attn_pv(ptr(R_o), R_o.stride(0), ptr(o), o.stride(0), ptr(A), A.shape[-1], ptr(v), v.stride(0),
                                       ptr(cam_A), ptr(cam_v), cam_v.stride(0), B, H, T, S, hd, current_stream()))
        rec.cam = cam_A
        return cam_A, cam_v

    def attn_qk(self, cam1, rec: AttnRecord, zscale: float):
        sv = rec.saved
        B, H, T, S, hd = sv["B"], sv["H"], sv["T"], sv["S"], sv["hd"]
        q, k = sv["q"].v, sv["k"].v
        cam_q, cam_k = torch.empty_like(q), torch.empty_like(k)
        check(self.lib.mmx_lrp_attn_qk(ptr(cam1), cam1.shape[-1], ptr(q), q.stride(0), ptr(k), k.stride(0), C.c_float(zscale),
                                       ptr(cam_q), cam_q.stride(0), ptr(cam_k), cam_k.stride(0), B, H, T, S, hd, current_stream()))
        return cam_q, cam_k

    def add_scores(self, cam1, rec: AttnRecord, key_bias: torch.Tensor, zscale: float):
        """The Add between the attention scores and the additive mask as VisualBERT's BertSelfAttention.relprop sees it
        (BERT_ours.py:352-395): returns the share of ``cam1`` that goes to the scores."""
        sv = rec.saved
        B, H, T, S, hd = sv["B"], sv["H"], sv["T"], sv["S"], sv["hd"]
        q, k = sv["q"].v, sv["k"].v
        ld = cam1.shape[-1]
        scores, mask = torch.empty_like(cam1), torch.empty_like(cam1)
        check(self.lib.mmx_lrp_attn_scores(ptr(q), q.stride(0), ptr(k), k.stride(0), ptr(key_bias), C.c_float(zscale), ptr(scores),
                                           ptr(mask), ld, B, H, T, S, hd, current_stream()))
        a = torch.zeros_like(cam1)
        ws = torch.empty(self.B, self.nch, 3, device=self.device, dtype=torch.float64)
        check(self.lib.mmx_lrp_add(ptr(cam1), ld, ptr(scores), ld, ptr(mask), ld, ptr(a), ld, None, 0, H * T, S, B, ptr(ws),
                                   current_stream()))
        return a

    def zero_value_fix(self, cam_q, cam_k, v_pre, v_post, cam_sums):
        check(self.lib.mmx_lrp_zero_value_fix(ptr(cam_q), cam_q.stride(0), cam_q.shape[0] // self.B, ptr(cam_k), cam_k.stride(0),
                                              cam_k.shape[0] // self.B, cam_q.shape[1], self.B, ptr(v_pre), ptr(v_post),
                                              ptr(self.sums(cam_q)), ptr(self.sums(cam_k)), ptr(cam_sums), current_stream()))


# ======================================================================================================================
# DETR (DETR/models/transformer.py:68-79,166-199,256-275,410-436; DETR/models/detr.py:79-92)
# ======================================================================================================================
def _detr_mha(sw: Sweep, m, cam, stop_after_cam: bool = False):
    """MultiheadAttention.relprop (DETR/modules/layers.py:770-801) -> (cam_q, cam_k, cam_v)."""
    rec = m["rec"]
    sv = rec.saved
    cam = sw.linear(cam, sv["o"].v, m["o"], True)
    cam_A, cam_v = sw.attn_pv(cam, rec)
    if stop_after_cam:
        return None, None, None
    cam_q, cam_k = sw.attn_qk(cam_A, rec, float(sv["hd"]) ** -0.5)     # q is scaled before the product (:738)
    v_pre = sw.sums(cam_v)
    cam_v = sw.linear(cam_v, sv["Xv"].v, m["v"], True)
    cam_k = sw.linear(cam_k, sv["Xk"].v, m["k"], True)
    cam_q = sw.linear(cam_q, sv["Xq"].v, m["q"], True)
    sw.zero_value_fix(cam_q, cam_k, v_pre, sw.sums(cam_v), sw.sums(cam))
    return cam_q, cam_k, cam_v


def _detr_enc_layer(sw: Sweep, L, cam, first: bool):
    """TransformerEncoderLayer.forward_post_relprop (DETR/models/transformer.py:256-275)."""
    s = L.saved
    cam_2, cam2 = sw.add(cam, s["x1"].v, s["ff"].v)
    cam_1 = sw.linear(cam2, s["h"].v, L.l2, True)
    cam_1 = sw.linear(cam_1, s["x1"].v, L.l1, True)
    cam = sw.clone([cam_1, cam_2], s["x1"].v)
    cam_3, cam_drop = sw.add(cam, s["src"].v, s["drop"].v)
    cam_q, cam_k, cam_v = _detr_mha(sw, L.self_attn, cam_drop, stop_after_cam=first)
    if first:
        return None
    cam_w = sw.clone([cam_q, cam_k], s["webmd"].v)
    return sw.clone([cam_w, cam_v, cam_3], s["src"].v)


def _detr_dec_layer(sw: Sweep, L, cam):
    """TransformerDecoderLayer.forward_post_relprop (DETR/models/transformer.py:410-436) -> (cam_tgt, cam_memory)."""
    s = L.saved
    cam_2, cam2 = sw.add(cam, s["x2"].v, s["ff"].v)
    cam2 = sw.linear(cam2, s["h"].v, L.l2, True)
    cam_1 = sw.linear(cam2, s["x2"].v, L.l1, True)
    cam = sw.clone([cam_1, cam_2], s["x2"].v)
    cam_2, cam_drop = sw.add(cam, s["x1"].v, s["drop2"].v)
    cam_q, cam_k, cam_mem_2 = _detr_mha(sw, L.multihead_attn, cam_drop)
    cam_mem = sw.clone([cam_k, cam_mem_2], s["memory"].v)
    cam = sw.clone([cam_q, cam_2], s["x1"].v)
    cam_3, cam_drop = sw.add(cam, s["tgt"].v, s["drop1"].v)
    cam_q, cam_k, cam_v = _detr_mha(sw, L.self_attn, cam_drop)
    cam_w = sw.clone([cam_q, cam_k], s["webmd"].v)
    return sw.plus(sw.plus(cam_w, cam_v), cam_3), cam_mem                # the reference sums here (:434)


def detr_sweep(engine, one_hot: torch.Tensor):
    """DETR.relprop (DETR/models/detr.py:79-92) from the one-hot class logit; leaves ``rec.cam`` ([B,H,T,ld]) on every
    attention record of the engine.  ``one_hot``: [B*Q, C+1]."""
    sv = engine.saved
    B = sv["B"]
    with torch.cuda.device(engine.device):
        sw = Sweep(engine.device, B)
        cam = sw.clone([one_hot], sv["logits"].v)                       # IndexSelect.relprop: the last decoder level only
        cam = sw.linear(cam, sv["hs"].v, engine.class_embed, True)
        n = len(engine.decoder)
        cam_mem_list: List[torch.Tensor] = []
        for j in reversed(range(n)):                                    # TransformerDecoder.relprop (transformer.py:179-195)
            if j != n - 1:
                cam = sw.clone([cam], engine.decoder[j].saved["out"].v)  # clone of layer j's output; its own level carries zero
            cam, cam_mem_j = _detr_dec_layer(sw, engine.decoder[j], cam)
            cam_mem_list.append(cam_mem_j)
        memory = sv["memory"].v
        cam = sw.clone([sw.clone(cam_mem_list, memory)], memory)         # transformer.py:70-74 (the zero `mem` branch adds 0)
        for i in reversed(range(len(engine.encoder))):
            cam = _detr_enc_layer(sw, engine.encoder[i], cam, first=(i == 0))


# ======================================================================================================================
# LXMERT (lxmert/lxmert/src/lxmert_lrp.py) and VisualBERT (BERT_ours.py): Linear.relprop without renormalisation
# ======================================================================================================================
def _lx_attention(sw: Sweep, a, rec: AttnRecord, cam, stop_after_cam: bool = False):
    """LxmertAttention.relprop (lxmert_lrp.py:422-461) -> (cam_hidden, cam_context)."""
    sv = rec.saved
    cam1, cam2 = sw.attn_pv(cam, rec)
    if stop_after_cam:
        return None, None
    cam_q, cam_k = sw.attn_qk(cam1, rec, 1.0)                          # the scores are divided by sqrt(d) after the product
    cam_q = sw.linear(cam_q, sv["Xq"].v, a.q, False)
    cam_k = sw.linear(cam_k, sv["Xk"].v, a.k, False)
    cam_v = sw.linear(cam2, sv["Xv"].v, a.v, False)
    return cam_q, sw.clone([cam_k, cam_v], sv["Xk"].v)


def _lx_att_block(sw: Sweep, a, rec: AttnRecord, cam, self_attention: bool, stop_after_cam: bool = False):
    """LxmertSelfAttentionLayer / LxmertCrossAttentionLayer (+ LxmertAttentionOutput) relprop (lxmert_lrp.py:479-484,505-509,535-539)."""
    sv = rec.saved
    x = sv["Xq"].v
    cam_dense, cam_res = sw.add(cam, sv["dense"].v, x)
    cam_out = sw.linear(cam_dense, sv["o"].v, a.o, False)
    cam_hidden, cam_ctx = _lx_attention(sw, a, rec, cam_out, stop_after_cam)
    if stop_after_cam:
        return None, None
    if self_attention:
        return sw.clone([cam_hidden, cam_ctx, cam_res], x), None
    return sw.clone([cam_hidden, cam_res], x), cam_ctx


def _lx_ffn(sw: Sweep, f, saved, cam):
    """LxmertOutput + LxmertIntermediate + the clone around them (lxmert_lrp.py:554-557,575-580,601-606)."""
    x = saved["x"].v
    cam1, cam2 = sw.add(cam, saved["dense"].v, x)
    cam1 = sw.linear(cam1, saved["inter"].v, f.fc2, False)
    cam1 = sw.linear(cam1, x, f.fc1, False)
    return sw.clone([cam1, cam2], x)


def lxmert_sweep(engine, one_hot: torch.Tensor):
    """LxmertForQuestionAnswering.relprop (lxmert_lrp.py:1689-1693 -> :1253-1257 -> :855-863) from the one-hot answer."""
    sv = engine.saved
    B, T = sv["B"], sv["T"]
    with torch.cuda.device(engine.device):
        sw = Sweep(engine.device, B)
        cam = sw.linear(one_hot, sv["h"].v, engine.head3, False)          # LxmertVisualAnswerHead (:955-958): LayerNorm, GELU pass
        cam = sw.linear(cam, sv["pooled"].v, engine.head0, False)
        cam = sw.linear(cam, sv["first"].v, engine.pooler, False)         # LxmertPooler (:886-892): Tanh passes
        lang, vis = sv["lang"].v, sv["vis"].v
        # IndexSelect.relprop on token 0: X[:,0] * safe_divide(cam, X[:,0]) scattered into zeros
        first_rel = sw.clone([cam], sv["first"].v)
        cam_lang = torch.zeros_like(lang)
        check(sw.lib.mmx_scatter_add_rows(ptr(first_rel), first_rel.stride(0), ptr(sv["rows0"]), ptr(cam_lang), cam_lang.stride(0), B,
                                          lang.shape[1], current_stream()))
        cam_vis = torch.zeros_like(vis)
        nx = len(engine.x_layers)
        for i in reversed(range(nx)):                                   # LxmertXLayer.relprop (:657-733)
            b = engine.x_layers[i]
            s = b.saved
            cam_vis, cam_lang = _lx_ffn(sw, b.visn_ffn, s["visn_ffn"], cam_vis), _lx_ffn(sw, b.lang_ffn, s["lang_ffn"], cam_lang)
            cam_vis, _ = _lx_att_block(sw, b.visn_self, b.visn_self.recs[0], cam_vis, True)
            cam_lang, _ = _lx_att_block(sw, b.lang_self, b.lang_self.recs[0], cam_lang, True)
            cam_vis2, cam_lang2 = _lx_att_block(sw, b.cross, b.cross.recs[1], cam_vis, False)
            cam_lang1, cam_vis1 = _lx_att_block(sw, b.cross, b.cross.recs[0], cam_lang, False)
            cam_lang = sw.clone([cam_lang1, cam_lang2], s["lang_in"].v)
            cam_vis = sw.clone([cam_vis1, cam_vis2], s["vis_in"].v)
        for j in reversed(range(len(engine.r_layers))):
            b = engine.r_layers[j]
            cam_vis = _lx_ffn(sw, b.ffn, b.saved["ffn"], cam_vis)
            cam_vis, _ = _lx_att_block(sw, b.att, b.att.recs[0], cam_vis, True, stop_after_cam=(j == 0))
        for j in reversed(range(len(engine.layer))):
            b = engine.layer[j]
            cam_lang = _lx_ffn(sw, b.ffn, b.saved["ffn"], cam_lang)
            cam_lang, _ = _lx_att_block(sw, b.att, b.att.recs[0], cam_lang, True, stop_after_cam=(j == 0))


def visualbert_sweep(engine, one_hot: torch.Tensor):
    """VisualBERTForClassification.relprop (VisualBERT/mmf/models/visual_bert.py:398-403) from the one-hot answer:
    classifier -> IndexSelect at cls_index -> BertEncoder.relprop (BERT_ours.py:152-156) -> BertLayer.relprop (:506-514)."""
    sv = engine.saved
    B = sv["B"]
    with torch.cuda.device(engine.device):
        sw = Sweep(engine.device, B)
        cam = sw.linear(one_hot, sv["h"].v, engine.head_out, False)       # BertPredictionHeadTransform: LayerNorm, GELU pass
        cam = sw.linear(cam, sv["pooled"].v, engine.head_dense, False)
        x_last = sv["x"].v
        sel = sw.clone([cam], sv["pooled"].v)                             # IndexSelect.relprop at cls_index
        cam = torch.zeros_like(x_last)
        check(sw.lib.mmx_scatter_add_rows(ptr(sel), sel.stride(0), ptr(sv["rows"]), ptr(cam), cam.stride(0), B, x_last.shape[1],
                                          current_stream()))
        key_bias = sv["key_bias"]
        for i in reversed(range(len(engine.layers))):
            L = engine.layers[i]
            s, rec = L.saved, L.rec
            x = s["x"].v
            cam = _lx_ffn(sw, L, s["ffn"], cam)                           # BertOutput + BertIntermediate + clone
            cam_dense, cam_res = sw.add(cam, s["dense"].v, x)             # BertSelfOutput (:412-419)
            cam_ctx = sw.linear(cam_dense, rec.saved["o"].v, L.o, False)
            cam1, cam2 = sw.attn_pv(cam_ctx, rec)                         # BertSelfAttention.relprop (:352-395)
            if i == 0:
                break                                                     # every attn_cam is staged; the rest feeds nothing
            if key_bias is not None:
                cam1 = sw.add_scores(cam1, rec, key_bias, float(rec.saved["hd"]) ** -0.5)
            cam_q, cam_k = sw.attn_qk(cam1, rec, 1.0)
            cam_q = sw.linear(cam_q, x, L.q, False)
            cam_k = sw.linear(cam_k, x, L.k, False)
            cam_v = sw.linear(cam2, x, L.v, False)
            cam = sw.clone([sw.clone([cam_q, cam_k, cam_v], x), cam_res], x)
