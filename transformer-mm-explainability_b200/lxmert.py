"""LXMERT bi-modal relevancy: ``GeneratorOurs(model_usage).generate_ours(...)`` with
the reference signature (lxmert/lxmert/src/ExplanationGenerator.py:131-211) over the libmmx kernels.

``LxmertEngine`` is built from a ``LxmertForQuestionAnswering`` state_dict (lxmert/lxmert/src/lxmert_lrp.py:1532) and
runs on (input_ids, visual_feats, visual_pos); the Faster-RCNN feature extractor and the tokenizer of the
reference's ``ModelUsage`` sit below every attention layer and are outside the hot-path scope.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import torch

from ._lib import lib, check, ptr, current_stream, MmxError
from .nn import fp32_gemms, Tape, Var, Weight, AttnRecord, ACT_GELU, ACT_TANH, ATTN_SCALE_SCORES, _f32
from . import rules

EPS = 1e-12


class _Att:
    """LxmertAttention (+ LxmertAttentionOutput) weights with one A/dA record per direction of use."""

    def __init__(self, sd, p_att, p_out, device, n_records=1):
        self.q = Weight(sd[p_att + "query.weight"], sd[p_att + "query.bias"], device)
        self.k = Weight(sd[p_att + "key.weight"], sd[p_att + "key.bias"], device)
        self.v = Weight(sd[p_att + "value.weight"], sd[p_att + "value.bias"], device)
        self.o = Weight(sd[p_out + "dense.weight"], sd[p_out + "dense.bias"], device)
        self.ln = (_f32(sd[p_out + "LayerNorm.weight"], device), _f32(sd[p_out + "LayerNorm.bias"], device))
        self.recs = [AttnRecord() for _ in range(n_records)]


class _Ffn:
    def __init__(self, sd, p_inter, p_out, device):
        self.fc1 = Weight(sd[p_inter + "dense.weight"], sd[p_inter + "dense.bias"], device)
        self.fc2 = Weight(sd[p_out + "dense.weight"], sd[p_out + "dense.bias"], device)
        self.ln = (_f32(sd[p_out + "LayerNorm.weight"], device), _f32(sd[p_out + "LayerNorm.bias"], device))


class _Block:
    pass


class LxmertEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], num_heads: int = 12, device=None):
        if not torch.cuda.is_available():
            raise MmxError("mmx_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
        self.device = d = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        sd = state_dict
        self.heads = num_heads
        e = "lxmert.embeddings."
        self.word = _f32(sd[e + "word_embeddings.weight"], d)
        self.posemb = _f32(sd[e + "position_embeddings.weight"], d)
        self.type0 = _f32(sd[e + "token_type_embeddings.weight"][0:1], d)
        self.emb_ln = (_f32(sd[e + "LayerNorm.weight"], d), _f32(sd[e + "LayerNorm.bias"], d))
        v = "lxmert.encoder.visn_fc."
        self.visn_fc = Weight(sd[v + "visn_fc.weight"], sd[v + "visn_fc.bias"], d)
        self.box_fc = Weight(sd[v + "box_fc.weight"], sd[v + "box_fc.bias"], d)
        self.visn_ln = (_f32(sd[v + "visn_layer_norm.weight"], d), _f32(sd[v + "visn_layer_norm.bias"], d))
        self.box_ln = (_f32(sd[v + "box_layer_norm.weight"], d), _f32(sd[v + "box_layer_norm.bias"], d))
        self.hidden = self.word.shape[1]
        count = lambda pre: len({k[len(pre):].split(".")[0] for k in sd if k.startswith(pre)})
        self.layer, self.r_layers, self.x_layers = [], [], []
        for name, dst in (("layer", self.layer), ("r_layers", self.r_layers)):
            for i in range(count(f"lxmert.encoder.{name}.")):
                p = f"lxmert.encoder.{name}.{i}."
                b = _Block()
                b.att = _Att(sd, p + "attention.self.", p + "attention.output.", d)
                b.ffn = _Ffn(sd, p + "intermediate.", p + "output.", d)
                dst.append(b)
        for i in range(count("lxmert.encoder.x_layers.")):
            p = f"lxmert.encoder.x_layers.{i}."
            b = _Block()
            # records[0]: text queries over image keys (visual_attention); records[1]: the deepcopy direction
            # (visual_attention_copy, lxmert_lrp.py:640-641) - same weights, separate A / dA
            b.cross = _Att(sd, p + "visual_attention.att.", p + "visual_attention.output.", d, n_records=2)
            b.lang_self = _Att(sd, p + "lang_self_att.self.", p + "lang_self_att.output.", d)
            b.visn_self = _Att(sd, p + "visn_self_att.self.", p + "visn_self_att.output.", d)
            b.lang_ffn = _Ffn(sd, p + "lang_inter.", p + "lang_output.", d)
            b.visn_ffn = _Ffn(sd, p + "visn_inter.", p + "visn_output.", d)
            self.x_layers.append(b)
        self.pooler = Weight(sd["lxmert.pooler.dense.weight"], sd["lxmert.pooler.dense.bias"], d)
        self.head0 = Weight(sd["answer_head.logit_fc.0.weight"], sd["answer_head.logit_fc.0.bias"], d)
        self.head_ln = (_f32(sd["answer_head.logit_fc.2.weight"], d), _f32(sd["answer_head.logit_fc.2.bias"], d))
        self.head3 = Weight(sd["answer_head.logit_fc.3.weight"], sd["answer_head.logit_fc.3.bias"], d)
        self.question_answering_score: Optional[torch.Tensor] = None
        self.text_len = self.image_boxes_len = 0

    def eval(self):
        return self

    def zero_grad(self):
        return None

    # ------------------------------------------------------------------ building blocks
    def _att_layer(self, tape: Tape, a: _Att, rec: AttnRecord, x: Var, ctx: Var, B, T, S, key_bias) -> Var:
        """LxmertAttention + LxmertAttentionOutput (lxmert_lrp.py:385-420, :472-477)."""
        H = self.heads
        q, k, v = tape.linear(x, a.q), tape.linear(ctx, a.k), tape.linear(ctx, a.v)
        o = tape.attention(q, k, v, B, H, T, S, 1.0 / math.sqrt(self.hidden // H), ATTN_SCALE_SCORES, key_bias, rec)
        dense = tape.linear(o, a.o)
        rec.saved.update(Xq=x, Xk=ctx, Xv=ctx, dense=dense)
        return tape.layernorm(tape.add(dense, x), *a.ln, EPS)

    def _ffn(self, tape: Tape, f: _Ffn, x: Var, saved: Optional[dict] = None) -> Var:
        inter = tape.linear(x, f.fc1, ACT_GELU)
        dense = tape.linear(inter, f.fc2)
        if saved is not None:
            saved.update(x=x, inter=inter, dense=dense)
        return tape.layernorm(tape.add(dense, x), *f.ln, EPS)

    def forward_backward(self, *args, lrp: bool = False, **kwargs):
        """See ``_forward_backward``.  With ``lrp=True`` every GEMM of the call (forward, dgrad and the relprop sweep) runs
        on the fp32 FFMA backend (``nn.fp32_gemms``: the sweep amplifies tensor-core split rounding)."""
        if lrp:
            with fp32_gemms():
                return self._forward_backward(*args, lrp=True, **kwargs)
        return self._forward_backward(*args, lrp=False, **kwargs)

    def _forward_backward(self, input_ids, visual_feats, visual_pos, index=None, attention_mask=None,
                         visual_attention_mask=None, backward: bool = True, lang_key_bias=None, vis_key_bias=None,
                         lrp: bool = False):
        """Forward staging every A (29 attention maps for the base model), one-hot on the answer logit, backward
        staging every dA.  input_ids [B,T] int, visual_feats [B,I,F], visual_pos [B,I,4]."""
        dev, l = self.device, lib()
        with torch.cuda.device(dev):
            ids = input_ids.to(dev).long()
            B, T = ids.shape
            feats, boxes = _f32(visual_feats, dev), _f32(visual_pos, dev)
            I = feats.shape[1]
            Hd = self.hidden
            self.text_len, self.image_boxes_len = T, I
            tape = Tape(dev)
            bias_t = None if attention_mask is None else ((1.0 - _f32(attention_mask, dev)) * -10000.0).contiguous()
            bias_i = None if visual_attention_mask is None else ((1.0 - _f32(visual_attention_mask, dev)) * -10000.0).contiguous()
            if lang_key_bias is not None:          # additive key biases given directly ([B,T] / [B,I]); -inf removes a key
                bias_t = _f32(lang_key_bias, dev)
            if vis_key_bias is not None:
                bias_i = _f32(vis_key_bias, dev)
            # embeddings: (type + position) + word, LayerNorm  (lxmert_lrp.py:285-310).  No gradient is needed below
            # the first attention layer, so these run outside the tape.
            emb = torch.empty(B * T, Hd, device=dev)
            check(l.mmx_gather_rows(ptr(self.word), Hd, ptr(ids.to(torch.int32).reshape(-1).contiguous()), ptr(emb), Hd, B * T, Hd,
                                    current_stream()))
            tp = (self.type0 + self.posemb[:T]).repeat(B, 1).contiguous()      # constants of the model: host-side setup
            check(l.mmx_add(ptr(tp), Hd, ptr(emb), Hd, C.c_float(1.0), ptr(emb), Hd, B * T, Hd, current_stream()))
            lang = tape.layernorm(Var(emb), *self.emb_ln, EPS)
            x = tape.layernorm(tape.linear(Var(feats.reshape(B * I, -1)), self.visn_fc), *self.visn_ln, EPS)
            y = tape.layernorm(tape.linear(Var(boxes.reshape(B * I, -1)), self.box_fc), *self.box_ln, EPS)
            vis_sum = tape.add(x, y)                                            # (x + y) / 2   (:763)
            half = torch.empty_like(vis_sum.v)
            check(l.mmx_add(ptr(vis_sum.v), Hd, ptr(vis_sum.v), Hd, C.c_float(-0.5), ptr(half), Hd, B * I, Hd, current_stream()))
            vis = Var(half)
            for b in self.layer:                                                # 9 language layers (:818-823)
                b.saved = dict(ffn={})
                lang = self._ffn(tape, b.ffn, self._att_layer(tape, b.att, b.att.recs[0], lang, lang, B, T, T, bias_t), b.saved["ffn"])
            for b in self.r_layers:                                             # 5 relational layers (:826-831)
                b.saved = dict(ffn={})
                vis = self._ffn(tape, b.ffn, self._att_layer(tape, b.att, b.att.recs[0], vis, vis, B, I, I, bias_i), b.saved["ffn"])
            for b in self.x_layers:                                             # 5 cross layers (:701-733)
                b.saved = dict(lang_in=lang, vis_in=vis, lang_ffn={}, visn_ffn={})
                l2 = self._att_layer(tape, b.cross, b.cross.recs[0], lang, vis, B, T, I, bias_i)
                v2 = self._att_layer(tape, b.cross, b.cross.recs[1], vis, lang, B, I, T, bias_t)
                l3 = self._att_layer(tape, b.lang_self, b.lang_self.recs[0], l2, l2, B, T, T, bias_t)
                v3 = self._att_layer(tape, b.visn_self, b.visn_self.recs[0], v2, v2, B, I, I, bias_i)
                lang = self._ffn(tape, b.lang_ffn, l3, b.saved["lang_ffn"])
                vis = self._ffn(tape, b.visn_ffn, v3, b.saved["visn_ffn"])
            rows0 = torch.arange(B, device=dev, dtype=torch.int32) * T
            first = tape.gather_rows(lang, rows0)
            pooled = tape.linear(first, self.pooler, ACT_TANH)                  # (:876-884)
            h = tape.layernorm(tape.linear(pooled, self.head0, ACT_GELU), *self.head_ln, EPS)
            logits = tape.linear(h, self.head3)
            self.question_answering_score = logits.v
            self._shape = (B, T, I)
            self.saved = dict(B=B, T=T, I=I, lang=lang, vis=vis, first=first, rows0=rows0, pooled=pooled, h=h)
            if not backward:
                return self.question_answering_score
            idx = logits.v.argmax(-1) if index is None else torch.as_tensor(index, device=dev).reshape(B).long()
            one_hot = torch.zeros_like(logits.v)
            one_hot.scatter_(1, idx.reshape(B, 1), 1.0)                         # ExplanationGenerator.py:152-160 (graph-capturable)
            tape.seed(logits, one_hot, B)
            tape.backward()
            if lrp:
                from .lrp import lxmert_sweep
                lxmert_sweep(self, one_hot)
        return self.question_answering_score


class GeneratorOurs:
    """lxmert/lxmert/src/ExplanationGenerator.py:56-211.  ``model_usage`` is a :class:`LxmertEngine`; ``input`` is
    ``(input_ids, visual_feats, visual_pos)``."""

    def __init__(self, model_usage: LxmertEngine, save_visualization=False):
        if not isinstance(model_usage, LxmertEngine):
            raise MmxError("model_usage must be a mmx_b200.LxmertEngine")
        self.model_usage = model_usage
        self.save_visualization = save_visualization

    def _self(self, rec, lang: bool):
        cam = rules.avg_heads_record(rec, self.B, self.use_lrp)
        if lang:                                                               # EG:61-71 / :85-94
            self.R_t_t, self.R_t_i = rules.self_update(self.R_t_t, cam, self.R_t_i)
        else:                                                                  # EG:73-83 / :96-105
            self.R_i_i, self.R_i_t = rules.self_update(self.R_i_i, cam, self.R_i_t)

    def _mm(self, R_ss, R_qq, R_qs, rec):
        cam = rules.avg_heads_record(rec, self.B, self.use_lrp)
        sq, ss, md = rules.mm_update_batched(R_ss, R_qq, R_qs, cam, self.normalize_self_attention, self.apply_self_in_rule_10)
        self._min_diag.append(md)
        return sq, ss

    def generate_ours(self, input, index=None, use_lrp=True, normalize_self_attention=True, apply_self_in_rule_10=True,
                      method_name="ours"):
        self.use_lrp = bool(use_lrp)
        self.normalize_self_attention = normalize_self_attention
        self.apply_self_in_rule_10 = apply_self_in_rule_10
        m = self.model_usage
        ids, feats, boxes = input
        m.forward_backward(ids, feats, boxes, index, lrp=self.use_lrp)
        B, T, I = m._shape
        self.B = B
        dev = m.device
        self._min_diag: List[torch.Tensor] = []
        self.R_t_t = torch.eye(T, device=dev).repeat(B, 1, 1)                   # EG:144-150
        self.R_i_i = torch.eye(I, device=dev).repeat(B, 1, 1)
        self.R_t_i = torch.zeros(B, T, I, device=dev)
        self.R_i_t = torch.zeros(B, I, T, device=dev)
        for b in m.layer:
            self._self(b.att.recs[0], True)
        for b in m.r_layers:
            self._self(b.att.recs[0], False)
        n = len(m.x_layers)
        for i, b in enumerate(m.x_layers):
            last = i == n - 1                                                    # EG:181-182, :200-207
            ti_add, tt_add = self._mm(self.R_t_t, self.R_i_i, self.R_i_t, b.cross.recs[0])
            if not last:
                it_add, ii_add = self._mm(self.R_i_i, self.R_t_t, self.R_t_i, b.cross.recs[1])
            self.R_t_i, self.R_t_t = rules.add(self.R_t_i, ti_add), rules.add(self.R_t_t, tt_add)
            if not last:
                self.R_i_t, self.R_i_i = rules.add(self.R_i_t, it_add), rules.add(self.R_i_i, ii_add)
            self._self(b.lang_self.recs[0], True)
            if not last:
                self._self(b.visn_self.recs[0], False)
        self.min_diag = torch.stack(self._min_diag).min() if normalize_self_attention else None
        if not getattr(self, "_defer_assert", False):
            self.check_min_diag()
        self.R_t_t[:, 0, 0] = 0                                                  # EG:210
        if B == 1:
            self.R_t_t, self.R_t_i, self.R_i_i, self.R_i_t = self.R_t_t[0], self.R_t_i[0], self.R_i_i[0], self.R_i_t[0]
        return self.R_t_t, self.R_t_i

    def check_min_diag(self):
        """handle_residual's ``assert self_attention[diag].min() >= 0`` (EG:50) over every cross-modal step of the last call
        (one host read; deferred to after the replay when the call runs as a CUDA graph)."""
        if getattr(self, "min_diag", None) is not None:
            assert self.min_diag.item() >= 0

    def capture(self, input, index=None, use_lrp=True, normalize_self_attention=True, apply_self_in_rule_10=True):
        """``generate_ours`` for these shapes / flags as a CUDA graph (``mmx_b200.graphs.Graphed``): returns a callable
        ``g(input_ids, visual_feats, visual_pos)`` that replays the captured kernels on new inputs and yields the (static)
        ``(R_t_t, R_t_i)``; ``g.check()`` runs the deferred ``diag(R - I) >= 0`` assert."""
        from .graphs import Graphed
        dev = self.model_usage.device
        idx = None if index is None else torch.as_tensor(index).reshape(-1).to(dev)

        def fn(ids, feats, boxes):
            return self.generate_ours((ids, feats, boxes), idx, use_lrp, normalize_self_attention, apply_self_in_rule_10)
        self._defer_assert = True
        try:
            return Graphed(fn, tuple(input), dev, deferred_checks=[self.check_min_diag])
        finally:
            self._defer_assert = False


class GeneratorBaselines:
    """lxmert/lxmert/src/ExplanationGenerator.py:368-666: raw attention, attn-GradCAM and rollout (the LRP based
    ``generate_transformer_attr`` / ``generate_partial_lrp`` need relprop and raise)."""

    def __init__(self, model_usage: LxmertEngine, save_visualization=False):
        if not isinstance(model_usage, LxmertEngine):
            raise MmxError("model_usage must be a mmx_b200.LxmertEngine")
        self.model_usage = model_usage
        self.save_visualization = save_visualization

    def _finish(self):
        self.R_t_t[:, 0, 0] = 0                                                    # "disregard the [CLS] token itself"
        if self.R_t_t.shape[0] == 1:
            self.R_t_t, self.R_t_i = self.R_t_t[0], self.R_t_i[0]
        return self.R_t_t, self.R_t_i

    def generate_raw_attn(self, input, method_name="raw_attention"):               # EG:508-540
        m = self.model_usage
        m.forward_backward(*input, backward=False)
        B = m._shape[0]
        blk = m.x_layers[-1]
        self.R_t_i = rules.head_mean_record(blk.cross.recs[0], B).contiguous()
        self.R_t_t = rules.head_mean_record(blk.lang_self.recs[0], B).contiguous()
        return self._finish()

    def generate_attn_gradcam(self, input, index=None, method_name="gradcam"):     # EG:549-593
        m = self.model_usage
        m.forward_backward(*input, index=index)
        B = m._shape[0]
        blk = m.x_layers[-1]
        self.R_t_i = rules.gradcam_record(blk.cross.recs[0], B)
        self.R_t_t = rules.gradcam_record(blk.lang_self.recs[0], B)
        return self._finish()

    def generate_rollout(self, input, method_name="rollout"):                      # EG:595-666
        m = self.model_usage
        m.forward_backward(*input, backward=False)
        B = m._shape[0]
        hm = lambda rec: rules.head_mean_record(rec, B)
        cams_text = [hm(b.att.recs[0]) for b in m.layer]
        cams_image = [hm(b.att.recs[0]) for b in m.r_layers]
        for b in m.x_layers[:-1]:
            cams_text.append(hm(b.lang_self.recs[0]))
            cams_image.append(hm(b.visn_self.recs[0]))
        last = m.x_layers[-1]
        cam_t_i = hm(last.cross.recs[0])
        R_t_t = rules.compute_rollout_attention(cams_text)
        self.R_i_i = rules.compute_rollout_attention(cams_image)
        self.R_t_i, _, _ = rules.mm_update_batched(R_t_t, self.R_i_i, None, cam_t_i, apply_normalization=False,
                                                   apply_self_in_rule_10=True)
        cams_text.append(hm(last.lang_self.recs[0]))
        self.R_t_t = rules.compute_rollout_attention(cams_text)
        return self._finish()

    def generate_transformer_attr(self, input, index=None, method_name="transformer_attr"):      # EG:373-457
        """Transformer attribution: rule 5 with the LRP relevance over the self-attention layers only; R_t_i is the last
        cross layer's rule-5 map."""
        m = self.model_usage
        m.forward_backward(*input, index=index, lrp=True)
        B, T, I = m._shape
        dev = m.device
        R_t_t = torch.eye(T, device=dev).repeat(B, 1, 1)
        R_i_i = torch.eye(I, device=dev).repeat(B, 1, 1)
        cam_of = lambda rec: rules.avg_heads_record(rec, B, use_cam=True)
        for b in m.layer:
            R_t_t, _ = rules.self_update(R_t_t, cam_of(b.att.recs[0]))
        for b in m.r_layers:
            R_i_i, _ = rules.self_update(R_i_i, cam_of(b.att.recs[0]))
        for b in m.x_layers[:-1]:
            R_t_t, _ = rules.self_update(R_t_t, cam_of(b.lang_self.recs[0]))
            R_i_i, _ = rules.self_update(R_i_i, cam_of(b.visn_self.recs[0]))
        last = m.x_layers[-1]
        self.R_t_i = cam_of(last.cross.recs[0]).contiguous()
        R_t_t, _ = rules.self_update(R_t_t, cam_of(last.lang_self.recs[0]))
        self.R_t_t, self.R_i_i = R_t_t.contiguous(), R_i_i
        return self._finish()

    def generate_partial_lrp(self, input, index=None, method_name="partial_lrp"):                  # EG:459-506
        """Partial LRP: head mean of the LRP relevance of the last cross layer (text -> image and the text self-attention),
        each min-max normalised."""
        m = self.model_usage
        m.forward_backward(*input, index=index, lrp=True)
        B = m._shape[0]
        last = m.x_layers[-1]
        self.R_t_i = rules.minmax_normalize(rules.head_mean_record(last.cross.recs[0], B, use_cam=True).contiguous())
        self.R_t_t = rules.minmax_normalize(rules.head_mean_record(last.lang_self.recs[0], B, use_cam=True).contiguous())
        return self._finish()


class GeneratorOursAblationNoAggregation:
    """lxmert/lxmert/src/ExplanationGenerator.py:215-365: ``generate_ours_no_agg`` - the updates replace the relevancy
    matrices instead of adding to them."""

    def __init__(self, model_usage: LxmertEngine, save_visualization=False):
        if not isinstance(model_usage, LxmertEngine):
            raise MmxError("model_usage must be a mmx_b200.LxmertEngine")
        self.model_usage = model_usage
        self.save_visualization = save_visualization

    def _self(self, rec, lang: bool):                                            # EG:220-268
        cam = rules.avg_heads_record(rec, self.B, self.use_lrp)
        if lang:
            self.R_t_t, self.R_t_i = rules.bmm(cam, self.R_t_t), rules.bmm(cam, self.R_t_i)
        else:
            self.R_i_i, self.R_i_t = rules.bmm(cam, self.R_i_i), rules.bmm(cam, self.R_i_t)

    def _mm(self, R_ss, R_qq, R_qs, rec):                                        # EG:270-288 (rule 10 always with R_ss, R_qq)
        cam = rules.avg_heads_record(rec, self.B, self.use_lrp)
        sq, ss, md = rules.mm_update_batched(R_ss, R_qq, R_qs, cam, self.normalize_self_attention, True)
        self._min_diag.append(md)
        return sq, ss

    def generate_ours_no_agg(self, input, index=None, use_lrp=False, normalize_self_attention=True, method_name="ours_no_agg"):
        self.use_lrp = bool(use_lrp)
        self.normalize_self_attention = normalize_self_attention
        m = self.model_usage
        ids, feats, boxes = input
        m.forward_backward(ids, feats, boxes, index, lrp=self.use_lrp)
        B, T, I = m._shape
        self.B = B
        dev = m.device
        self._min_diag: List[torch.Tensor] = []
        self.R_t_t = torch.eye(T, device=dev).repeat(B, 1, 1)                   # EG:296-304
        self.R_i_i = torch.eye(I, device=dev).repeat(B, 1, 1)
        self.R_t_i = torch.zeros(B, T, I, device=dev)
        self.R_i_t = torch.zeros(B, I, T, device=dev)
        for b in m.layer:
            self._self(b.att.recs[0], True)
        for b in m.r_layers:
            self._self(b.att.recs[0], False)
        for b in m.x_layers[:-1]:                                                # EG:330-350
            ti, tt = self._mm(self.R_t_t, self.R_i_i, self.R_i_t, b.cross.recs[0])
            it, ii = self._mm(self.R_i_i, self.R_t_t, self.R_t_i, b.cross.recs[1])
            self.R_t_i, self.R_t_t, self.R_i_t, self.R_i_i = ti, tt, it, ii
            self._self(b.lang_self.recs[0], True)
            self._self(b.visn_self.recs[0], False)
        last = m.x_layers[-1]                                                    # EG:353-361: text side only
        self.R_t_i, self.R_t_t = self._mm(self.R_t_t, self.R_i_i, self.R_i_t, last.cross.recs[0])
        self._self(last.lang_self.recs[0], True)
        if normalize_self_attention:
            assert torch.stack(self._min_diag).min().item() >= 0                 # handle_residual's assert (EG:50)
        self.R_t_t[:, 0, 0] = 0                                                  # EG:364
        if B == 1:
            self.R_t_t, self.R_t_i, self.R_i_i, self.R_i_t = self.R_t_t[0], self.R_t_i[0], self.R_i_i[0], self.R_i_t[0]
        return self.R_t_t, self.R_t_i
