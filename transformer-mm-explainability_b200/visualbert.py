"""VisualBERT single-stream relevancy: ``SelfAttentionGenerator(model).generate_ours(input)`` with
the reference signature (VisualBERT/mmf/models/transformers/backends/ExplanationGenerator.py:20-214) over the libmmx
kernels.

``VisualBertEngine`` is built from a ``VisualBERTForClassification`` state_dict (VisualBERT/mmf/models/visual_bert.py:263;
keys ``bert.embeddings.*``, ``bert.encoder.layer.N.*``, ``classifier.{0,1}.*``, with or without the mmf ``model.``
prefix) and runs on the post-flatten sample fields the mmf model passes down (visual_bert.py:590-599): ``input_ids``,
``token_type_ids``, ``input_mask``, ``visual_embeddings``, ``visual_embeddings_type``, ``attention_mask``.  Unlike the
reference (one sample per call) the engine accepts a batch; every sample keeps its own ``cls_index``.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import torch

from ._lib import lib, check, ptr, current_stream, MmxError
from .nn import fp32_gemms, Tape, Var, Weight, AttnRecord, ACT_GELU, ATTN_SCALE_SCORES, _f32
from . import rules

EPS = 1e-12


class _Layer:
    """BertLayer weights (BERT_ours.py:475-506) with the A / dA record of its self-attention."""

    def __init__(self, sd, p, device):
        W = lambda n: Weight(sd[p + n + ".weight"], sd[p + n + ".bias"], device)
        LN = lambda n: (_f32(sd[p + n + ".weight"], device), _f32(sd[p + n + ".bias"], device))
        self.q, self.k, self.v = W("attention.self.query"), W("attention.self.key"), W("attention.self.value")
        self.o, self.ln1 = W("attention.output.dense"), LN("attention.output.LayerNorm")
        self.fc1, self.fc2, self.ln2 = W("intermediate.dense"), W("output.dense"), LN("output.LayerNorm")
        self.rec = AttnRecord()


class VisualBertEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], num_heads: int = 12, device=None):
        if not torch.cuda.is_available():
            raise MmxError("mmx_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
        self.device = d = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        sd = state_dict
        if not any(k.startswith("bert.") for k in sd) and any(k.startswith("model.bert.") for k in sd):
            sd = {k[len("model."):]: v for k, v in sd.items() if k.startswith("model.")}
        self.heads = num_heads
        e = "bert.embeddings."
        emb = lambda n: _f32(sd[e + n + ".weight"], d)
        self.word, self.posemb, self.typeemb = emb("word_embeddings"), emb("position_embeddings"), emb("token_type_embeddings")
        self.pos_vis, self.type_vis = emb("position_embeddings_visual"), emb("token_type_embeddings_visual")
        self.emb_ln = (_f32(sd[e + "LayerNorm.weight"], d), _f32(sd[e + "LayerNorm.bias"], d))
        self.projection = Weight(sd[e + "projection.weight"], sd[e + "projection.bias"], d)
        self.hidden = self.word.shape[1]
        n_layers = len({k.split(".")[3] for k in sd if k.startswith("bert.encoder.layer.")})
        self.layers: List[_Layer] = [_Layer(sd, f"bert.encoder.layer.{i}.", d) for i in range(n_layers)]
        self.head_dense = Weight(sd["classifier.0.dense.weight"], sd["classifier.0.dense.bias"], d)
        self.head_ln = (_f32(sd["classifier.0.LayerNorm.weight"], d), _f32(sd["classifier.0.LayerNorm.bias"], d))
        self.head_out = Weight(sd["classifier.1.weight"], sd["classifier.1.bias"], d)
        self.scores: Optional[torch.Tensor] = None
        self.cls_index: Optional[torch.Tensor] = None
        self._shape = (0, 0)

    def eval(self):
        return self

    def zero_grad(self):
        return None

    def forward_backward(self, *args, lrp: bool = False, **kwargs):
        """See ``_forward_backward``.  With ``lrp=True`` every GEMM of the call (forward, dgrad and the relprop sweep) runs
        on the fp32 FFMA backend (``nn.fp32_gemms``: the sweep amplifies tensor-core split rounding)."""
        if lrp:
            with fp32_gemms():
                return self._forward_backward(*args, lrp=True, **kwargs)
        return self._forward_backward(*args, lrp=False, **kwargs)

    def _forward_backward(self, input: Dict[str, torch.Tensor], index=None, backward: bool = True, lrp: bool = False) -> torch.Tensor:
        """Forward staging A of every layer, one-hot on the answer score, backward staging dA (EG:68-81); ``lrp=True``
        adds the relprop sweep from the same one-hot (visual_bert.py:398-403), staging the relevance of every A."""
        dev, l = self.device, lib()
        with torch.cuda.device(dev):
            ids = input["input_ids"].to(dev).long()
            B, T = ids.shape
            vis = _f32(input["visual_embeddings"], dev)
            V = vis.shape[1]
            S, Hd, H = T + V, self.hidden, self.heads
            tt = input.get("token_type_ids")
            tt = torch.zeros_like(ids) if tt is None else tt.to(dev).long()
            vt = input.get("visual_embeddings_type")
            vt = torch.zeros(B, V, dtype=torch.long, device=dev) if vt is None else vt.to(dev).long()
            # BertVisioLinguisticEmbeddings (mmf/modules/embeddings.py:325-451, plain strategy).  Word rows and the
            # visual projection are written straight into the [B, T+V, hidden] sequence; the position / type rows are
            # table look-ups of the model's constants.
            emb = torch.empty(B, S, Hd, device=dev)
            ids32 = ids.to(torch.int32).contiguous()
            for b in range(B):
                check(l.mmx_gather_rows(ptr(self.word), Hd, ptr(ids32[b]), ptr(emb[b]), Hd, T, Hd, current_stream()))
                check(l.mmx_linear(ptr(vis[b]), vis.shape[2], ptr(self.projection.w), vis.shape[2], ptr(self.projection.b), None, 0,
                                   ptr(emb[b, T:]), Hd, None, 0, V, Hd, vis.shape[2], current_stream()))
            const = torch.cat((self.posemb[:T].unsqueeze(0) + self.typeemb[tt], self.pos_vis[0] + self.type_vis[vt]), dim=1).contiguous()
            check(l.mmx_add(ptr(emb), Hd, ptr(const), Hd, C.c_float(1.0), ptr(emb), Hd, B * S, Hd, current_stream()))
            am = input.get("attention_mask")
            key_bias = None if am is None else ((1.0 - _f32(am, dev)) * -10000.0).contiguous()      # visual_bert.py:85-97
            if input.get("key_bias") is not None:      # additive key bias given directly [B, T+V]; -inf removes a key
                key_bias = _f32(input["key_bias"], dev)
            tape = Tape(dev)
            x = tape.layernorm(Var(emb.view(B * S, Hd)), *self.emb_ln, EPS)
            scale = 1.0 / math.sqrt(Hd // H)
            for L in self.layers:
                q, k, v = tape.linear(x, L.q), tape.linear(x, L.k), tape.linear(x, L.v)
                o = tape.attention(q, k, v, B, H, S, S, scale, ATTN_SCALE_SCORES, key_bias, L.rec)       # BERT_ours.py:322-338
                dense = tape.linear(o, L.o)
                a = tape.layernorm(tape.add(dense, x), *L.ln1, EPS)
                inter = tape.linear(a, L.fc1, ACT_GELU)
                dense2 = tape.linear(inter, L.fc2)
                L.saved = dict(x=x, dense=dense, ffn=dict(x=a, inter=inter, dense=dense2))
                x = tape.layernorm(tape.add(dense2, a), *L.ln2, EPS)
            self.cls_index = input["input_mask"].to(dev).long().sum(1) - 2                             # visual_bert.py:382
            rows = (torch.arange(B, device=dev) * S + self.cls_index).to(torch.int32)
            pooled = tape.gather_rows(x, rows)
            h = tape.layernorm(tape.linear(pooled, self.head_dense, ACT_GELU), *self.head_ln, EPS)
            scores = tape.linear(h, self.head_out)
            self.scores, self._shape = scores.v, (B, S)
            self.saved = dict(B=B, x=x, rows=rows, pooled=pooled, h=h, key_bias=key_bias)
            if backward:
                idx = scores.v.argmax(-1) if index is None else torch.as_tensor(index, device=dev).reshape(-1).expand(B).long()
                one_hot = torch.zeros_like(scores.v)
                one_hot.scatter_(1, idx.reshape(B, 1), 1.0)
                tape.seed(scores, one_hot, B)
                tape.backward()
                if lrp:
                    from .lrp import visualbert_sweep
                    visualbert_sweep(self, one_hot)
        return self.scores

    def __call__(self, input):
        return {"scores": self.forward_backward(input, backward=False)}


class SelfAttentionGenerator:
    """VisualBERT/mmf/models/transformers/backends/ExplanationGenerator.py:20-214.  ``model`` is a
    :class:`VisualBertEngine`; every method returns ``cls_per_token_score`` [B, T+V]."""

    def __init__(self, model: VisualBertEngine):
        if not isinstance(model, VisualBertEngine):
            raise MmxError("model must be a mmx_b200.VisualBertEngine")
        self.model = model.eval()

    def _cls_row(self, M: torch.Tensor) -> torch.Tensor:
        """M [B,S,S] -> row cls_index of every sample with its own entry zeroed (EG:94-97)."""
        B = M.shape[0]
        ar = torch.arange(B, device=M.device)
        out = M[ar, self.model.cls_index].clone()
        out[ar, self.model.cls_index] = 0
        return out

    def generate_ours(self, input, index=None, save_visualization=False, save_visualization_per_token=False):
        m = self.model
        m.forward_backward(input, index)
        B, S = m._shape
        R = torch.eye(S, device=m.device).repeat(B, 1, 1)                                               # EG:84-85
        for L in m.layers:                                                                             # EG:86-93: rules 5 + 6
            R, _ = rules.self_update(R, rules.avg_heads_record(L.rec, B))
        self.R = R
        return self._cls_row(R)

    def generate_raw_attn(self, input, save_visualization=False):                                       # EG:154-167
        m = self.model
        m.forward_backward(input, backward=False)
        return self._cls_row(rules.head_mean_record(m.layers[-1].rec, m._shape[0]))

    def generate_rollout(self, input, start_layer=0, save_visualization=False):                         # EG:169-185
        m = self.model
        m.forward_backward(input, backward=False)
        B = m._shape[0]
        mats = [rules.head_mean_record(L.rec, B) for L in m.layers]
        return self._cls_row(rules.compute_rollout_attention(mats, start_layer=start_layer, normalize=False))

    def generate_attn_gradcam(self, input, index=None, save_visualization=False):                       # EG:187-214
        m = self.model
        m.forward_backward(input, index)
        cam = rules.gradcam_record(m.layers[-1].rec, m._shape[0])
        return self._cls_row(rules.minmax_normalize(cam))

    def generate_transformer_att(self, input, index=None, start_layer=0, save_visualization=False,
                                 save_visualization_per_token=False):
        """Transformer attribution (EG:24-66): rule 5 on (grad, LRP relevance of A) per layer, then the non-normalising
        rollout (EG:5-18) from ``start_layer``."""
        m = self.model
        m.forward_backward(input, index, lrp=True)
        B = m._shape[0]
        mats = [rules.avg_heads_record(L.rec, B, use_cam=True) for L in m.layers]
        return self._cls_row(rules.compute_rollout_attention(mats, start_layer=start_layer, normalize=False))

    def generate_partial_lrp(self, input, index=None, save_visualization=False):
        """Partial LRP (EG:109-131): head mean of the last layer's LRP relevance, min-max normalised."""
        m = self.model
        m.forward_backward(input, index, lrp=True)
        cam = rules.head_mean_record(m.layers[-1].rec, m._shape[0], use_cam=True).contiguous()
        return self._cls_row(rules.minmax_normalize(cam))
