"""DETR encoder-decoder relevancy: ``Generator(model).generate_ours(...)`` with the
reference signature (DETR/modules/ExplanationGenerator.py:142-195) over the libmmx kernels.

``DetrEngine`` holds the DETR transformer + class head (DETR/models/transformer.py, DETR/models/detr.py:46-77) built
from a reference / official-checkpoint ``state_dict`` and runs on backbone features: ``img`` is ``(src, pos)`` with
``src`` [B, d_model, h, w] (after ``input_proj``) and ``pos`` the sine embedding [B or 1, d_model, h, w].  The ResNet
backbone sits below every attention layer (no gradient needed) and is outside the hot-path scope of this round.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from ._lib import lib, check, ptr, current_stream, MmxError
from .nn import fp32_gemms, Tape, Var, Weight, AttnRecord, ACT_RELU, _f32
from . import rules


def _split_mha(sd, p, device):
    """q/k/v/out projections of one attention module; accepts split keys or the packed ``in_proj_weight`` of the
    official checkpoints (split by the reference in DETR/modules/layers.py:711-726)."""
    if p + "q_proj.weight" in sd:
        q, k, v = ((sd[p + n + ".weight"], sd[p + n + ".bias"]) for n in ("q_proj", "k_proj", "v_proj"))
    else:
        w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
        d = w.shape[1]
        q, k, v = ((w[i * d:(i + 1) * d], b[i * d:(i + 1) * d]) for i in range(3))
    return dict(q=Weight(*q, device), k=Weight(*k, device), v=Weight(*v, device),
                o=Weight(sd[p + "out_proj.weight"], sd[p + "out_proj.bias"], device), rec=AttnRecord())


class _Layer:
    pass


class DetrEngine:
    def __init__(self, state_dict: Dict[str, torch.Tensor], nhead: int = 8, device=None):
        if not torch.cuda.is_available():
            raise MmxError("mmx_b200 needs a CUDA (sm_100) device; there is no CPU fallback")
        self.device = d = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        sd = {(k if k.startswith(("transformer.", "class_embed", "query_embed")) else "transformer." + k): v
              for k, v in state_dict.items()}
        self.nhead = nhead
        n_enc = len({k.split(".")[3] for k in sd if k.startswith("transformer.encoder.layers.")})
        n_dec = len({k.split(".")[3] for k in sd if k.startswith("transformer.decoder.layers.")})
        ln = lambda p: (_f32(sd[p + ".weight"], d), _f32(sd[p + ".bias"], d))
        self.encoder, self.decoder = [], []
        for i in range(n_enc):
            p = f"transformer.encoder.layers.{i}."
            L = _Layer()
            L.self_attn = _split_mha(sd, p + "self_attn.", d)
            L.l1 = Weight(sd[p + "linear1.weight"], sd[p + "linear1.bias"], d)
            L.l2 = Weight(sd[p + "linear2.weight"], sd[p + "linear2.bias"], d)
            L.n1, L.n2 = ln(p + "norm1"), ln(p + "norm2")
            self.encoder.append(L)
        for i in range(n_dec):
            p = f"transformer.decoder.layers.{i}."
            L = _Layer()
            L.self_attn = _split_mha(sd, p + "self_attn.", d)
            L.multihead_attn = _split_mha(sd, p + "multihead_attn.", d)
            L.l1 = Weight(sd[p + "linear1.weight"], sd[p + "linear1.bias"], d)
            L.l2 = Weight(sd[p + "linear2.weight"], sd[p + "linear2.bias"], d)
            L.n1, L.n2, L.n3 = ln(p + "norm1"), ln(p + "norm2"), ln(p + "norm3")
            self.decoder.append(L)
        self.dec_norm = ln("transformer.decoder.norm")
        self.query_embed = _f32(sd["query_embed.weight"], d)
        self.class_embed = Weight(sd["class_embed.weight"], sd["class_embed.bias"], d)
        self.d_model = self.query_embed.shape[1]
        self.queries = self.query_embed.shape[0]
        self.pred_logits: Optional[torch.Tensor] = None

    def eval(self):
        return self

    def zero_grad(self):
        return None

    def _mha(self, tape: Tape, m, query: Var, key: Var, value: Var, B, T, S) -> Var:
        """DETR/modules/layers.py:728-768: separate projections, q scaled before q k^T, masks ignored."""
        H = self.nhead
        q, k, v = tape.linear(query, m["q"]), tape.linear(key, m["k"]), tape.linear(value, m["v"])
        o = tape.attention(q, k, v, B, H, T, S, float(self.d_model // H) ** -0.5, 0, None, m["rec"])
        m["rec"].saved.update(Xq=query, Xk=key, Xv=value)
        return tape.linear(o, m["o"])

    def forward_backward(self, *args, lrp: bool = False, **kwargs):
        """See ``_forward_backward``.  With ``lrp=True`` every GEMM of the call (forward, dgrad and the relprop sweep) runs
        on the fp32 FFMA backend (``nn.fp32_gemms``: the sweep amplifies tensor-core split rounding)."""
        if lrp:
            with fp32_gemms():
                return self._forward_backward(*args, lrp=True, **kwargs)
        return self._forward_backward(*args, lrp=False, **kwargs)

    def _forward_backward(self, src: torch.Tensor, pos: torch.Tensor, target_index, index=None, backward: bool = True,
                         lrp: bool = False):
        """Forward staging every A, one-hot on pred_logits[b, target_b, class_b], backward staging every dA
        (``backward=False``: forward only, for the raw-attention / rollout baselines); ``lrp=True`` adds the relprop
        sweep from the same one-hot (DETR/models/detr.py:79-92), staging the relevance of every A."""
        dev = self.device
        with torch.cuda.device(dev):
            B, d, h, w = src.shape
            S, Q = h * w, self.queries
            tape = Tape(dev)
            x = Var(_f32(src, dev).flatten(2).permute(0, 2, 1).reshape(B * S, d).contiguous())
            pe = _f32(pos, dev).expand(B, d, h, w).flatten(2).permute(0, 2, 1).reshape(B * S, d).contiguous()
            qe = self.query_embed.repeat(B, 1)                                    # [B*Q, d]
            for L in self.encoder:                                                # transformer.py:230-254
                qk = tape.add_const(x, pe)
                att = self._mha(tape, L.self_attn, qk, qk, x, B, S, S)
                x1 = tape.layernorm(tape.add(x, att), *L.n1, 1e-5)
                hdn = tape.linear(x1, L.l1, ACT_RELU)
                ff = tape.linear(hdn, L.l2)
                L.saved = dict(src=x, webmd=qk, drop=att, x1=x1, h=hdn, ff=ff)
                x = tape.layernorm(tape.add(x1, ff), *L.n2, 1e-5)
            memory = x
            mem_pe = tape.add_const(memory, pe)
            t = Var(torch.zeros(B * Q, d, device=dev))
            for L in self.decoder:                                                # transformer.py:372-408
                qk = tape.add_const(t, qe)
                drop1 = self._mha(tape, L.self_attn, qk, qk, t, B, Q, Q)
                x1 = tape.layernorm(tape.add(t, drop1), *L.n1, 1e-5)
                tq = tape.add_const(x1, qe)
                drop2 = self._mha(tape, L.multihead_attn, tq, mem_pe, memory, B, Q, S)
                x2 = tape.layernorm(tape.add(x1, drop2), *L.n2, 1e-5)
                hdn = tape.linear(x2, L.l1, ACT_RELU)
                ff = tape.linear(hdn, L.l2)
                L.saved = dict(tgt=t, memory=memory, webmd=qk, drop1=drop1, x1=x1, drop2=drop2, x2=x2, h=hdn, ff=ff)
                t = tape.layernorm(tape.add(x2, ff), *L.n3, 1e-5)
                L.saved["out"] = t
            hs = tape.layernorm(t, *self.dec_norm, 1e-5)
            logits = tape.linear(hs, self.class_embed)                             # [B*Q, C+1]
            C1 = logits.cols
            self.pred_logits = logits.v.view(B, Q, C1)
            self._shape = (B, S, Q)
            self.saved = dict(B=B, memory=memory, hs=hs, logits=logits)
            if not backward:
                return {"pred_logits": self.pred_logits}
            tq_idx = torch.as_tensor(target_index, device=dev).reshape(B).long()
            rows = torch.arange(B, device=dev) * Q + tq_idx
            if index is None:                                                      # ExplanationGenerator.py:151-152
                cls = logits.v[rows, :-1].argmax(-1)
            else:
                cls = torch.as_tensor(index, device=dev).reshape(B).long()
            one_hot = torch.zeros_like(logits.v)
            one_hot.view(-1).index_fill_(0, rows * C1 + cls, 1.0)          # one_hot[rows, cls] = 1 without a host scalar (graph-capturable)
            tape.seed(logits, one_hot, B)
            tape.backward()
            if lrp:
                from .lrp import detr_sweep
                detr_sweep(self, one_hot)
        return {"pred_logits": self.pred_logits}


class Generator:
    """DETR/modules/ExplanationGenerator.py:56-195 (``generate_ours`` and its three handlers)."""

    def __init__(self, model: DetrEngine):
        if not isinstance(model, DetrEngine):
            raise MmxError("model must be a mmx_b200.DetrEngine")
        self.model = model
        self.model.eval()

    def handle_self_attention_image(self, blocks):                                 # :110-118
        for blk in blocks:
            cam = rules.avg_heads_record(blk.self_attn["rec"], self.B, self.use_lrp)
            self.R_i_i, _ = rules.self_update(self.R_i_i, cam)

    def handle_co_attn_self_query(self, block):                                    # :120-129
        cam = rules.avg_heads_record(block.self_attn["rec"], self.B, self.use_lrp)
        self.R_q_q, self.R_q_i = rules.self_update(self.R_q_q, cam, self.R_q_i)

    def handle_co_attn_query(self, block):                                         # :131-140
        cam_q_i = rules.avg_heads_record(block.multihead_attn["rec"], self.B, self.use_lrp)
        add, _, md = rules.mm_update_batched(self.R_q_q, self.R_i_i, None, cam_q_i,
                                             apply_normalization=self.normalize_self_attention,
                                             apply_self_in_rule_10=self.apply_self_in_rule_10, nan_to_zero=True)
        self._min_diag.append(md)
        self.R_q_i = rules.add(self.R_q_i, add)

    def generate_ours(self, img, target_index, index=None, use_lrp=True, normalize_self_attention=True,
                      apply_self_in_rule_10=True):
        self.use_lrp = bool(use_lrp)
        self.normalize_self_attention = normalize_self_attention
        self.apply_self_in_rule_10 = apply_self_in_rule_10
        src, pos = img
        m = self.model
        m.forward_backward(src, pos, target_index, index, lrp=self.use_lrp)
        B, S, Q = m._shape
        self.B = B
        dev = m.device
        self._min_diag: List[torch.Tensor] = []
        self.R_i_i = torch.eye(S, device=dev).repeat(B, 1, 1)                       # :176-180
        self.R_q_q = torch.eye(Q, device=dev).repeat(B, 1, 1)
        self.R_q_i = torch.zeros(B, Q, S, device=dev)
        self.handle_self_attention_image(m.encoder)
        for blk in m.decoder:
            self.handle_co_attn_self_query(blk)
            self.handle_co_attn_query(blk)
        self.min_diag = torch.stack(self._min_diag).min() if (normalize_self_attention and self._min_diag) else None
        if not getattr(self, "_defer_assert", False):
            self.check_min_diag()
        return self._pick(target_index)

    def check_min_diag(self):
        """handle_residual's ``assert self_attention[diag].min() >= 0`` (:50) over every cross-attention step of the last
        call (one host read; deferred to after the replay when the call runs as a CUDA graph)."""
        if getattr(self, "min_diag", None) is not None:
            assert self.min_diag.item() >= 0

    def capture(self, img, target_index, index=None, use_lrp=True, normalize_self_attention=True, apply_self_in_rule_10=True):
        """``generate_ours`` for these shapes / flags as a CUDA graph (``mmx_b200.graphs.Graphed``): returns a callable
        ``g(src, pos, target_index)`` that replays the captured kernels on new inputs and yields the (static) result
        tensor; ``g.check()`` runs the deferred ``diag(R - I) >= 0`` assert.  ``target_index`` must be a tensor."""
        from .graphs import Graphed
        src, pos = img
        dev = self.model.device
        tq = torch.as_tensor(target_index).reshape(-1).to(dev)
        idx = None if index is None else torch.as_tensor(index).reshape(-1).to(dev)

        def fn(s, p, t):
            return self.generate_ours((s, p), t, idx, use_lrp, normalize_self_attention, apply_self_in_rule_10)
        self._defer_assert = True
        try:
            return Graphed(fn, (src, pos, tq), dev, deferred_checks=[self.check_min_diag])
        finally:
            self._defer_assert = False

    def _pick(self, target_index):
        """``aggregated[:, target_index, :].unsqueeze_(0)`` (:192-194): [1,1,1,S] for a tensor index, [1,1,S] for an int;
        [B,S] for a batch."""
        B, S, Q = self.model._shape
        dev = self.model.device
        tq = torch.as_tensor(target_index, device=dev).reshape(B).long()
        picked = self.R_q_i[torch.arange(B, device=dev), tq].contiguous()
        if B == 1:
            return picked.view(1, 1, 1, S) if isinstance(target_index, torch.Tensor) else picked.view(1, 1, S)
        return picked

    # ---- baselines sharing the API; the LRP ones need relprop and are not implemented
    def generate_raw_attn(self, img, target_index):
        """Head-mean of the last decoder layer's cross-attention (DETR/modules/ExplanationGenerator.py:224-236)."""
        src, pos = img
        m = self.model
        m.forward_backward(src, pos, target_index, backward=False)
        self.B = m._shape[0]
        self.R_q_i = rules.head_mean_record(m.decoder[-1].multihead_attn["rec"], self.B)
        return self._pick(target_index)

    def generate_rollout(self, img, target_index):
        """Attention rollout (DETR/modules/ExplanationGenerator.py:238-273): rollout of the encoder and decoder
        self-attentions, combined through the last cross-attention as R_qq^T (A_qi R_ii)."""
        src, pos = img
        m = self.model
        m.forward_backward(src, pos, target_index, backward=False)
        self.B = B = m._shape[0]
        cams_image = [rules.head_mean_record(blk.self_attn["rec"], B) for blk in m.encoder]
        cams_queries = [rules.head_mean_record(blk.self_attn["rec"], B) for blk in m.decoder]
        self.R_i_i = rules.compute_rollout_attention(cams_image)
        self.R_q_q = rules.compute_rollout_attention(cams_queries)
        cam_q_i = rules.head_mean_record(m.decoder[-1].multihead_attn["rec"], B)
        self.R_q_i, _, _ = rules.mm_update_batched(self.R_q_q, self.R_i_i, None, cam_q_i, apply_normalization=False,
                                                   apply_self_in_rule_10=True)
        return self._pick(target_index)

    def generate_attn_gradcam(self, img, target_index, index=None):
        """attn-GradCAM of the last cross-attention (DETR/modules/ExplanationGenerator.py:275-305)."""
        src, pos = img
        m = self.model
        m.forward_backward(src, pos, target_index, index)
        self.B = m._shape[0]
        self.R_q_i = rules.gradcam_record(m.decoder[-1].multihead_attn["rec"], self.B)
        return self._pick(target_index)

    def generate_partial_lrp(self, img, target_index, index=None):
        """Partial LRP (DETR/modules/ExplanationGenerator.py:197-224): head mean of the LRP relevance of the last decoder
        cross-attention, min-max normalised over the whole [Q, S] map."""
        src, pos = img
        m = self.model
        m.forward_backward(src, pos, target_index, index, lrp=True)
        self.B = B = m._shape[0]
        cam = rules.head_mean_record(m.decoder[-1].multihead_attn["rec"], B, use_cam=True).contiguous()
        self.R_q_i = rules.minmax_normalize(cam)
        return self._pick(target_index)

    def generate_transformer_att(self, img, target_index, index=None):
        """Transformer attribution (DETR/modules/ExplanationGenerator.py:64-108): rule 5 on the last decoder
        cross-attention with the LRP relevance in place of the probabilities."""
        src, pos = img
        m = self.model
        m.forward_backward(src, pos, target_index, index, lrp=True)
        self.B = B = m._shape[0]
        self.R_q_i = rules.avg_heads_record(m.decoder[-1].multihead_attn["rec"], B, use_cam=True).contiguous()
        return self._pick(target_index)


class GeneratorAlbationNoAgg(Generator):
    """DETR/modules/ExplanationGenerator.py:306-403 (sic): the ablation without aggregation - every update REPLACES the
    relevancy instead of adding to it (``R = cam R`` for the self-attention rules, ``R_q_i = rule 10`` for the
    cross-attention)."""

    def handle_self_attention_image(self, blocks):                                 # :314-322
        for blk in blocks:
            self.R_i_i = rules.bmm(rules.avg_heads_record(blk.self_attn["rec"], self.B, self.use_lrp), self.R_i_i)

    def handle_co_attn_self_query(self, block):                                    # :324-333
        cam = rules.avg_heads_record(block.self_attn["rec"], self.B, self.use_lrp)
        self.R_q_q, self.R_q_i = rules.bmm(cam, self.R_q_q), rules.bmm(cam, self.R_q_i)

    def handle_co_attn_query(self, block):                                         # :335-344
        cam_q_i = rules.avg_heads_record(block.multihead_attn["rec"], self.B, self.use_lrp)
        self.R_q_i, _, md = rules.mm_update_batched(self.R_q_q, self.R_i_i, None, cam_q_i,
                                                    apply_normalization=self.normalize_self_attention,
                                                    apply_self_in_rule_10=self.apply_self_in_rule_10, nan_to_zero=True)
        self._min_diag.append(md)

    def generate_ours_abl(self, img, target_index, index=None, use_lrp=False, normalize_self_attention=False,
                          apply_self_in_rule_10=True):                             # :346-403
        return Generator.generate_ours(self, img, target_index, index, use_lrp, normalize_self_attention, apply_self_in_rule_10)

    def generate_ours(self, *a, **k):
        raise AttributeError("GeneratorAlbationNoAgg has generate_ours_abl only (DETR/modules/ExplanationGenerator.py:346)")


class MaskGenerator:
    """The relevance -> segmentation-mask step of DETR/mask_generator.py:39-125 (``get_panoptic``): one relevance map per
    kept query through the ``method`` switch (:93-113), then min-max / Otsu (:115-121).  The reference loops over the
    kept queries with one forward + backward and one CPU cv2 call each; here the queries of an image form ONE batch and
    the masks come from one kernel launch.  The visualisation / COCO panoptic bookkeeping around it is out of scope."""

    METHODS = ("ours_with_lrp", "ours_no_lrp", "ablation_no_self_in_10", "ablation_no_aggregation", "ours_no_lrp_no_norm",
               "transformer_att", "partial_lrp", "raw_attn", "attn_gradcam", "rollout")

    def __init__(self, model: DetrEngine):
        self.gen = Generator(model)
        self.abl = GeneratorAlbationNoAgg(model)
        self.model = model

    def get_masks(self, img, queries, method: str = "ours_no_lrp"):
        """img = (src [1,C,h,w], pos [1,C,h,w]); queries: the kept query indices [K].  Returns (masks [K,h,w] in {0,255},
        thresholds [K], cams [K,h*w])."""
        src, pos = img
        q = torch.as_tensor(queries).reshape(-1).long()
        K = q.numel()
        if K == 0:
            raise MmxError("no queries to segment")
        batch = (src.expand(K, *src.shape[1:]).contiguous(), pos.expand(K, *pos.shape[1:]).contiguous())
        if method == "ours_no_lrp":
            cam = self.gen.generate_ours(batch, q, use_lrp=False)
        elif method == "ablation_no_self_in_10":
            cam = self.gen.generate_ours(batch, q, use_lrp=False, apply_self_in_rule_10=False)
        elif method == "ablation_no_aggregation":
            cam = self.abl.generate_ours_abl(batch, q, use_lrp=False, normalize_self_attention=False)
        elif method == "ours_no_lrp_no_norm":
            cam = self.gen.generate_ours(batch, q, use_lrp=False, normalize_self_attention=False)
        elif method == "raw_attn":
            cam = self.gen.generate_raw_attn(batch, q)
        elif method == "attn_gradcam":
            cam = self.gen.generate_attn_gradcam(batch, q)
        elif method == "rollout":
            cam = self.gen.generate_rollout(batch, q)
        elif method == "ours_with_lrp":                                       # mask_generator.py:93-94
            cam = self.gen.generate_ours(batch, q, use_lrp=True)
        elif method == "transformer_att":
            cam = self.gen.generate_transformer_att(batch, q)
        elif method == "partial_lrp":
            cam = self.gen.generate_partial_lrp(batch, q)
        else:
            print("please provide a valid explainability method")            # mask_generator.py:111-113
            return None
        cam = cam.reshape(K, -1)
        masks, thr = rules.otsu_masks(cam)
        return masks.view(K, src.shape[-2], src.shape[-1]), thr, cam
