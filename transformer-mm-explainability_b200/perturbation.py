"""Perturbation tests of the LXMERT evaluation driver: ``perturbation_image`` / ``perturbation_text`` of
lxmert/lxmert/perturbation.py:85-194 - remove the least (negative test) or most (positive test) relevant boxes / tokens
in steps and re-run the model.

The reference runs one forward per step on gathered, shorter inputs.  Here all steps of an item form ONE batch:
  * boxes are removed with a ``-inf`` additive key bias (``exp(-inf) = 0``: exactly what the gathered softmax sums; a
    step that removes every box is the softmax over an empty key set, output 0 - attention.cu), so box order and count
    never change the arithmetic of the kept keys;
  * tokens are removed by compaction: kept tokens move to positions 0..k+1 (the reference re-indexes positions through
    its ``sorted`` gather, :172-178), the tail is padding behind a ``-inf`` key bias.
Top-k selection and compaction run in ``mmx_topk_select``.  The Faster-RCNN extractor, tokenizer, COCO / VQA files and
the accuracy bookkeeping of ``ModelPert`` are outside the hot-path scope; ``answers`` (argmax per step) is what
``self.vqa_answers[...argmax()]`` consumes.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from ._lib import lib, check, ptr, current_stream, MmxError
from .lxmert import LxmertEngine
from .visualbert import VisualBertEngine

PERT_STEPS = [0, 0.25, 0.5, 0.75, 0.8, 0.85, 0.9, 0.95, 1]          # perturbation.py:42; evaluation_loop.py:96 (text test)
VISUALBERT_IMAGE_STEPS = [0, 0.5, 0.75, 0.95, 0.96, 0.97, 0.98, 0.99, 1]   # evaluation_loop.py:94 (image test of VisualBERT)


def topk_select(scores: torch.Tensor, k: Sequence[int]):
    """scores [n] (one relevance vector), k: kept count per step -> (keep [steps,n] int32, pos [steps,n] int32)."""
    if not scores.is_cuda:
        raise MmxError("mmx_b200 needs CUDA tensors (no CPU fallback)")
    steps, n = len(k), scores.numel()
    rows = scores.detach().float().reshape(1, n).expand(steps, n).contiguous()
    kk = torch.tensor(list(k), dtype=torch.int32, device=scores.device)
    keep = torch.empty(steps, n, dtype=torch.int32, device=scores.device)
    pos = torch.empty_like(keep)
    check(lib().mmx_topk_select(ptr(rows), ptr(kk), ptr(keep), ptr(pos), steps, n, current_stream()))
    return keep, pos


class LxmertPerturbation:
    """``ModelPert.perturbation_image`` / ``perturbation_text`` on a :class:`LxmertEngine`; ``item`` is
    ``(input_ids [1,T], visual_feats [1,I,F], visual_pos [1,I,4])``."""

    def __init__(self, model_usage: LxmertEngine, pert_steps: Sequence[float] = PERT_STEPS):
        if not isinstance(model_usage, LxmertEngine):
            raise MmxError("model_usage must be a mmx_b200.LxmertEngine")
        self.model_usage = model_usage
        self.pert_steps = list(pert_steps)
        self.scores: torch.Tensor | None = None      # [steps, num_labels] of the last call

    def perturbation_image(self, item, cam_image: torch.Tensor, cam_text: torch.Tensor = None, is_positive_pert: bool = False):
        """perturbation.py:85-133.  Returns the argmax answer index per step [steps]."""
        ids, feats, boxes = item
        m = self.model_usage
        cam = cam_image.reshape(-1) * (-1 if is_positive_pert else 1)
        I = cam.numel()
        k = [int((1 - step) * I) for step in self.pert_steps]                       # :112
        keep, _ = topk_select(cam, k)
        n = len(k)
        bias = torch.zeros(n, I, device=cam.device).masked_fill_(keep == 0, float("-inf"))
        rep = lambda t: t.to(m.device).expand(n, *t.shape[1:]).contiguous()
        self.scores = m.forward_backward(rep(ids), rep(feats), rep(boxes), backward=False, vis_key_bias=bias)
        return self.scores.argmax(-1)

    def perturbation_text(self, item, cam_image: torch.Tensor, cam_text: torch.Tensor, is_positive_pert: bool = False):
        """perturbation.py:135-194: [CLS] and [SEP] always stay, the kept tokens keep their order."""
        ids, feats, boxes = item
        m = self.model_usage
        cam = cam_text.reshape(-1) * (-1 if is_positive_pert else 1)
        T = cam.numel()
        pure = cam[1:-1].contiguous()                                               # :163
        text_len = pure.numel()
        k = [int((1 - step) * text_len) for step in self.pert_steps]                # :166
        keep, pos = topk_select(pure, k)
        n = len(k)
        dev = cam.device
        kk = torch.tensor(k, device=dev)
        src_ids = ids.to(dev).reshape(-1).long()
        new_ids = torch.zeros(n, T, dtype=torch.long, device=dev)
        new_ids[:, 0] = src_ids[0]                                                  # [CLS]
        rows = torch.arange(n, device=dev).unsqueeze(1).expand(n, text_len)
        kept = keep.bool()
        new_ids[rows[kept], (pos[kept] + 1).long()] = src_ids[1:-1].unsqueeze(0).expand(n, text_len)[kept]
        new_ids[torch.arange(n, device=dev), kk + 1] = src_ids[-1]                  # [SEP] right after the kept tokens
        bias = torch.zeros(n, T, device=dev).masked_fill_(torch.arange(T, device=dev).unsqueeze(0) > (kk + 1).unsqueeze(1),
                                                          float("-inf"))
        rep = lambda t: t.to(m.device).expand(n, *t.shape[1:]).contiguous()
        self.scores = m.forward_backward(new_ids, rep(feats), rep(boxes), backward=False, lang_key_bias=bias)
        return self.scores.argmax(-1)


class VisualBertPerturbation:
    """The perturbation branch of VisualBERT/mmf/trainers/core/evaluation_loop.py:99-169 on a :class:`VisualBertEngine`:
    ``input`` is one sample's fields (see visualbert.py), ``method_cam`` the generator's ``cls_per_token_score`` [1, T+V].
    Image test: the visual tokens are ranked (:107-120); text test: tokens 1 .. cls_index-1 are ranked and token 0, the
    pooled token ``cls_index`` and the final [SEP] always stay (:133-150).  All steps run as one batch."""

    def __init__(self, model: VisualBertEngine, pert_steps: Sequence[float] | None = None):
        if not isinstance(model, VisualBertEngine):
            raise MmxError("model must be a mmx_b200.VisualBertEngine")
        self.model = model
        self.pert_steps = None if pert_steps is None else list(pert_steps)      # None: the reference's per-modality lists
        self.scores: torch.Tensor | None = None

    def steps_for(self, modality: str):
        """evaluation_loop.py:93-96: the image test removes boxes on a finer grid near 100 % than the text test."""
        if self.pert_steps is not None:
            return self.pert_steps
        return list(VISUALBERT_IMAGE_STEPS if modality == "image" else PERT_STEPS)

    def _base(self, input, n):
        dev = self.model.device
        inp = {k: v.to(dev).expand(n, *v.shape[1:]).contiguous() for k, v in input.items() if k != "key_bias"}
        T = int(input["input_mask"].sum())
        S = T + input["visual_embeddings"].shape[1]
        am = input.get("attention_mask")
        bias = torch.zeros(n, S, device=dev) if am is None else ((1.0 - am.to(dev).float()) * -10000.0).expand(n, S).contiguous()
        return inp, bias, T

    def perturbation_image(self, input, method_cam: torch.Tensor, is_positive_pert: bool = False):
        cam = method_cam.reshape(-1) * (-1 if is_positive_pert else 1)
        steps = self.steps_for("image")
        n = len(steps)
        inp, bias, T = self._base(input, n)
        bbox_scores = cam[T:].contiguous()                                           # :107
        k = [int((1 - step) * bbox_scores.numel()) for step in steps]               # :112
        keep, _ = topk_select(bbox_scores, k)
        bias[:, T:] = bias[:, T:].masked_fill(keep == 0, float("-inf"))
        inp["key_bias"] = bias
        self.scores = self.model.forward_backward(inp, backward=False)
        return self.scores.argmax(-1)

    def perturbation_text(self, input, method_cam: torch.Tensor, is_positive_pert: bool = False):
        cam = method_cam.reshape(-1) * (-1 if is_positive_pert else 1)
        steps = self.steps_for("text")
        n = len(steps)
        inp, bias, T = self._base(input, n)
        dev = self.model.device
        cls_index = T - 2                                                            # :130
        text_scores = cam[1:cls_index].contiguous()                                  # :134
        text_len = text_scores.numel()
        k = [int((1 - step) * text_len) for step in steps]
        keep, pos = topk_select(text_scores, k)
        kk = torch.tensor(k, device=dev)
        ar = torch.arange(n, device=dev)
        kept = keep.bool()
        rows = ar.unsqueeze(1).expand(n, text_len)
        for name in ("input_ids", "token_type_ids"):
            if inp.get(name) is None:
                continue
            src = inp[name][0].clone()
            new = torch.zeros_like(inp[name])
            new[:, 0] = src[0]
            new[rows[kept], (pos[kept] + 1).long()] = src[1:cls_index].unsqueeze(0).expand(n, text_len)[kept]
            new[ar, kk + 1] = src[cls_index]                                        # the pooled token, then [SEP] (:146-148)
            new[ar, kk + 2] = src[cls_index + 1]
            inp[name] = new
        live = torch.arange(T, device=dev).unsqueeze(0) < (kk + 3).unsqueeze(1)
        inp["input_mask"] = live.long()
        bias[:, :T] = bias[:, :T].masked_fill(~live, float("-inf"))
        inp["key_bias"] = bias
        self.scores = self.model.forward_backward(inp, backward=False)
        return self.scores.argmax(-1)
