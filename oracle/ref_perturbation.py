"""Runs the UNMODIFIED perturbation loops of the reference's LXMERT evaluation driver on CPU (build container only; TEST
INFRASTRUCTURE).  ``ModelPert.perturbation_image`` / ``perturbation_text`` (lxmert/lxmert/perturbation.py:85-194) are
methods of a class whose constructor downloads Faster-RCNN, the tokenizer, the VQA answer list and needs COCO files; the
loops themselves only touch ``self.<attribute>``.  Their source is taken straight from the reference file (between
``def perturbation_image`` and ``def main``), exec-ed as plain functions, and driven with a duck-typed ``self`` whose
feature extractor / tokenizer return the given tensors and whose ``lxmert_vqa`` is a callable on
(input_ids, visual_feats, visual_pos).  What comes back is the answer-score vector of every step - the quantity the
reference turns into an accuracy with the VQA labels (``self.pert_acc``)."""
from __future__ import annotations

import contextlib
import os
import textwrap
import types

import torch

from . import ref_shims as rs


@contextlib.contextmanager
def _to_cuda_is_identity():
    orig = torch.Tensor.to

    def to(self, *a, **k):
        if a and isinstance(a[0], str) and a[0].startswith("cuda"):
            return self
        return orig(self, *a, **k)
    torch.Tensor.to = to
    try:
        yield
    finally:
        torch.Tensor.to = orig


def _loops():
    rs._ensure_path()
    with open(os.path.join(rs.REFERENCE_ROOT, "lxmert", "lxmert", "perturbation.py")) as f:
        src = f.read()
    body = src[src.index("    def perturbation_image("):src.index("def main(")]
    ns = {"torch": torch}
    exec(compile(textwrap.dedent(body), "lxmert/lxmert/perturbation.py:85-194", "exec"), ns)
    return ns["perturbation_image"], ns["perturbation_text"]


def run(model_fn, ids, feats, boxes, cam_image, cam_text, modality: str, is_positive_pert: bool, num_labels: int):
    """model_fn(input_ids [1,t], visual_feats [1,i,F], visual_pos [1,i,4]) -> answer scores [1, num_labels].
    Returns (scores per step [steps, num_labels], pert_steps)."""
    pert_image, pert_text = _loops()
    calls = []

    def lxmert_vqa(input_ids, attention_mask, visual_feats, visual_pos, token_type_ids, return_dict, output_attentions):
        assert bool(attention_mask.all())                     # the tokenizer's mask of an unpadded sentence
        scores = model_fn(input_ids, visual_feats, visual_pos)
        calls.append(scores.detach().clone()[0])
        return types.SimpleNamespace(question_answering_score=scores)

    T, I = ids.shape[1], feats.shape[1]
    tok = types.SimpleNamespace(input_ids=ids, attention_mask=torch.ones_like(ids), token_type_ids=torch.zeros_like(ids))
    me = types.SimpleNamespace(
        COCO_VAL_PATH="", image_preprocess=lambda path: (None, None, None),
        frcnn=lambda images, sizes, **kw: {"normalized_boxes": boxes, "roi_features": feats},
        frcnn_cfg=types.SimpleNamespace(max_detections=I), lxmert_tokenizer=lambda sent, **kw: tok, lxmert_vqa=lxmert_vqa,
        vqa_answers=list(range(num_labels)), pert_steps=[0, 0.25, 0.5, 0.75, 0.8, 0.85, 0.9, 0.95, 1],     # perturbation.py:42
        image_boxes_len=I, text_len=T)
    me.pert_acc = [0] * len(me.pert_steps)
    item = {"img_id": "synthetic", "sent": "synthetic", "label": {}}
    with torch.no_grad(), _to_cuda_is_identity():
        (pert_image if modality == "image" else pert_text)(me, item, cam_image.clone(), cam_text.clone(), is_positive_pert)
    return torch.stack(calls), me.pert_steps
