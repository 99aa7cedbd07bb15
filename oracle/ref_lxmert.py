"""Runs the UNMODIFIED reference LXMERT modules + GeneratorOurs on CPU (build container only; TEST INFRASTRUCTURE).
Shims: ``transformers.configuration_lxmert`` alias, no-op docstring decorators, ``.cuda()`` identity; the
``PreTrainedModel`` subclasses do not instantiate under transformers 5.x, so the plain-nn.Module parts
(LxmertEmbeddings / LxmertEncoder / LxmertPooler / LxmertVisualAnswerHead) are wrapped with the same attribute names."""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

import torch

from . import ref_shims as rs


def _import_ref():
    rs._ensure_path()
    import transformers
    if "transformers.configuration_lxmert" not in sys.modules:
        sys.modules["transformers.configuration_lxmert"] = importlib.import_module("transformers.models.lxmert.configuration_lxmert")
    fu = importlib.import_module("transformers.file_utils") if importlib.util.find_spec("transformers.file_utils") else types.ModuleType("transformers.file_utils")
    ident = lambda *a, **k: (lambda f: f)
    for n in ("add_code_sample_docstrings", "add_start_docstrings", "add_start_docstrings_to_callable", "replace_return_docstrings",
              "add_start_docstrings_to_model_forward"):
        setattr(fu, n, ident)
    if not hasattr(fu, "ModelOutput"):
        from transformers.utils import ModelOutput
        fu.ModelOutput = ModelOutput
    sys.modules["transformers.file_utils"] = fu
    src = os.path.join(rs.REFERENCE_ROOT, "lxmert", "lxmert")
    if src not in sys.path:
        sys.path.insert(0, src)
    lrp = importlib.import_module("src.lxmert_lrp")
    eg = importlib.import_module("src.ExplanationGenerator")
    return lrp, eg


def build(cfg, sd):
    lrp, eg = _import_ref()
    from transformers.models.lxmert.configuration_lxmert import LxmertConfig as HFConfig
    hc = HFConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_attention_heads=cfg.heads, intermediate_size=cfg.intermediate,
                  l_layers=cfg.l_layers, x_layers=cfg.x_layers, r_layers=cfg.r_layers, max_position_embeddings=cfg.max_pos,
                  visual_feat_dim=cfg.feat_dim, visual_pos_dim=cfg.pos_dim, hidden_dropout_prob=0.0,
                  attention_probs_dropout_prob=0.0, num_qa_labels=cfg.num_labels)
    hc.output_attentions = False

    class Lx(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.embeddings = lrp.LxmertEmbeddings(hc)
            self.encoder = lrp.LxmertEncoder(hc)
            self.pooler = lrp.LxmertPooler(hc)

    class QA(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lxmert = Lx()
            self.answer_head = lrp.LxmertVisualAnswerHead(hc, cfg.num_labels)
            self.device = torch.device("cpu")

        def forward(self, ids, feats, boxes):
            emb = self.lxmert.embeddings(ids, torch.zeros_like(ids))
            vis_out, lang_out, _ = self.lxmert.encoder(emb, None, feats, boxes, None, output_attentions=False)
            self.vis_shape = vis_out[0][-1].shape                      # lxmert_lrp.py:1677
            pooled = self.lxmert.pooler(lang_out[0][-1])
            return self.answer_head(pooled)

        def relprop(self, cam, **kwargs):
            # LxmertForQuestionAnswering.relprop (lxmert_lrp.py:1689-1693) + LxmertModel.relprop (:1253-1257): the
            # PreTrainedModel subclasses cannot be instantiated here, their five relprop lines are restated
            cam_lang = self.answer_head.relprop(cam, **kwargs)
            cam_vis = torch.zeros(self.vis_shape).to(cam_lang.device)
            cam_lang = self.lxmert.pooler.relprop(cam_lang, **kwargs)
            return self.lxmert.encoder.relprop((cam_lang, cam_vis), **kwargs)

    m = QA().eval()
    missing, unexpected = m.load_state_dict(sd, strict=False)
    missing = [k for k in missing if "position_ids" not in k]
    assert not missing and not unexpected, (missing, unexpected)
    return m, eg


def generate_ours(cfg, sd, ids, feats, boxes, use_lrp=False, **kw):
    m, eg = build(cfg, sd)
    Rtt, Rti = [], []
    with rs.cuda_is_identity():
        for b in range(ids.shape[0]):
            class Usage:
                model = m
                text_len = ids.shape[1]
                image_boxes_len = feats.shape[1]

                def forward(self, item):
                    out = types.SimpleNamespace()
                    out.question_answering_score = m(ids[b:b + 1], feats[b:b + 1], boxes[b:b + 1])
                    return out
            gen = eg.GeneratorOurs(Usage())
            a, c = gen.generate_ours(None, use_lrp=use_lrp, **kw)
            Rtt.append(a.detach().clone()); Rti.append(c.detach().clone())
    return torch.stack(Rtt), torch.stack(Rti)


def generate_baseline(cfg, sd, ids, feats, boxes, method):
    """method in {"raw_attn", "rollout", "attn_gradcam"}: GeneratorBaselines of the reference, per sample."""
    m, eg = build(cfg, sd)
    Rtt, Rti = [], []
    with rs.cuda_is_identity():
        for b in range(ids.shape[0]):
            class Usage:
                model = m
                text_len = ids.shape[1]
                image_boxes_len = feats.shape[1]

                def forward(self, item):
                    out = types.SimpleNamespace()
                    out.question_answering_score = m(ids[b:b + 1], feats[b:b + 1], boxes[b:b + 1])
                    return out
            gen = eg.GeneratorBaselines(Usage())
            a, c = getattr(gen, "generate_" + method)(None)
            Rtt.append(a.detach().clone()); Rti.append(c.detach().clone())
    return torch.stack(Rtt), torch.stack(Rti)


def generate_ours_no_agg(cfg, sd, ids, feats, boxes, **kw):
    """The reference's GeneratorOursAblationNoAggregation.generate_ours_no_agg (ExplanationGenerator.py:215-365), per sample."""
    m, eg = build(cfg, sd)
    Rtt, Rti = [], []
    with rs.cuda_is_identity():
        for b in range(ids.shape[0]):
            class Usage:
                model = m
                text_len = ids.shape[1]
                image_boxes_len = feats.shape[1]

                def forward(self, item):
                    out = types.SimpleNamespace()
                    out.question_answering_score = m(ids[b:b + 1], feats[b:b + 1], boxes[b:b + 1])
                    return out
            gen = eg.GeneratorOursAblationNoAggregation(Usage())
            a, c = gen.generate_ours_no_agg(None, use_lrp=False, **kw)
            Rtt.append(a.detach().clone()); Rti.append(c.detach().clone())
    return torch.stack(Rtt), torch.stack(Rti)


def forward_fn(cfg, sd):
    """(input_ids, visual_feats, visual_pos) -> answer scores through the unmodified reference LXMERT modules (what
    ``ModelPert.lxmert_vqa(...)`` returns as ``question_answering_score``), for oracle/ref_perturbation.py."""
    m, _ = build(cfg, sd)

    def fn(ids, feats, boxes):
        with torch.enable_grad():                       # the attention hooks of lxmert_lrp.py register on tensors that require grad
            return m(ids, feats, boxes).detach()
    return fn
