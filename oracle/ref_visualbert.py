"""Runs the UNMODIFIED reference VisualBERT encoder / head modules + SelfAttentionGenerator on CPU (build container only;
TEST INFRASTRUCTURE).  ``BERT_ours.py``, ``layers_ours.py`` and ``ExplanationGenerator.py`` of
VisualBERT/mmf/models/transformers/backends are loaded by file path under a private package name (importing the ``mmf``
package itself needs omegaconf and the mmf registry).  The mmf embeddings module has the same problem, so
``BertVisioLinguisticEmbeddings`` (VisualBERT/mmf/modules/embeddings.py:305-451) is restated here with nn.Embedding /
nn.Linear / nn.LayerNorm under the reference's parameter names; everything above it (12 x BertLayer with the attention
hooks, BertPredictionHeadTransform, the generator) is the reference's own code."""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch
from torch import nn

from . import ref_shims as rs

_PKG = "mmx_ref_visualbert_backends"


def _import_ref():
    rs._ensure_path()
    base = os.path.join(rs.REFERENCE_ROOT, "VisualBERT", "mmf", "models", "transformers", "backends")
    if _PKG not in sys.modules:
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [base]
        sys.modules[_PKG] = pkg
    mods = []
    for name in ("layers_ours", "BERT_ours", "ExplanationGenerator"):
        full = f"{_PKG}.{name}"
        if full not in sys.modules:
            spec = importlib.util.spec_from_file_location(full, os.path.join(base, name + ".py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[full] = m
            spec.loader.exec_module(m)
        mods.append(sys.modules[full])
    return mods


def build(cfg, sd):
    lo, bo, eg = _import_ref()
    from transformers import BertConfig
    hc = BertConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                    intermediate_size=cfg.intermediate, hidden_act="gelu", hidden_dropout_prob=0.0,
                    attention_probs_dropout_prob=0.0, max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab,
                    layer_norm_eps=1e-12)

    class Embeddings(nn.Module):            # embeddings.py:305-451, "plain" strategy
        def __init__(self):
            super().__init__()
            self.word_embeddings = nn.Embedding(cfg.vocab, cfg.hidden)
            self.position_embeddings = nn.Embedding(cfg.max_pos, cfg.hidden)
            self.token_type_embeddings = nn.Embedding(cfg.type_vocab, cfg.hidden)
            self.LayerNorm = nn.LayerNorm(cfg.hidden, eps=1e-12)
            self.token_type_embeddings_visual = nn.Embedding(cfg.type_vocab, cfg.hidden)
            self.position_embeddings_visual = nn.Embedding(cfg.max_pos, cfg.hidden)
            self.projection = nn.Linear(cfg.visual_dim, cfg.hidden)

        def forward(self, input_ids, token_type_ids, visual_embeddings, visual_embeddings_type):
            pos = torch.arange(input_ids.size(1)).unsqueeze(0).expand_as(input_ids)
            text = self.word_embeddings(input_ids) + self.position_embeddings(pos) + self.token_type_embeddings(token_type_ids)
            v = self.projection(visual_embeddings)
            v = v + self.position_embeddings_visual(torch.zeros(v.shape[:-1], dtype=torch.long)) \
                + self.token_type_embeddings_visual(visual_embeddings_type)
            return self.LayerNorm(torch.cat((text, v), dim=1))

    class Bert(nn.Module):
        def __init__(self):
            super().__init__()
            self.embeddings = Embeddings()
            self.encoder = bo.BertEncoder(hc)

        def relprop(self, cam, **kwargs):          # VisualBERTBase.relprop, visual_bert.py:150-152
            return self.encoder.relprop(cam, **kwargs)

    class ForClassification(nn.Module):     # visual_bert.py:263-396, pooler_strategy "vqa"
        def __init__(self):
            super().__init__()
            self.bert = Bert()
            self.classifier = lo.Sequential(bo.BertPredictionHeadTransform(hc), lo.Linear(cfg.hidden, cfg.num_labels))
            self.vqa_pooler = lo.IndexSelect()

        def forward(self, inp):
            am = inp["attention_mask"]
            ext = (1.0 - am.unsqueeze(1).unsqueeze(2).to(torch.float32)) * -10000.0
            emb = self.bert.embeddings(inp["input_ids"], inp["token_type_ids"], inp["visual_embeddings"],
                                       inp["visual_embeddings_type"])
            seq = self.bert.encoder(emb, ext)[0]
            idx = inp["input_mask"].sum(1) - 2
            pooled = self.vqa_pooler(seq, 1, idx.clone().detach())
            return self.classifier(pooled).contiguous().view(-1, cfg.num_labels)

        def relprop(self, cam, **kwargs):          # VisualBERTForClassification.relprop, visual_bert.py:398-403
            for m_ in reversed(self.classifier._modules.values()):
                cam = m_.relprop(cam, **kwargs)
            cam = self.vqa_pooler.relprop(cam, **kwargs)
            return self.bert.relprop(cam, **kwargs)

    class Wrapper(nn.Module):               # the mmf ``VisualBERT`` model: model(input)['scores'], .model.bert.encoder.layer
        def __init__(self):
            super().__init__()
            self.model = ForClassification()

        def forward(self, inp):
            return {"scores": self.model(inp)}

        def relprop(self, cam, **kwargs):          # VisualBERT.relprop, visual_bert.py:615-616
            return self.model.relprop(cam, **kwargs)

    w = Wrapper().eval()
    missing, unexpected = w.model.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return w, eg


def generate(cfg, sd, inp, method="ours", **kw):
    """method in {"ours", "raw_attn", "rollout", "attn_gradcam", "transformer_att", "partial_lrp"}: the reference
    generator, one sample at a time (the last two run the relprop sweep)."""
    w, eg = build(cfg, sd)
    outs = []
    with rs.cuda_is_identity():
        for b in range(inp["input_ids"].shape[0]):
            one = {k: v[b:b + 1] for k, v in inp.items()}
            gen = eg.SelfAttentionGenerator(w)
            outs.append(getattr(gen, "generate_" + method)(one, **kw).detach().clone()[0])
    return torch.stack(outs)
