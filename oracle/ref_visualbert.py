"""Runs the UNMODIFIED reference VisualBERT encoder / head modules + SelfAttentionGenerator on CPU (build container only;
TEST INFRASTRUCTURE).  ``BERT_ours.py``, ``layers_ours.py`` and ``ExplanationGenerator.py`` of
VisualBERT/mmf/models/transformers/backends are loaded by file path under a private package name (importing the ``mmf``
package itself needs omegaconf and the mmf registry).  ``BertVisioLinguisticEmbeddings`` is the reference's own class too:
VisualBERT/mmf/modules/embeddings.py is executed by file path with its unrelated imports (attention / bottleneck / layers /
file_io / vocab modules, none of which the class touches) stubbed and ``transformers.modeling_bert`` aliased to the module
transformers 5.x keeps ``BertEmbeddings`` in.  So the whole model - embeddings, 12 x BertLayer with the attention hooks,
BertPredictionHeadTransform - and the generator are the reference's code."""
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


def _import_embeddings():
    """``BertVisioLinguisticEmbeddings`` (VisualBERT/mmf/modules/embeddings.py:305-451) from the reference file itself."""
    rs._ensure_path()
    full = f"{_PKG}.mmf_embeddings"
    if full in sys.modules:
        return sys.modules[full].BertVisioLinguisticEmbeddings
    import transformers.models.bert.modeling_bert as mb
    sys.modules.setdefault("transformers.modeling_bert", mb)
    stubs = {"VisualBERT": [], "VisualBERT.mmf": [], "VisualBERT.mmf.modules": [], "VisualBERT.mmf.utils": [],
             "VisualBERT.mmf.modules.attention": ["AttentionLayer", "SelfAttention", "SelfGuidedAttention"],
             "VisualBERT.mmf.modules.bottleneck": ["MovieBottleneck"],
             "VisualBERT.mmf.modules.layers": ["AttnPool1d", "Identity"],
             "VisualBERT.mmf.utils.file_io": ["PathManager"], "VisualBERT.mmf.utils.vocab": ["Vocab"]}
    added = []
    for name, attrs in stubs.items():
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            for a in attrs:
                setattr(m, a, type(a, (), {}))
            sys.modules[name] = m
            added.append(name)
    try:
        spec = importlib.util.spec_from_file_location(full, os.path.join(rs.REFERENCE_ROOT, "VisualBERT", "mmf", "modules", "embeddings.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[full] = m
        spec.loader.exec_module(m)
    finally:
        for name in added:                 # the stubs must not shadow anything for later imports
            sys.modules.pop(name, None)
    return m.BertVisioLinguisticEmbeddings


def build(cfg, sd):
    lo, bo, eg = _import_ref()
    RefEmbeddings = _import_embeddings()
    from transformers import BertConfig
    hc = BertConfig(vocab_size=cfg.vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers, num_attention_heads=cfg.heads,
                    intermediate_size=cfg.intermediate, hidden_act="gelu", hidden_dropout_prob=0.0,
                    attention_probs_dropout_prob=0.0, max_position_embeddings=cfg.max_pos, type_vocab_size=cfg.type_vocab,
                    layer_norm_eps=1e-12)
    hc.visual_embedding_dim = cfg.visual_dim

    class Bert(nn.Module):
        def __init__(self):
            super().__init__()
            self.embeddings = RefEmbeddings(hc)       # the reference class (embeddings.py:305-451), "plain" strategy
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
            emb = self.bert.embeddings(inp["input_ids"], inp["token_type_ids"], visual_embeddings=inp["visual_embeddings"],
                                       visual_embeddings_type=inp["visual_embeddings_type"])       # visual_bert.py:112-118
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
