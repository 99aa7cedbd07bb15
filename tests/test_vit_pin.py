"""CPU: pins the ViT model forward of the oracle.

The ViT the notebook imports (``baselines/ViT/ViT_new.py`` of hila-chefer/Transformer-Explainability,
Transformer_MM_explainability_ViT.ipynb:47,1212) is not vendored under the reference tree.  It wraps the published timm
``vit_base_patch16_224`` block; torchvision's ``VisionTransformer`` (``vit_b_16``) is an independent implementation of the
SAME architecture (pre-LN with eps 1e-6, packed qkv with bias, exact-erf GELU MLP, class token + learned positions, final
LayerNorm, linear head) and ships in this image (torchvision 0.26), also on the GPU box.  With the oracle's weights mapped
onto its parameter names, the logits and every block's attention probabilities (the tensors the relevancy rule hooks) must
agree - at the tiny test size and at the full ViT-B/16 size of BASELINE config 1."""
import pytest
import torch

from oracle import vit_oracle as vo
from util import rel_err

tv = pytest.importorskip("torchvision.models.vision_transformer")


def _torchvision_vit(cfg, sd):
    m = tv.VisionTransformer(image_size=cfg.image, patch_size=cfg.patch, num_layers=cfg.depth, num_heads=cfg.heads,
                             hidden_dim=cfg.dim, mlp_dim=cfg.mlp_ratio * cfg.dim, num_classes=cfg.num_classes).eval()
    t = {"class_token": sd["cls_token"], "conv_proj.weight": sd["patch_embed.proj.weight"], "conv_proj.bias": sd["patch_embed.proj.bias"],
         "encoder.pos_embedding": sd["pos_embed"], "encoder.ln.weight": sd["norm.weight"], "encoder.ln.bias": sd["norm.bias"],
         "heads.head.weight": sd["head.weight"], "heads.head.bias": sd["head.bias"]}
    names = {"norm1": "ln_1", "norm2": "ln_2", "attn.qkv": "self_attention.in_proj_", "attn.proj": "self_attention.out_proj.",
             "mlp.fc1": "mlp.0.", "mlp.fc2": "mlp.3."}
    for i in range(cfg.depth):
        src, dst = f"blocks.{i}.", f"encoder.layers.encoder_layer_{i}."
        for a, b in names.items():
            for wb in ("weight", "bias"):
                key = dst + (b + wb if b.endswith(("_", ".")) else b + "." + wb)
                t[key] = sd[src + a + "." + wb]
    missing, unexpected = m.load_state_dict(t, strict=True)
    assert not missing and not unexpected
    return m


def _attention_probs(m, cfg, images):
    """Every block's softmax(q k^T / sqrt(d)) from torchvision's own modules: the encoder is stepped block by block and the
    block's nn.MultiheadAttention is asked for its per-head weights."""
    x = m._process_input(images)
    x = torch.cat([m.class_token.expand(x.shape[0], -1, -1), x], dim=1)
    x = x + m.encoder.pos_embedding
    probs = []
    for blk in m.encoder.layers:
        h = blk.ln_1(x)
        _, w = blk.self_attention(h, h, h, need_weights=True, average_attn_weights=False)
        probs.append(w)
        x = blk(x)
    return probs


@pytest.mark.parametrize("cfg,batch", [(vo.VIT_TINY, 3), (vo.VIT_B16, 1)])
def test_oracle_forward_matches_torchvision_vit(cfg, batch):
    torch.manual_seed(0)
    sd = vo.init_state_dict(cfg, seed=3)
    images = torch.randn(batch, 3, cfg.image, cfg.image)
    m = _torchvision_vit(cfg, sd)
    with torch.no_grad():
        stage = []
        logits = vo.vit_forward(sd, cfg, images, stage)
        ref_logits = m(images)
        ref_probs = _attention_probs(m, cfg, images)
    assert rel_err(logits, ref_logits) < 1e-5
    assert len(stage) == len(ref_probs) == cfg.depth
    for a, r in zip(stage, ref_probs):
        assert a.shape == r.shape and rel_err(a, r) < 1e-5
