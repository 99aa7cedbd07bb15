"""GPU: ViT generate_relevance vs the oracle restatement (model forward: parity unpinned, see
oracle/vit_oracle.py; rule: pinned by tests/golden/rules.npz)."""
import pytest
import torch

from oracle import vit_oracle as vo
from util import rel_err, TOL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,B,index", [(vo.VIT_TINY, 3, None), (vo.VIT_TINY, 2, [1, 7]), (vo.VIT_B16, 1, None)])
def test_generate_relevance(cfg, B, index):
    import mmx_b200
    sd = vo.init_state_dict(cfg, seed=3)
    g = torch.Generator().manual_seed(11)
    images = torch.randn(B, 3, cfg.image, cfg.image, generator=g)
    ref, ref_logits = vo.generate_relevance(sd, cfg, images, index)
    eng = mmx_b200.ViTEngine(sd, heads=cfg.heads, device="cuda:0")
    out = mmx_b200.generate_relevance(eng, images.cuda(), index)
    assert rel_err(eng.logits, ref_logits) < TOL
    if B == 1:
        assert out.shape == (cfg.tokens - 1,)            # the notebook returns R[0, 1:] of one image
        out = out[None]
    assert rel_err(out, ref) < TOL
    # stage check on the last block: A and dA as the reference hooks would see them
    A = eng.blocks[-1]["attn"].get_attn()
    assert A.shape == (B, cfg.heads, cfg.tokens, cfg.tokens)
    assert rel_err(A.sum(-1), torch.ones(B, cfg.heads, cfg.tokens)) < 1e-5
