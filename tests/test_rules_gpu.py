"""GPU: rule kernels (rules 5, 6, 7, 10, 11, eq. 8-9, rollout) through the C ABI vs the reference goldens and the
oracle on seeded inputs, including ragged / unaligned / empty shapes."""
import os

import numpy as np
import pytest
import torch

from oracle import rules as orules
from util import rel_err, TOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mmx():
    import mmx_b200
    assert torch.cuda.is_available()
    return mmx_b200


@pytest.fixture(scope="module")
def g(golden_dir):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, "rules.npz")).items()}


def test_golden_rules(mmx, g):
    c = {k: v.cuda() for k, v in g.items()}
    assert rel_err(mmx.avg_heads(c["cam_ss"], c["grad_ss"]), g["abar_ss"]) < 1e-6
    assert rel_err(mmx.avg_heads(c["cam_sq"], c["grad_sq"]), g["abar_sq"]) < 1e-6
    a, b = mmx.apply_self_attention_rules(c["R_ss"], c["R_sq"], c["abar_ss"])
    assert rel_err(a, g["self.R_ss_add"]) < 1e-6 and rel_err(b, g["self.R_sq_add"]) < 1e-6
    assert rel_err(mmx.handle_residual(c["R_ss"]), g["hr.R_ss"]) < 1e-6
    assert rel_err(mmx.handle_residual(c["R_qq"]), g["hr.R_qq"]) < 1e-6
    for norm in (True, False):
        for s10 in (True, False):
            key = f"n{int(norm)}s{int(s10)}"
            d = mmx.apply_mm_attention_rules(c["R_ss"], c["R_qq"], c["abar_sq"], norm, s10)
            assert rel_err(d, g[f"mm_detr.{key}"]) < 1e-5
            x, y = mmx.apply_mm_attention_rules_lxmert(c["R_ss"], c["R_qq"], c["R_qs"], c["abar_sq"], norm, s10)
            assert rel_err(x, g[f"mm_lx.{key}.sq"]) < 1e-5 and rel_err(y, g[f"mm_lx.{key}.ss"]) < 1e-5
    for sl in (0, 2):
        r = mmx.compute_rollout_attention(list(c["rollout.mats"]), sl)
        assert rel_err(r, g[f"rollout.sl{sl}"]) < 1e-5


def test_nan_guard_and_assert(mmx, g):
    T, S = g["R_ss"].shape[0], g["R_qq"].shape[0]
    d = mmx.apply_mm_attention_rules(torch.eye(T).cuda(), torch.eye(S).cuda(), g["abar_sq"].cuda())
    assert torch.equal(d.cpu(), g["mm_detr.identity_state"])        # 0/0 -> NaN -> 0, exactly like DETR
    x, _ = mmx.apply_mm_attention_rules_lxmert(torch.eye(T).cuda(), torch.eye(S).cuda(), torch.zeros(S, T).cuda(),
                                               g["abar_sq"].cuda())
    assert torch.isnan(x).all()                                      # LXMERT has no guard
    R = torch.eye(4)
    R[1, 1] = 0.5
    with pytest.raises(AssertionError):
        mmx.handle_residual(R.cuda())


@pytest.mark.parametrize("B,H,T,S", [(1, 12, 50, 50), (3, 8, 77, 77), (2, 12, 20, 36), (1, 8, 100, 625), (5, 1, 1, 1),
                                     (2, 3, 7, 13), (64, 12, 50, 52)])
def test_avg_heads_shapes(mmx, B, H, T, S):
    gen = torch.Generator().manual_seed(B * 1000 + S)
    A = torch.softmax(torch.randn(B, H, T, S, generator=gen), -1)
    G = torch.randn(B, H, T, S, generator=gen)
    ref = orules.avg_heads_batched(A, G, B)
    out = mmx.avg_heads_batched(A.cuda(), G.cuda(), B)
    assert rel_err(out, ref) < 1e-6
    # linearity in the gradient on the positive cone: scaling dA by c > 0 scales Abar by c
    out2 = mmx.avg_heads_batched(A.cuda(), (3.0 * G).cuda(), B)
    assert rel_err(out2, 3.0 * ref) < 1e-6


def test_avg_heads_full_size_property(mmx):
    """C2 size (12 layers x 64 samples, text tower): Abar >= 0, zero where A is zero (causal mask), and equal to the
    oracle on a slice."""
    B, H, S = 12 * 64, 8, 77
    gen = torch.Generator(device="cuda").manual_seed(1)
    logits = torch.randn(B, H, S, S, device="cuda", generator=gen)
    mask = torch.full((S, S), float("-inf"), device="cuda").triu_(1)
    A = torch.softmax(logits + mask, -1)
    G = torch.randn(B, H, S, S, device="cuda", generator=gen)
    out = mmx.avg_heads_batched(A, G, B)
    assert (out >= 0).all()
    assert (out.triu(1) == 0).all()
    ref = orules.avg_heads_batched(A[:4].cpu(), G[:4].cpu(), 4)
    assert rel_err(out[:4], ref) < 1e-6


@pytest.mark.parametrize("B,S,Q", [(1, 50, 0), (4, 77, 0), (2, 20, 36), (2, 36, 20), (1, 100, 625), (1, 197, 0), (3, 1, 1),
                                   (2, 256, 300), (1, 625, 0), (2, 577, 0), (5, 130, 150), (16, 200, 0)])
def test_self_update(mmx, B, S, Q):
    # S >= 128 runs on the tensor cores (ONE batched tcgen05 3xTF32 launch, tile -> sample, "+R" fused in the epilogue; a
    # tile's operand boxes run into the next sample's rows when S % 128 != 0); smaller S on the FFMA kernel
    tol = 1e-6 if S < 128 else 1e-5
    gen = torch.Generator().manual_seed(S)
    Ab = torch.rand(B, S, S, generator=gen) * 0.01
    R = torch.eye(S).expand(B, S, S) + torch.rand(B, S, S, generator=gen) * 0.01
    Rq = torch.rand(B, S, Q, generator=gen) if Q else None
    o, oq = mmx.self_update(R.cuda(), Ab.cuda(), Rq.cuda() if Q else None)
    assert rel_err(o, (R.double() + torch.bmm(Ab.double(), R.double()))) < tol
    if Q:
        assert rel_err(oq, (Rq.double() + torch.bmm(Ab.double(), Rq.double()))) < tol


def test_self_update_chain_matches_product(mmx):
    """Size-independent property: after L updates from R=I, R == (I+A_L)...(I+A_1) (checked in fp64)."""
    S, L = 77, 12
    gen = torch.Generator().manual_seed(0)
    As = [torch.rand(1, S, S, generator=gen) * 0.02 for _ in range(L)]
    R = torch.eye(S).unsqueeze(0).cuda()
    for a in As:
        R, _ = mmx.self_update(R, a.cuda())
    P = torch.eye(S, dtype=torch.float64)
    for a in As:
        P = (torch.eye(S, dtype=torch.float64) + a[0].double()) @ P
    assert rel_err(R[0], P) < 1e-5


@pytest.mark.parametrize("T,S", [(20, 36), (36, 20), (100, 625), (5, 3)])
def test_mm_update_vs_oracle(mmx, T, S):
    gen = torch.Generator().manual_seed(T * 7 + S)
    R_ss = torch.eye(T) + torch.rand(T, T, generator=gen) * 0.05
    R_qq = torch.eye(S) + torch.rand(S, S, generator=gen) * 0.05
    R_qs = torch.rand(S, T, generator=gen) * 0.05
    cam = torch.rand(T, S, generator=gen) * 0.05
    for norm in (True, False):
        for s10 in (True, False):
            x, y = mmx.apply_mm_attention_rules_lxmert(R_ss.cuda(), R_qq.cuda(), R_qs.cuda(), cam.cuda(), norm, s10)
            ox, oy = orules.apply_mm_attention_rules_lxmert(R_ss, R_qq, R_qs, cam, norm, s10)
            assert rel_err(x, ox) < 1e-5 and rel_err(y, oy) < 1e-5
            d = mmx.apply_mm_attention_rules(R_ss.cuda(), R_qq.cuda(), cam.cuda(), norm, s10)
            assert rel_err(d, orules.apply_mm_attention_rules_detr(R_ss, R_qq, cam, norm, s10)) < 1e-5


def test_rollout_batched_visualbert_variant(mmx):
    gen = torch.Generator().manual_seed(5)
    mats = [torch.softmax(torch.randn(3, 30, 30, generator=gen), -1) for _ in range(4)]
    out = mmx.compute_rollout_attention([m.cuda() for m in mats], 1, normalize=False)
    ref = orules.compute_rollout_attention(mats, 1, normalize=False)
    assert rel_err(out, ref) < 1e-5
    out = mmx.compute_rollout_attention([m.cuda() for m in mats], 0, normalize=True)
    assert rel_err(out, orules.compute_rollout_attention(mats, 0, normalize=True)) < 1e-5
    # rows of a normalised rollout sum to 1
    assert rel_err(out.sum(-1), torch.ones(3, 30)) < 1e-5


@pytest.mark.parametrize("S,L,B", [(50, 12, 5), (77, 12, 3), (20, 9, 4), (36, 5, 2), (100, 6, 2), (128, 2, 1), (7, 1, 3), (77, 1, 2)])
def test_rule6_chain_kernel(S, L, B):
    """mmx_self_chain: R = I; R <- R + Abar_l R for all layers in ONE launch (R resident in shared memory, mma.sync 3xTF32)
    vs the per-layer fp64 loop of CLIP_explainability.ipynb:169-183."""
    import ctypes as C
    import mmx_b200
    from mmx_b200._lib import lib, check, ptr, current_stream
    g = torch.Generator().manual_seed(S * 100 + L)
    ld = (S + 3) // 4 * 4
    Ab = torch.zeros(L, B, S, ld)
    Ab[..., :S] = torch.rand(L, B, S, S, generator=g) * (2.0 / S)           # rule-5 outputs: non-negative, rows of O(1) mass
    if B > 1:
        Ab[:, 1, S // 2:, :] = 0                                            # a ragged sample: dead rows / columns stay identity
        Ab[:, 1, :, S // 2:] = 0
    out = torch.full((B, S, ld), float("nan"), device="cuda")
    check(lib().mmx_self_chain(ptr(Ab.cuda()), B * S * ld, ld, ptr(out), ld, B, S, L, current_stream()))
    R = torch.eye(S, dtype=torch.float64).repeat(B, 1, 1)
    for l in range(L):
        R = R + Ab[l, :, :, :S].double() @ R
    assert rel_err(out[..., :S], R) < 1e-5
    if ld > S:
        assert (out[..., S:] == 0).all()
    if B > 1:
        assert torch.equal(out[1, S // 2:, :S].cpu(), torch.eye(S)[S // 2:])
    # one sample alone == inside the batch, bit for bit
    one = torch.empty(1, S, ld, device="cuda")
    check(lib().mmx_self_chain(ptr(Ab[:, :1].contiguous().cuda()), S * ld, ld, ptr(one), ld, 1, S, L, current_stream()))
    assert torch.equal(one[0], out[0])
