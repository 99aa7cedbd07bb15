"""GPU: the transformer primitives that produce A and dA (linear, LayerNorm, attention fwd/bwd) through the C ABI vs
plain PyTorch fp32/fp64 autograd on the CPU (floating-point kernels: tolerance stated per test)."""
import ctypes as C
import math

import pytest
import torch
import torch.nn.functional as F

from util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    import mmx_b200
    from mmx_b200._lib import lib, check, ptr, current_stream
    return dict(lib=lib(), check=check, ptr=ptr, st=current_stream)


def _pack(lib, check, ptr, st, Wd):
    """fp16x3 packed copy of a device [N, K] matrix (None below the tensor-core kernel's minimum tile)."""
    N, K = Wd.shape
    buf = torch.empty(lib.mmx_pack_weight_bytes(N, K), dtype=torch.uint8, device="cuda")
    check(lib.mmx_pack_weight(ptr(Wd), K, N, K, ptr(buf), st()))
    return buf


ACTS = {0: lambda x: x, 1: lambda x: x * torch.sigmoid(1.702 * x), 2: lambda x: F.gelu(x), 3: lambda x: F.relu(x)}


@pytest.mark.parametrize("backend", [0, 1, 2, 3])
@pytest.mark.parametrize("M,N,K", [(3200, 2304, 768), (4928, 512, 2048), (64, 512, 768), (37, 96, 64), (1, 8, 4),
                                   (130, 260, 36), (3136, 768, 3072)])
def test_linear_and_dgrad(L, backend, M, N, K):
    lib, check, ptr, st = L["lib"], L["check"], L["ptr"], L["st"]
    got = lib.mmx_set_gemm_backend(backend)
    if got != backend:
        pytest.skip("tcgen05 backend not available")
    try:
        # fp32 FFMA: plain fp32 rounding.  tcgen05 3xTF32 / fp16x3: the tensor core accumulates with truncation, measured
        # ~2e-6 (K=768) .. 6e-6 (K=3072) of max|C|; all far inside the 1e-4 budget of the maps.  Backend 2 takes the
        # packed fp16 hi / lo planes of the weight (mmx_pack_weight) beside the fp32 pointer; backend 3 is the same arithmetic
        # with A pre-split by a separate pass and the product on CTA pairs (cta_group::2, 256-row tiles).
        tol = 5e-6 if backend == 0 else 2e-5
        gen = torch.Generator().manual_seed(M + N + K)
        A = torch.randn(M, K, generator=gen)
        W = torch.randn(N, K, generator=gen) / math.sqrt(K)
        bias = torch.randn(N, generator=gen)
        res = torch.randn(M, N, generator=gen)
        for act in (0, 1, 2, 3):
            Ad, Wd, bd, rd = A.cuda(), W.cuda(), bias.cuda(), res.cuda()
            Cd = torch.empty(M, N, device="cuda")
            Ca = torch.empty(M, N, device="cuda") if act else None
            pk = _pack(lib, check, ptr, st, Wd) if backend >= 2 else None
            check(lib.mmx_linear_packed(ptr(Ad), K, ptr(Wd), K, ptr(pk), ptr(bd), ptr(rd), N, ptr(Cd), N, ptr(Ca), act, M, N, K,
                                        st()))
            ref = (A.double() @ W.double().t() + bias.double() + res.double())
            assert rel_err(Cd, ref) < tol
            if act:
                assert rel_err(Ca, ACTS[act](ref)) < tol
        # dgrad through an activation: dX = (dY W) . act'(pre)
        dY = torch.randn(M, N, generator=gen)
        pre = torch.randn(M, K, generator=gen)
        Wt = W.t().contiguous()
        dYd, Wtd, pred = dY.cuda(), Wt.cuda(), pre.cuda()      # keep device tensors alive across the call
        pkt = _pack(lib, check, ptr, st, Wtd) if backend >= 2 else None
        for act in (0, 1, 2, 3):
            dX = torch.empty(M, K, device="cuda")
            check(lib.mmx_linear_dgrad_packed(ptr(dYd), N, ptr(Wtd), N, ptr(pkt), ptr(pred) if act else None, K, act,
                                              ptr(dX), K, M, N, K, st()))
            p = pre.double().requires_grad_(True)
            y = ACTS[act](p) @ W.double().t()
            y.backward(dY.double())
            base = dY.double() @ W.double()
            ref = p.grad if act else base
            assert rel_err(dX, ref, base=base) < tol
    finally:
        lib.mmx_set_gemm_backend(2)


def test_f16x3_range_and_tiny_operands(L):
    """fp16x3 backend: operands spanning fp16's useful window (rows of A scaled from 1e-4 to 1e3) keep the fp32-faithful
    error, an operand above 65504 yields a non-finite result (loud), and gradient-sized operands (1e-7) keep an ABSOLUTE
    error of ~2^-35 |W| K (documented limit: scale such streams by a power of two)."""
    lib, check, ptr, st = L["lib"], L["check"], L["ptr"], L["st"]
    if lib.mmx_set_gemm_backend(2) != 2:
        pytest.skip("tcgen05 backend not available")
    M, N, K = 256, 256, 512
    gen = torch.Generator().manual_seed(11)
    A = torch.randn(M, K, generator=gen) * torch.logspace(-4, 3, M).unsqueeze(1)
    W = torch.randn(N, K, generator=gen) / math.sqrt(K)
    Ad, Wd = A.cuda(), W.cuda()
    pk = _pack(lib, check, ptr, st, Wd)
    Cd = torch.empty(M, N, device="cuda")
    check(lib.mmx_linear_packed(ptr(Ad), K, ptr(Wd), K, ptr(pk), None, None, 0, ptr(Cd), N, None, 0, M, N, K, st()))
    ref = A.double() @ W.double().t()
    row_err = (Cd.cpu().double() - ref).abs().amax(1) / ref.abs().amax(1)
    assert row_err.max().item() < 2e-5, row_err.max().item()           # every row, whatever its magnitude
    A2 = A.clone(); A2[3, 5] = 7.0e4                                      # above fp16's largest finite value
    A2d = A2.cuda()
    check(lib.mmx_linear_packed(ptr(A2d), K, ptr(Wd), K, ptr(pk), None, None, 0, ptr(Cd), N, None, 0, M, N, K, st()))
    assert not torch.isfinite(Cd[3]).all() and torch.isfinite(Cd[4]).all()
    A3 = (torch.randn(M, K, generator=gen) * 1e-7)
    A3d = A3.cuda()
    check(lib.mmx_linear_packed(ptr(A3d), K, ptr(Wd), K, ptr(pk), None, None, 0, ptr(Cd), N, None, 0, M, N, K, st()))
    ref3 = A3.double() @ W.double().t()
    assert (Cd.cpu().double() - ref3).abs().max().item() < 2.0 ** -35 * K * W.abs().max().item()


@pytest.mark.parametrize("backend", [1, 2])
@pytest.mark.parametrize("M,N,K", [(3200, 768, 768), (3200, 2304, 768), (300, 260, 200), (129, 1000, 3072), (1, 132, 68)])
def test_tile_width_never_changes_a_bit(L, backend, M, N, K):
    """The tcgen05 kernel picks its tile width (128 / 144 / 160 columns) from the tile count; each output element
    still sums its K products in the same order, so all three widths must agree bit for bit (batch invariance and
    sharded == single-GPU rest on this) and match fp64 to the 3xTF32 tolerance."""
    lib, check, ptr, st = L["lib"], L["check"], L["ptr"], L["st"]
    if lib.mmx_set_gemm_backend(backend) != backend:
        pytest.skip("tcgen05 backend not available")
    try:
        gen = torch.Generator().manual_seed(7 * M + N + K)
        A = torch.randn(M, K, generator=gen)
        W = torch.randn(N, K, generator=gen) / math.sqrt(K)
        bias, res = torch.randn(N, generator=gen), torch.randn(M, N, generator=gen)
        Ad, Wd, bd, rd = A.cuda(), W.cuda(), bias.cuda(), res.cuda()
        pk = _pack(lib, check, ptr, st, Wd) if backend == 2 else None
        outs = {}
        for bn in (128, 144, 160, 0):
            assert lib.mmx_set_gemm_tile_n(bn) == bn
            Cd = torch.full((M, N), float("nan"), device="cuda")
            Ca = torch.full((M, N), float("nan"), device="cuda")
            check(lib.mmx_linear_packed(ptr(Ad), K, ptr(Wd), K, ptr(pk), ptr(bd), ptr(rd), N, ptr(Cd), N, ptr(Ca), 2, M, N, K,
                                        st()))
            outs[bn] = (Cd.cpu(), Ca.cpu())
        ref = A.double() @ W.double().t() + bias.double() + res.double()
        assert rel_err(outs[128][0], ref) < 2e-5
        for bn in (144, 160, 0):
            assert torch.equal(outs[bn][0], outs[128][0]), bn
            assert torch.equal(outs[bn][1], outs[128][1]), bn
    finally:
        lib.mmx_set_gemm_tile_n(0)
        lib.mmx_set_gemm_backend(2)


@pytest.mark.parametrize("rows,D", [(3200, 768), (77, 512), (5, 32), (3, 1024), (9, 100)])
def test_layernorm_fwd_bwd(L, rows, D):
    lib, check, ptr, st = L["lib"], L["check"], L["ptr"], L["st"]
    gen = torch.Generator().manual_seed(rows + D)
    x = torch.randn(rows, D, generator=gen) * 2 + 0.5
    g = 1 + 0.1 * torch.randn(D, generator=gen)
    b = 0.1 * torch.randn(D, generator=gen)
    dy = torch.randn(rows, D, generator=gen)
    resid = torch.randn(rows, D, generator=gen)
    xd, gd, bd, dyd, resd = x.cuda(), g.cuda(), b.cuda(), dy.cuda(), resid.cuda()   # kept alive across the calls
    y = torch.empty_like(xd); mean = torch.empty(rows, device="cuda"); rstd = torch.empty(rows, device="cuda")
    check(lib.mmx_layernorm_fwd(ptr(xd), D, None, ptr(gd), ptr(bd), ptr(y), D, ptr(mean), ptr(rstd), rows, D,
                                C.c_float(1e-5), st()))
    xr = x.double().requires_grad_(True)
    yr = F.layer_norm(xr, (D,), g.double(), b.double(), 1e-5)
    assert rel_err(y, yr.detach()) < 2e-6
    yr.backward(dy.double())
    dx = torch.empty_like(xd)
    check(lib.mmx_layernorm_bwd(ptr(dyd), D, ptr(xd), D, None, ptr(gd), ptr(mean), ptr(rstd), ptr(resd),
                                D, ptr(dx), D, rows, D, st()))
    assert rel_err(dx, xr.grad + resid.double()) < 2e-6
    # gathered rows (the pooled cls / eot token): output row r reads x[row_map[r]], dx scatters back to it
    if rows >= 3:
        rm = torch.tensor([rows - 1, 0, rows // 2], dtype=torch.int32)
        rmd, dy3 = rm.cuda(), dy[:3].contiguous().cuda()
        yg = torch.empty(3, D, device="cuda"); mg = torch.empty(3, device="cuda"); rg = torch.empty(3, device="cuda")
        check(lib.mmx_layernorm_fwd(ptr(xd), D, ptr(rmd), ptr(gd), ptr(bd), ptr(yg), D, ptr(mg), ptr(rg), 3,
                                    D, C.c_float(1e-5), st()))
        assert rel_err(yg, yr.detach()[rm.long()]) < 2e-6
        dxs = torch.zeros_like(xd)
        check(lib.mmx_layernorm_bwd(ptr(dy3), D, ptr(xd), D, ptr(rmd), ptr(gd), ptr(mg), ptr(rg), None,
                                    0, ptr(dxs), D, 3, D, st()))
        x2 = x.double().requires_grad_(True)
        F.layer_norm(x2, (D,), g.double(), b.double(), 1e-5)[rm.long()].backward(dy[:3].double())
        assert rel_err(dxs, x2.grad) < 2e-6


def _attn_ref(q, k, v, H, scale, causal, key_bias, scale_scores):
    B, T, Dm = q.shape
    S = k.shape[1]
    hd = Dm // H
    qh = q.view(B, T, H, hd).transpose(1, 2)
    kh = k.view(B, S, H, hd).transpose(1, 2)
    vh = v.view(B, S, H, hd).transpose(1, 2)
    s = ((qh @ kh.transpose(-1, -2)) * scale) if scale_scores else ((qh * scale) @ kh.transpose(-1, -2))
    if key_bias is not None:
        s = s + key_bias[:, None, None, :]
    if causal:
        s = s + torch.full((T, S), float("-inf"), dtype=s.dtype).triu_(1)
    A = torch.softmax(s, -1)
    A.retain_grad()
    o = (A @ vh).transpose(1, 2).reshape(B, T, Dm)
    return A, o


@pytest.mark.parametrize("B,H,T,S,hd,causal,bias,ss", [
    (2, 12, 50, 50, 64, 0, 0, 0), (2, 8, 77, 77, 64, 1, 0, 0), (1, 2, 8, 8, 16, 1, 0, 0), (2, 2, 17, 17, 64, 0, 0, 0),
    (2, 12, 20, 36, 64, 0, 1, 1), (2, 12, 36, 20, 64, 0, 1, 1), (1, 8, 100, 625, 32, 0, 0, 0), (1, 8, 625, 625, 32, 0, 0, 0),
    (1, 16, 577, 577, 64, 0, 0, 0), (1, 2, 1, 1, 32, 0, 0, 0),
    (1, 8, 100, 850, 32, 0, 0, 0), (1, 8, 850, 850, 32, 0, 0, 0)])      # DETR at 800x1066: 850 keys -> the 32-row tile variant
def test_attention_fwd_bwd(L, B, H, T, S, hd, causal, bias, ss):
    lib, check, ptr, st = L["lib"], L["check"], L["ptr"], L["st"]
    gen = torch.Generator().manual_seed(T * 31 + S)
    Dm = H * hd
    q = torch.randn(B, T, Dm, generator=gen)
    k = torch.randn(B, S, Dm, generator=gen)
    v = torch.randn(B, S, Dm, generator=gen)
    dO = torch.randn(B, T, Dm, generator=gen)
    kb = (torch.randn(B, S, generator=gen) if bias else None)
    scale = 1.0 / math.sqrt(hd)
    flags = (1 if causal else 0) | (2 if ss else 0)
    ldA = (S + 3) // 4 * 4
    qd, kd, vd, dOd = q.cuda(), k.cuda(), v.cuda(), dO.cuda()
    kbd = kb.cuda() if bias else None
    A = torch.full((B, H, T, ldA), 7.0, device="cuda"); O = torch.empty(B, T, Dm, device="cuda")
    check(lib.mmx_attention_fwd(ptr(qd), Dm, ptr(kd), Dm, ptr(vd), Dm, ptr(kbd), ptr(A), ldA, ptr(O), Dm,
                                B, H, T, S, hd, C.c_float(scale), flags, st()))
    q64, k64, v64 = (t.double().requires_grad_(True) for t in (q, k, v))
    Ar, Or = _attn_ref(q64, k64, v64, H, scale, causal, kb.double() if bias else None, ss)
    # the products run as 3xTF32 tensor-core passes (same budget as the linear GEMMs), softmax in fp32
    errs = {"A": rel_err(A[..., :S], Ar.detach())}
    assert errs["A"] < 5e-6
    assert (A[..., S:] == 0).all()                       # padded columns are zero-filled (rule 5 relies on it)
    errs["O"] = rel_err(O, Or.detach())
    assert errs["O"] < 5e-6
    Or.backward(dO.double())
    dA = torch.full((B, H, T, ldA), 7.0, device="cuda"); delta = torch.empty(B, H, T, device="cuda")
    dq, dk, dv = (torch.empty_like(t) for t in (qd, kd, vd))
    check(lib.mmx_attention_bwd(ptr(dOd), Dm, ptr(qd), Dm, ptr(kd), Dm, ptr(vd), Dm, ptr(A), ptr(dA), ldA, ptr(delta),
                                ptr(dq), Dm, ptr(dk), Dm, ptr(dv), Dm, B, H, T, S, hd, C.c_float(scale), flags, st()))
    errs["dA"] = rel_err(dA[..., :S], Ar.grad)         # the tensor the reference's backward hook captures
    assert errs["dA"] < 5e-6
    assert (dA[..., S:] == 0).all()
    errs.update(dq=rel_err(dq, q64.grad), dk=rel_err(dk, k64.grad), dv=rel_err(dv, v64.grad))
    print("attention errors", (B, H, T, S, hd), {k_: f"{v_:.1e}" for k_, v_ in errs.items()})
    assert max(errs["dq"], errs["dk"], errs["dv"]) < 1e-5
    # stop-after-dA form (last relevant block)
    dA2 = torch.empty_like(dA)
    check(lib.mmx_attention_bwd(ptr(dOd), Dm, ptr(qd), Dm, ptr(kd), Dm, ptr(vd), Dm, ptr(A), ptr(dA2), ldA, None,
                                None, 0, None, 0, None, 0, B, H, T, S, hd, C.c_float(scale), flags, st()))
    assert torch.equal(dA2, dA)
