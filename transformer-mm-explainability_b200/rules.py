"""Reference-named rule functions (torch CUDA tensors in, torch CUDA tensors out) over the libmmx rule kernels.

Signatures mirror DETR/modules/ExplanationGenerator.py:5-53 and lxmert/lxmert/src/ExplanationGenerator.py:5-54.
Inputs must be fp32 CUDA tensors; nothing is computed by PyTorch - it only owns the memory and the stream.
"""
from __future__ import annotations

import torch

from ._lib import lib, check, ptr, current_stream, MmxError

MM_NORMALIZE, MM_SELF_IN_10, MM_NAN_TO_ZERO = 1, 2, 4


def _prep(t: torch.Tensor) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise MmxError("mmx_b200 rule functions need CUDA tensors (no CPU fallback)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def avg_heads_batched(cam: torch.Tensor, grad: torch.Tensor, batch: int) -> torch.Tensor:
    """Rule 5 for a batch; cam/grad [batch*H, T, S] or [batch, H, T, S] (head index b*H+h,
    CLIP_explainability.ipynb:176-181).  Returns [batch, T, S]."""
    cam, grad = _prep(cam), _prep(grad)
    T, S = cam.shape[-2], cam.shape[-1]
    H = cam.numel() // (batch * T * S)
    out = torch.empty(batch, T, S, device=cam.device, dtype=torch.float32)
    check(lib().mmx_avg_heads(ptr(cam), ptr(grad), ptr(out), batch, H, T, S, S, S, current_stream()))
    return out


def avg_heads_record(rec, batch: int, use_cam: bool = False) -> torch.Tensor:
    """Rule 5 straight from an attention record (nn.AttnRecord): reads the zero-padded staged A / dA in place with the
    128-bit kernel and returns a [B,T,S] VIEW of a zero-padded [B,T,ld] buffer (row stride ld = round_up(S,4)), the
    layout the tensor-core R update wants."""
    A, dA, ld = rec.padded(use_cam)
    if dA is None:
        raise MmxError("no attention gradient recorded (backward did not reach this attention)")
    B, H, T = A.shape[0], A.shape[1], A.shape[2]
    assert B == batch
    out = torch.empty(B, T, ld, device=A.device, dtype=torch.float32)
    # the pad columns of A and dA are zero, so the plane is treated as dense [T, ld]
    check(lib().mmx_avg_heads(ptr(A), ptr(dA), ptr(out), B, H, T, ld, ld, ld, current_stream()))
    return out[..., :rec.S]


def _padded3(t: torch.Tensor):
    """[B,R,C] fp32 CUDA tensor -> (tensor whose rows are 16-byte aligned with zero pads, ld).  Views produced by this
    module (row stride = round_up(C,4), dense planes) pass through without a copy."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise MmxError("mmx_b200 rule functions need CUDA tensors (no CPU fallback)")
    B, R, Cc = t.shape
    ldp = (Cc + 3) // 4 * 4
    if (t.dtype == torch.float32 and t.stride(2) == 1 and t.stride(1) == ldp and (B == 1 or t.stride(0) == R * ldp)
            and t.data_ptr() % 16 == 0):
        return t, ldp
    buf = torch.zeros(B, R, ldp, device=t.device, dtype=torch.float32)
    buf[..., :Cc] = t
    return buf[..., :Cc], ldp


def avg_heads(cam: torch.Tensor, grad: torch.Tensor) -> torch.Tensor:
    """Rule 5 (DETR/modules/ExplanationGenerator.py:19-24): all leading dims are heads of ONE sample."""
    return avg_heads_batched(cam, grad, 1)[0]


def self_update(R_ss: torch.Tensor, cam_ss: torch.Tensor, R_sq: torch.Tensor | None = None):
    """Fused rules 6+7 including the caller's ``+=``: returns (R_ss + cam*R_ss, R_sq + cam*R_sq).  Accepts [S,S] or
    batched [B,S,S]."""
    squeeze = R_ss.dim() == 2
    R, ld = _padded3(R_ss if not squeeze else R_ss.unsqueeze(0))
    Ab, lda = _padded3(cam_ss if not squeeze else cam_ss.unsqueeze(0))
    B, S = R.shape[0], R.shape[-1]
    out = torch.zeros(B, S, ld, device=R.device, dtype=torch.float32)[..., :S]     # zero pads, 16-byte aligned rows
    Rq = outq = None
    Q, ldq = 0, 4
    if R_sq is not None:
        Rq, ldq = _padded3(R_sq if not squeeze else R_sq.unsqueeze(0))
        Q = Rq.shape[-1]
        outq = torch.zeros(B, S, ldq, device=R.device, dtype=torch.float32)[..., :Q]
    check(lib().mmx_self_update(ptr(Ab), lda, ptr(R), ptr(out), ld, ptr(Rq), ptr(outq), ldq, B, S, Q, current_stream()))
    if squeeze:
        return out[0], (outq[0] if outq is not None else None)
    return out, outq


def apply_self_attention_rules(R_ss, R_sq, cam_ss):
    """Rules 6+7, reference signature and return order ``(R_ss_addition, R_sq_addition)``
    (DETR/modules/ExplanationGenerator.py:27-30)."""
    R, Rq, Ab = _prep(R_ss).unsqueeze(0), _prep(R_sq).unsqueeze(0), _prep(cam_ss).unsqueeze(0)
    S, Q = R.shape[-1], Rq.shape[-1]
    add_ss = torch.empty_like(R)
    add_sq = torch.empty_like(Rq)
    l = lib()
    check(l.mmx_bmm_add(ptr(Ab), S, S * S, 0, ptr(R), S, S * S, None, 0, 0, ptr(add_ss), S, S * S, 1, S, S, S, current_stream()))
    check(l.mmx_bmm_add(ptr(Ab), S, S * S, 0, ptr(Rq), Q, S * Q, None, 0, 0, ptr(add_sq), Q, S * Q, 1, S, Q, S, current_stream()))
    return add_ss[0], add_sq[0]


def handle_residual(orig_self_attention: torch.Tensor) -> torch.Tensor:
    """Eq. 8-9 (DETR/modules/ExplanationGenerator.py:46-53), including the reference's
    ``assert diag(R - I).min() >= 0`` (one host sync, like the reference)."""
    R = _prep(orig_self_attention)
    squeeze = R.dim() == 2
    if squeeze:
        R = R.unsqueeze(0)
    B, S = R.shape[0], R.shape[-1]
    out = torch.empty_like(R)
    md = torch.empty(1, device=R.device, dtype=torch.float32)
    check(lib().mmx_handle_residual(ptr(R), ptr(out), S, B, S, ptr(md), current_stream()))
    assert md.item() >= 0
    return out[0] if squeeze else out


def _mm(R_ss, R_qq, R_qs, cam_sq, flags):
    Rs, Rq, Ab = _prep(R_ss).unsqueeze(0), _prep(R_qq).unsqueeze(0), _prep(cam_sq).unsqueeze(0)
    T, S = Rs.shape[-1], Rq.shape[-1]
    l = lib()
    ws = torch.empty(l.mmx_mm_update_workspace(1, T, S) // 4 + 4, device=Rs.device, dtype=torch.float32)
    sq_add = torch.empty(1, T, S, device=Rs.device, dtype=torch.float32)
    Rqs = ss_add = None
    if R_qs is not None:
        Rqs = _prep(R_qs).unsqueeze(0)
        ss_add = torch.empty(1, T, T, device=Rs.device, dtype=torch.float32)
    md = torch.zeros(2, device=Rs.device, dtype=torch.float32)
    check(l.mmx_mm_update(ptr(Rs), T, ptr(Rq), S, ptr(Rqs), T, ptr(Ab), S, ptr(sq_add), S, ptr(ss_add), T, 1, T, S, flags,
                          ptr(ws), ptr(md), current_stream()))
    if flags & MM_NORMALIZE:
        assert md.min().item() >= 0          # the reference's assert in handle_residual (also when rule 10 drops the product)
    return sq_add[0], (ss_add[0] if ss_add is not None else None)


def apply_mm_attention_rules(R_ss, R_qq, cam_sq, apply_normalization=True, apply_self_in_rule_10=True):
    """Rule 10, DETR form (DETR/modules/ExplanationGenerator.py:33-43): returns ``R_sq_addition`` with NaN -> 0."""
    flags = MM_NAN_TO_ZERO | (MM_NORMALIZE if apply_normalization else 0) | (MM_SELF_IN_10 if apply_self_in_rule_10 else 0)
    return _mm(R_ss, R_qq, None, cam_sq, flags)[0]


def apply_mm_attention_rules_lxmert(R_ss, R_qq, R_qs, cam_sq, apply_normalization=True, apply_self_in_rule_10=True):
    """Rules 10+11, LXMERT form (lxmert/lxmert/src/ExplanationGenerator.py:32-42): returns
    ``(R_sq_addition, R_ss_addition)``; no NaN guard."""
    flags = (MM_NORMALIZE if apply_normalization else 0) | (MM_SELF_IN_10 if apply_self_in_rule_10 else 0)
    return _mm(R_ss, R_qq, R_qs, cam_sq, flags)


def compute_rollout_attention(all_layer_matrices, start_layer=0, normalize=True):
    """Rollout (DETR/modules/ExplanationGenerator.py:5-16).  Matrices [S,S] (DETR/LXMERT form) or [B,S,S]
    (VisualBERT form, which passes ``normalize=False``)."""
    mats = torch.stack([_prep(m) for m in all_layer_matrices])
    squeeze = mats.dim() == 3
    if squeeze:
        mats = mats.unsqueeze(1)
    L, B, S = mats.shape[0], mats.shape[1], mats.shape[-1]
    mats = mats.contiguous()
    out = torch.empty(B, S, S, device=mats.device, dtype=torch.float32)
    ws = torch.empty(2, B, S, S, device=mats.device, dtype=torch.float32)
    check(lib().mmx_rollout(ptr(mats), L, B, S, start_layer, 1 if normalize else 0, ptr(out), ptr(ws), current_stream()))
    return out[0] if squeeze else out


def add(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """a + b for equally shaped fp32 CUDA tensors (the ``R += addition`` of the generators)."""
    a, b = _prep(a), _prep(b)
    out = torch.empty_like(a)
    cols = a.shape[-1]
    rows = a.numel() // cols
    import ctypes as C
    check(lib().mmx_add(ptr(a), cols, ptr(b), cols, C.c_float(1.0), ptr(out), cols, rows, cols, current_stream()))
    return out


def mm_update_batched(R_ss, R_qq, R_qs, cam_sq, apply_normalization=True, apply_self_in_rule_10=True, nan_to_zero=False):
    """Batched rules 10+11: R_ss [B,T,T], R_qq [B,S,S], R_qs [B,S,T] or None, cam_sq [B,T,S].
    Returns (R_sq_addition [B,T,S], R_ss_addition [B,T,T] or None, min_diag [2] device tensor for the reference's
    ``assert diag(R - I) >= 0``)."""
    Rs, Rq, Ab = _prep(R_ss), _prep(R_qq), _prep(cam_sq)
    B, T, S = Ab.shape
    flags = (MM_NORMALIZE if apply_normalization else 0) | (MM_SELF_IN_10 if apply_self_in_rule_10 else 0) | \
            (MM_NAN_TO_ZERO if nan_to_zero else 0)
    l = lib()
    ws = torch.empty(l.mmx_mm_update_workspace(B, T, S) // 4 + 4, device=Rs.device, dtype=torch.float32)
    sq_add = torch.empty(B, T, S, device=Rs.device, dtype=torch.float32)
    Rqs = ss_add = None
    if R_qs is not None:
        Rqs = _prep(R_qs)
        ss_add = torch.empty(B, T, T, device=Rs.device, dtype=torch.float32)
    md = torch.zeros(2, device=Rs.device, dtype=torch.float32)
    check(l.mmx_mm_update(ptr(Rs), T, ptr(Rq), S, ptr(Rqs), T, ptr(Ab), S, ptr(sq_add), S, ptr(ss_add), T, B, T, S, flags,
                          ptr(ws), ptr(md), current_stream()))
    return sq_add, ss_add, md


def head_mean_record(rec, batch: int, use_cam: bool = False) -> torch.Tensor:
    """mean over heads of the staged A (raw-attention / rollout baselines: ``cam.mean(dim=0)``,
    DETR/modules/ExplanationGenerator.py:229-230,247-249).  A >= 0, so it is rule 5 with a gradient of ones.
    ``use_cam``: the head mean of the LRP relevance of A (partial LRP, :215-217), which is signed:
    mean(x) = mean(relu(x)) - mean(relu(-x)), two rule-5 passes."""
    A, _, ld = rec.padded(use_cam)
    B, H, T = A.shape[0], A.shape[1], A.shape[2]
    assert B == batch
    ones = torch.ones_like(A)
    out = torch.empty(B, T, ld, device=A.device, dtype=torch.float32)
    check(lib().mmx_avg_heads(ptr(A), ptr(ones), ptr(out), B, H, T, ld, ld, ld, current_stream()))
    if use_cam:
        import ctypes as C
        neg = torch.empty_like(out)
        check(lib().mmx_avg_heads(ptr(A), ptr(-ones), ptr(neg), B, H, T, ld, ld, ld, current_stream()))
        check(lib().mmx_add(ptr(out), ld, ptr(neg), ld, C.c_float(-1.0), ptr(out), ld, B * T, ld, current_stream()))
    return out[..., :rec.S]


def gradcam_record(rec, batch: int) -> torch.Tensor:
    """attn-GradCAM of one attention module (``gradcam``, DETR/modules/ExplanationGenerator.py:275-280)."""
    A, dA, ld = rec.padded()
    if dA is None:
        raise MmxError("no attention gradient recorded (backward did not reach this attention)")
    B, H, T = A.shape[0], A.shape[1], A.shape[2]
    assert B == batch
    out = torch.empty(B, T, rec.S, device=A.device, dtype=torch.float32)
    gbar = torch.empty(B * H, device=A.device, dtype=torch.float32)
    check(lib().mmx_attn_gradcam(ptr(A), ptr(dA), ptr(out), ptr(gbar), B, H, T, rec.S, ld, rec.S, current_stream()))
    return out


def minmax_normalize(maps: torch.Tensor) -> torch.Tensor:
    """(x - x.min()) / (x.max() - x.min()) per leading-dim map (VisualBERT ExplanationGenerator.py:203,
    DETR/mask_generator.py:115-121)."""
    x = _prep(maps)
    out = torch.empty_like(x)
    B = x.shape[0]
    check(lib().mmx_minmax_normalize(ptr(x), ptr(out), B, x.numel() // max(B, 1), current_stream()))
    return out


def bmm(cam: torch.Tensor, R: torch.Tensor) -> torch.Tensor:
    """Batched ``torch.matmul(cam, R)`` ([B,T,S] x [B,S,Q]) on the rule GEMM kernel: the un-aggregated update of the
    no-aggregation ablations (``self.R_i_i = torch.matmul(cam, self.R_i_i)``, DETR/modules/ExplanationGenerator.py:323)."""
    A, Bm = _prep(cam), _prep(R)
    B, T, S = A.shape
    Q = Bm.shape[-1]
    out = torch.empty(B, T, Q, device=A.device, dtype=torch.float32)
    check(lib().mmx_bmm_add(ptr(A), S, T * S, 0, ptr(Bm), Q, S * Q, None, 0, 0, ptr(out), Q, T * Q, B, T, Q, S, current_stream()))
    return out


def otsu_masks(cams: torch.Tensor):
    """Min-max -> uint8 -> Otsu -> 0/255 for a batch of maps [B, ...] (DETR/mask_generator.py:115-121).  Returns
    (masks, same shape, float32; thresholds [B] int32)."""
    x = _prep(cams)
    B = x.shape[0]
    masks = torch.empty_like(x)
    thr = torch.empty(B, device=x.device, dtype=torch.int32)
    check(lib().mmx_otsu_masks(ptr(x), ptr(masks), ptr(thr), B, x.numel() // max(B, 1), current_stream()))
    return masks, thr
