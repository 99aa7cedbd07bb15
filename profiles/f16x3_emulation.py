"""CPU emulation of the fp16x3 tensor-core GEMM numerics (round 2 design study; no GPU needed).

Every linear of the CLIP oracle (forward AND dgrad) is replaced by
    C = A_hi B_hi^T + (A_lo' B_hi^T + A_hi B_lo'^T) / 2048,   x_hi = fp16(x), x_lo' = fp16((x - x_hi) * 2048)
with exact fp16 x fp16 products summed in float64 and rounded to fp32 once (the tensor core's truncating fp32
accumulation is NOT modelled: it is identical for the tf32x3 kernel this replaces).  The script reports the end-to-end
error of the relevancy maps against the fp64 oracle and the dynamic range of every GEMM A operand, which is what
decides whether fp16's exponent range needs a scale.

    python profiles/f16x3_emulation.py [--full] [--grad-scale 1]
"""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmx_b200  # noqa: F401,E402  (package alias for the hyphenated directory)
from oracle import clip_oracle as co  # noqa: E402

RANGES = []
GRAD_SCALE = 1.0


def split(x):
    hi = x.to(torch.float16)
    lo = ((x - hi.float()) * 2048.0).to(torch.float16)
    return hi.double(), lo.double()


def emul_mm(a, bt, tag):
    """a [M,K] fp32, bt [N,K] fp32 -> [M,N] fp32 through the fp16 hi/lo planes."""
    nz = a[a != 0].abs()
    if nz.numel():
        RANGES.append((tag, float(nz.min()), float(nz.median()), float(nz.max())))
    ah, al = split(a)
    bh, bl = split(bt)
    main = ah @ bh.t()
    cross = al @ bh.t() + ah @ bl.t()
    return (main.float() + (cross / 2048.0).float()).float()


class EmulLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(w)
        ctx.shape = x.shape
        y = emul_mm(x.reshape(-1, x.shape[-1]).float(), w.float(), "fwd")
        if b is not None:
            y = y + b
        return y.reshape(*x.shape[:-1], w.shape[0]).to(x.dtype)

    @staticmethod
    def backward(ctx, gy):
        (w,) = ctx.saved_tensors
        g = gy.reshape(-1, gy.shape[-1]).float() * GRAD_SCALE
        gx = emul_mm(g, w.t().contiguous().float(), "bwd") / GRAD_SCALE
        return gx.reshape(ctx.shape).to(gy.dtype), None, None


def emul_linear(x, w, b=None):
    return EmulLinear.apply(x, w, b)


def main():
    global GRAD_SCALE
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="ViT-B/32 (default: a 4-layer, width-256 model)")
    ap.add_argument("--grad-scale", type=float, default=1.0)
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    GRAD_SCALE = args.grad_scale
    cfg = co.VIT_B32 if args.full else co.ClipConfig(128, 64, 4, 256, 16, 24, 1024, 128, 2, 4)
    sd = co.init_state_dict(cfg, seed=0)
    images, tokens = co.synthetic_inputs(cfg, args.batch, seed=7)
    ref_t, ref_i = co.clip_interpret(sd, cfg, images, tokens, 0, 0, dtype=torch.float64)
    f32_t, f32_i = co.clip_interpret(sd, cfg, images, tokens, 0, 0)
    orig = F.linear
    co.F.linear = emul_linear
    try:
        em_t, em_i = co.clip_interpret(sd, cfg, images, tokens, 0, 0)
    finally:
        co.F.linear = orig
    eye = torch.eye(cfg.context_length, dtype=torch.float64)

    def err(t, i):
        return (float((t.double() - ref_t).abs().max() / (ref_t - eye).abs().max()),
                float((i.double() - ref_i).abs().max() / ref_i.abs().max()))
    print("fp32 autograd vs fp64: text %.2e image %.2e" % err(f32_t, f32_i))
    print("fp16x3 emulation vs fp64: text %.2e image %.2e" % err(em_t, em_i))
    for kind in ("fwd", "bwd"):
        rs = [r for r in RANGES if r[0] == kind]
        print(f"{kind}: {len(rs)} GEMM A operands; min|x| {min(r[1] for r in rs):.2e}, median of medians "
              f"{sorted(r[2] for r in rs)[len(rs) // 2]:.2e}, smallest median {min(r[2] for r in rs):.2e}, "
              f"max|x| {max(r[3] for r in rs):.2e}")


if __name__ == "__main__":
    main()
