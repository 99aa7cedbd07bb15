"""Micro-benchmark of the staged-attention kernels through the C ABI at the CLIP ViT-B/32 batch-64 shapes.
usage (GPU box): python profiles/attn_bench.py            (CUDA-event timing, L2 not flushed: 40-60 MB working set)
                 ncu --set full -k regex:attention -c 3 python profiles/attn_bench.py --once"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmx_b200  # noqa: E402
from mmx_b200._lib import lib, check, ptr, current_stream  # noqa: E402

SHAPES = [("vision", 64, 12, 50, 64, 0), ("text-dense", 64, 8, 77, 64, 1)]   # name, B, H, T, hd, flags (1 = causal)


def main():
    l = lib()
    once = "--once" in sys.argv
    out = []
    for name, B, H, T, hd, flags in SHAPES[:1] if once else SHAPES:
        D = H * hd
        g = torch.Generator(device="cuda").manual_seed(1)
        qkv = torch.randn(B * T, 3 * D, device="cuda", generator=g)
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
        ld = (T + 3) // 4 * 4
        A = torch.empty(B, H, T, ld, device="cuda"); dA = torch.empty_like(A)
        O = torch.empty(B * T, D, device="cuda"); dO = torch.randn(B * T, D, device="cuda", generator=g)
        dqkv = torch.empty_like(qkv); delta = torch.empty(B, H, T, device="cuda")
        scale = hd ** -0.5

        def fwd():
            check(l.mmx_attention_fwd(ptr(q), 3 * D, ptr(k), 3 * D, ptr(v), 3 * D, None, ptr(A), ld, ptr(O), D, B, H, T, T, hd,
                                      C.c_float(scale), flags, current_stream()))

        def bwd():
            check(l.mmx_attention_bwd(ptr(dO), D, ptr(q), 3 * D, ptr(k), 3 * D, ptr(v), 3 * D, ptr(A), ptr(dA), ld, ptr(delta),
                                      ptr(dqkv), 3 * D, ptr(dqkv[:, D:]), 3 * D, ptr(dqkv[:, 2 * D:]), 3 * D, B, H, T, T, hd,
                                      C.c_float(scale), flags, current_stream()))
        for _ in range(20 if not once else 1):       # also ramps the clocks before the first timed loop
            fwd(); bwd()
        torch.cuda.synchronize()
        if once:
            continue
        for fn, label, nbytes in ((fwd, "fwd", 4 * (qkv.numel() + A.numel() + O.numel())),
                                  (bwd, "bwd (q + kv kernels)", 4 * (2 * qkv.numel() + 3 * A.numel() + O.numel()))):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 50
            e0.record()
            for _ in range(reps):
                fn()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            out.append(dict(shape=name, kernel=label, us=round(us, 2), alg_MB=round(nbytes / 1e6, 1), GBps=round(nbytes / us / 1e3, 1)))
            print(out[-1], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
