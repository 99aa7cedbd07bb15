"""clock64 phase timeline of block (0,0,0) of the attention forward kernel (mmx_gemm_trace) on the CLIP ViT-B/32 shape
(B=64, H=12, T=S=50, packed q|k|v rows).  usage (GPU box): python profiles/attn_trace.py [B H S hd]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmx_b200  # noqa: E402,F401
from mmx_b200._lib import lib, check, ptr, current_stream  # noqa: E402


def main():
    B, H, S, hd = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (64, 12, 50, 64)
    l = lib()
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(B * S, 3 * D, device="cuda", generator=g)
    ld = (S + 3) // 4 * 4
    A = torch.empty(B, H, S, ld, device="cuda")
    O = torch.empty(B * S, D, device="cuda")
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    run = lambda: check(l.mmx_attention_fwd(ptr(q), 3 * D, ptr(k), 3 * D, ptr(v), 3 * D, None, ptr(A), ld, ptr(O), D, B, H, S, S, hd,
                                            hd ** -0.5, 0, current_stream()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    buf = torch.zeros(4112, dtype=torch.int64, device="cuda")
    check(l.mmx_gemm_trace(ptr(buf)))
    run()
    torch.cuda.synchronize()
    check(l.mmx_gemm_trace(None))
    t = buf.cpu()[4096:4103].tolist()
    names = ["start", "loads issued", "scores done", "V issued + softmax start", "softmax done", "after sync", "PV done", "O stored"]
    print(f"B={B} H={H} S={S} hd={hd}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
    print("phase boundaries of block 0 (clk since start):", [int(x - t[0]) for x in t])


if __name__ == "__main__":
    main()
