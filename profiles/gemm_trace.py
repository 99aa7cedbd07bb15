"""clock64 timeline of CTA 0 of one fp16x3 GEMM launch (mmx_gemm_trace): per tile, when the producer issued its first / last
TMA, when the MMA warp started the tile / saw the first slab / issued the last slab, when the epilogue saw tmem_full / handed
TMEM back / finished its stores.  usage (GPU box): [MMX_F16X3_MODE=..] python profiles/gemm_trace.py M N K"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmx_b200  # noqa: E402,F401
from mmx_b200._lib import lib, check, ptr, current_stream  # noqa: E402


def main():
    M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (3200, 2304, 768)
    l = lib()
    l.mmx_set_gemm_backend(2)
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    Cm = torch.empty(M, N, device="cuda")
    pk = torch.empty(l.mmx_pack_weight_bytes(N, K), dtype=torch.uint8, device="cuda")
    check(l.mmx_pack_weight(ptr(W), K, N, K, ptr(pk), current_stream()))
    run = lambda: check(l.mmx_linear_packed(ptr(A), K, ptr(W), K, ptr(pk), ptr(bias), None, 0, ptr(Cm), N, None, 0, M, N, K, current_stream()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    buf = torch.zeros(4112, dtype=torch.int64, device="cuda")
    check(l.mmx_gemm_trace(ptr(buf)))
    run()
    torch.cuda.synchronize()
    check(l.mmx_gemm_trace(None))
    t = buf.cpu()[:1024].view(4, 64, 4)
    if not (t > 0).any():
        print("no trace events (this kernel variant is not instrumented)")
        return
    t0 = int(t[t > 0].min())
    names = {0: ("tma first", "tma last"), 1: ("mma tile start", "mma first slab", "mma last slab"), 2: ("epi tmem_full", "epi tmem released", "epi stores done")}
    print(f"M={M} N={N} K={K} mode={os.environ.get('MMX_F16X3_MODE', 'default')}  (clk since the first event, CTA 0)")
    for it in range(64):
        if not (t[:, it] > 0).any():
            break
        row = []
        for role, labels in names.items():
            for slot, lab in enumerate(labels):
                v = int(t[role, it, slot])
                row.append(f"{lab} {v - t0 if v else -1}")
        print(f"tile {it}: " + " | ".join(row))


if __name__ == "__main__":
    main()
