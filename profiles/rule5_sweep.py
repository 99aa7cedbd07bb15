"""Rule-5 (avg_heads) standalone bandwidth for one (unroll, CTAs/SM) setting from the environment
(MMX_AVG_UNROLL, MMX_AVG_CTAS_PER_SM), at the CLIP ViT-B/32 batch-64 all-layer sizes of both towers."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmx_b200
from mmx_b200._lib import lib, ptr, current_stream
l = lib()
res = {"unroll": os.environ.get("MMX_AVG_UNROLL", "4"), "ctas_per_sm": os.environ.get("MMX_AVG_CTAS_PER_SM", "8")}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, (LB, H, S) in {"text": (12 * 64, 8, 77), "vision": (12 * 64, 12, 50)}.items():
    ld = (S + 3) // 4 * 4
    A = torch.rand(LB, H, S, ld, device="cuda"); G = torch.randn(LB, H, S, ld, device="cuda")
    out = torch.empty(LB, S, ld, device="cuda")
    for _ in range(3):
        l.mmx_avg_heads(ptr(A), ptr(G), ptr(out), LB, H, S, ld, ld, ld, current_stream())
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); l.mmx_avg_heads(ptr(A), ptr(G), ptr(out), LB, H, S, ld, ld, ld, current_stream()); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    byts = (2 * A.numel() + out.numel()) * 4
    res[name] = round(byts / (ts[len(ts) // 2] * 1e-3) / 1e9, 1)
print(json.dumps(res))
