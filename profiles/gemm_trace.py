"""Timeline of one CTA of the tcgen05 GEMM (clock64 stamps per role and K-slab).
Needs a library built with tracing: MMX_EXTRA_NVCC_FLAGS=-DMMX_TC_TRACE python transformer-mm-explainability_b200/build.py --force
usage: python profiles/gemm_trace.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmx_b200
from mmx_b200._lib import lib, check, ptr, current_stream
l = lib()
M, N, K = 3200, 2304, 768
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda"); Cm = torch.empty(M, N, device="cuda")
bias = torch.randn(N, device="cuda")
for _ in range(3):
    check(l.mmx_linear(ptr(A), K, ptr(W), K, ptr(bias), None, 0, ptr(Cm), N, None, 0, M, N, K, current_stream()))
buf = torch.zeros(4, 256, 4, dtype=torch.int64, device="cuda")
l.mmx_gemm_trace(ptr(buf))
check(l.mmx_linear(ptr(A), K, ptr(W), K, ptr(bias), None, 0, ptr(Cm), N, None, 0, M, N, K, current_stream()))
torch.cuda.synchronize()
l.mmx_gemm_trace(None)
t = buf.cpu()
t0 = int(t[t > 0].min())
rel = lambda x: int(x) - t0 if int(x) > 0 else -1
print("slab | prod: empty_ok tma_issued | split: full_ok arrived | mma: split_ok issued committed")
for i in range(0, 60):
    print(f"{i:4d} | {rel(t[0,i,0]):7d} {rel(t[0,i,1]):7d} | {rel(t[2,i,0]):7d} {rel(t[2,i,1]):7d} | {rel(t[1,i,0]):7d} {rel(t[1,i,1]):7d} {rel(t[1,i,2]):7d}")
print("tile | epi: tfull_ok tmem_released stores_done")
for i in range(4):
    print(f"{i:4d} | {rel(t[3,i,0]):7d} {rel(t[3,i,1]):7d} {rel(t[3,i,2]):7d}")
