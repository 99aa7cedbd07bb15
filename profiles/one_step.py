"""One call of a workload between cudaProfilerStart / cudaProfilerStop, for short ncu runs:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x.csv \
        python profiles/one_step.py detr_r50 --batch 4
    python profiles/summarize_launches.py gpurun_out/x.csv > profiles/x.md

workloads: clip_b32 | clip_l14_336 | detr_r50 | lxmert   (per-launch times under ncu are cold-cache and serialised: compare
SHARES, not absolutes)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmx_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["clip_b32", "clip_l14_336", "detr_r50", "lxmert"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--lrp", action="store_true")
    ap.add_argument("--warm", type=int, default=2)
    a = ap.parse_args()
    dev = "cuda:0"
    if a.workload.startswith("clip"):
        cfg = mmx_b200.VIT_B32 if a.workload == "clip_b32" else mmx_b200.VIT_L14_336
        B = a.batch or (64 if a.workload == "clip_b32" else 8)
        eng = mmx_b200.ClipEngine(cfg, mmx_b200.clip_init_state_dict(cfg, seed=0), max_batch=B, device=dev)
        images, tokens = mmx_b200.clip_synthetic_inputs(cfg, B, seed=1234, length_seed=4321)
        images, tokens = images.to(dev), tokens.to(dev)
        call = lambda: eng.interpret(images, tokens, 0, 0, validate=False)
    elif a.workload == "detr_r50":
        cfg = mmx_b200.DETR_R50
        B = a.batch or 16
        eng = mmx_b200.DetrEngine(mmx_b200.detr_init_state_dict(cfg, seed=9), nhead=cfg.nhead, device=dev)
        src, pos, tq = (t.to(dev) for t in mmx_b200.detr_synthetic_inputs(cfg, B, 25, 25, seed=4))
        gen = mmx_b200.Generator(eng)
        call = lambda: gen.generate_ours((src, pos), tq, use_lrp=a.lrp)
    else:
        cfg = mmx_b200.LXMERT_BASE
        B = a.batch or 16
        eng = mmx_b200.LxmertEngine(mmx_b200.lxmert_init_state_dict(cfg, seed=1), num_heads=cfg.heads, device=dev)
        item = tuple(t.to(dev) for t in mmx_b200.lxmert_synthetic_inputs(cfg, B, 20, 36, seed=8))
        gen = mmx_b200.GeneratorOurs(eng)
        call = lambda: gen.generate_ours(item, use_lrp=a.lrp)
    for _ in range(a.warm):
        call()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    call()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
