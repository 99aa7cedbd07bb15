"""Extra data point: CLIP ViT-L/14@336 (BASELINE.json config 5 model) relevancy maps/s on one GPU, all layers.
usage: python profiles/clip_l14_throughput.py [batch]"""
import os, sys, time, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmx_b200
from oracle import clip_oracle as co   # synthetic weights / inputs only

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = co.VIT_L14_336
sd = co.init_state_dict(cfg, seed=0)
eng = mmx_b200.ClipEngine(mmx_b200.ClipConfig(*cfg.ref_args()), sd, max_batch=B, device="cuda:0")
images, tokens = co.synthetic_inputs(cfg, B, seed=1)
ic, tc = images.cuda(), tokens.cuda()
out = {}
for sl in (0, -1):
    for _ in range(2):
        eng.interpret(ic, tc, sl, sl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        eng.interpret(ic, tc, sl, sl)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out[f"start_layer={sl}"] = {"ms_per_step": ms, "maps_per_s": B / (ms * 1e-3)}
out["batch"] = B
out["mem_GB"] = torch.cuda.max_memory_allocated() / 2**30
print(json.dumps(out))
