"""Run-to-run bit reproducibility and per-sample independence of ClipEngine.interpret on one GPU (the properties
`sharded_equals_single` rests on).  usage (GPU box): [MMX_PDL=0] python profiles/determinism_check.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.getcwd())
import mmx_b200

cfg = mmx_b200.VIT_B32
eng = mmx_b200.ClipEngine(cfg, mmx_b200.clip_init_state_dict(cfg, seed=0), max_batch=64, device="cuda:0")
images, tokens = mmx_b200.clip_synthetic_inputs(cfg, 64, seed=33, length_seed=4321)
ic, tc = images.cuda(), tokens.cuda()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rt, ri = eng.interpret(ic, tc, 0, 0)
bad = 0
worst = 0.0
for i in range(reps):
    rt2, ri2 = eng.interpret(ic, tc, 0, 0)
    same = torch.equal(rt, rt2) and torch.equal(ri, ri2)
    if not same:
        bad += 1
        worst = max(worst, (rt - rt2).abs().max().item(), (ri - ri2).abs().max().item())
print(f"PDL={os.environ.get('MMX_PDL', '1')}: {reps} repeats of the same batch: {bad} differ bitwise (max abs diff {worst:.3e})")
g = torch.cat([ic, ic.flip(0)]), torch.cat([tc, tc.flip(0)])          # the same samples at other batch positions / other batch
rt3, ri3 = eng.interpret(g[0][32:96], g[1][32:96], 0, 0)
idx = list(range(32, 64)) + list(range(63, 31, -1))
print("same samples inside another batch:", torch.equal(rt3, rt[idx]), torch.equal(ri3, ri[idx]))
for b in (0, 17, 63):
    rt1, ri1 = eng.interpret(ic[b:b + 1], tc[b:b + 1], 0, 0)
    print("alone", b, torch.equal(rt1[0], rt[b]), torch.equal(ri1[0], ri[b]), (rt1[0] - rt[b]).abs().max().item(), (ri1[0] - ri[b]).abs().max().item())
