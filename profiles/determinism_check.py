import sys, os, torch
sys.path.insert(0, os.getcwd())
import mmx_b200
from oracle import clip_oracle as co
cfg = co.VIT_B32
sd = co.init_state_dict(cfg, seed=0)
eng = mmx_b200.ClipEngine(mmx_b200.ClipConfig(*cfg.ref_args()), sd, max_batch=64, device="cuda:0")
images, tokens = co.synthetic_inputs(cfg, 64, seed=33)
ic, tc = images.cuda(), tokens.cuda()
rt, ri = eng.interpret(ic, tc, 0, 0)
for i in range(3):
    rt2, ri2 = eng.interpret(ic, tc, 0, 0)
    print("repeat", i, torch.equal(rt, rt2), torch.equal(ri, ri2), (rt - rt2).abs().max().item(), (ri - ri2).abs().max().item())
for b in (0, 17, 63):
    rt1, ri1 = eng.interpret(ic[b:b+1], tc[b:b+1], 0, 0)
    print("alone", b, torch.equal(rt1[0], rt[b]), torch.equal(ri1[0], ri[b]), (rt1[0]-rt[b]).abs().max().item(), (ri1[0]-ri[b]).abs().max().item())
