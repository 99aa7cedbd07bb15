"""Stage-by-stage parity of the CUDA engine vs the CPU oracle for CLIP ViT-B/32 (run on the GPU box).
Prints max|x - ref| / max|ref| for logits, A_l, dA_l, Abar_l and the final maps, for both GEMM backends."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import mmx_b200  # noqa: E402
from oracle import clip_oracle as co  # noqa: E402
from util import rel_err, text_rel_err  # noqa: E402


def main():
    cfg = co.VIT_B32
    sd = co.init_state_dict(cfg, seed=0)
    B = 3
    images, tokens = co.synthetic_inputs(cfg, B, seed=21)
    eng = mmx_b200.ClipEngine(mmx_b200.ClipConfig(*cfg.ref_args()), sd, max_batch=B, device="cuda:0")
    dtype = torch.float64 if "--fp64" in sys.argv else torch.float32
    ot, oi, stg = co.clip_interpret(sd, cfg, images, tokens, 0, 0, return_stages=True, dtype=dtype)
    for backend in (0, 1, 2):
        if mmx_b200.lib().mmx_set_gemm_backend(backend) != backend:
            continue
        rt, ri = mmx_b200.interpret(images.cuda(), tokens.cuda(), eng, "cuda:0", 0, 0)
        print(f"== backend {backend} (oracle {dtype})")
        print("logits", rel_err(eng.tap("logits"), stg["logits"]))
        for tower, (Ak, Gk, Bk, L, H) in enumerate((("A_v", "G_v", "bar_v", 12, 12), ("A_t", "G_t", "bar_t", 12, 8))):
            for l in (11, 8, 4, 0):
                S = stg[Ak][l].shape[-1]
                print(f" tower {tower} layer {l:2d}: A {rel_err(eng.tap('A', tower, l), stg[Ak][l].reshape(B, H, S, S)):.2e}"
                      f"  dA {rel_err(eng.tap('dA', tower, l), stg[Gk][l].reshape(B, H, S, S)):.2e}"
                      f"  Abar {rel_err(eng.tap('Abar', tower, l), stg[Bk][l]):.2e}")
        print(f" R_text {text_rel_err(rt, ot):.2e}  R_image {rel_err(ri, oi):.2e}")
    mmx_b200.lib().mmx_set_gemm_backend(1)


if __name__ == "__main__":
    main()
