"""Run under torchrun on N GPUs (not collected by pytest):  sharded interpret + one NCCL all-gather must equal the
single-GPU result BITWISE on every rank.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmx_b200  # noqa: E402
from mmx_b200.distributed import interpret_sharded  # noqa: E402
from oracle import clip_oracle as co  # noqa: E402  (synthetic weights / inputs only)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    cfg = co.VIT_B32
    sd = co.init_state_dict(cfg, seed=0)
    B = 22                                               # ragged over 4 or 8 ranks on purpose
    images, tokens = co.synthetic_inputs(cfg, B, seed=5)
    eng = mmx_b200.ClipEngine(mmx_b200.ClipConfig(*cfg.ref_args()), sd, max_batch=B, device=f"cuda:{local}")
    ic, tc = images.cuda(), tokens.cuda()
    for sl in (0, -1):
        rt, ri = interpret_sharded(eng.interpret, ic, tc, sl, sl)
        ft, fi = eng.interpret(ic, tc, sl, sl)           # the whole batch on this GPU
        ok = torch.equal(rt, ft) and torch.equal(ri, fi)
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"start_layer={sl}: sharded x{world} == single GPU bitwise: {bool(flag.item())}", flush=True)
        assert flag.item() == 1
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
