"""CPU (gloo, world_size 2 and 3): the host-side sharding / all-gather logic of the N>1 path, with a stand-in
interpret function (the CUDA engine needs a GPU; the sharding logic does not)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_interpret(images, tokens, start_layer, start_layer_text):
    """Deterministic per-sample function of the inputs (sample-independent, like the real path)."""
    B, ctx = tokens.shape
    base = tokens.float().sum(-1)
    img = images.float().reshape(images.shape[0], -1).sum(-1)
    if img.shape[0] == 1:
        img = img.expand(B)
    rt = (base + img)[:, None, None] * torch.ones(B, ctx, ctx) + start_layer
    ri = (base - img)[:, None] * torch.ones(B, 5) + start_layer_text
    return rt, ri


def _worker(rank, world, port, B, repeat, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import mmx_b200  # noqa: F401
    from mmx_b200.distributed import interpret_sharded, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    images = torch.randn(1 if repeat else B, 3, 4, 4, generator=g)
    tokens = torch.randint(0, 100, (B, 6), generator=g)
    rt, ri = interpret_sharded(_fake_interpret, images, tokens, 2, 3)
    ref_t, ref_i = _fake_interpret(images, tokens, 2, 3)
    ok = torch.equal(rt, ref_t) and torch.equal(ri, ref_i)
    # the generic unit form used by the DETR / LXMERT / VisualBERT / ViT generators
    from mmx_b200.distributed import map_sharded
    units = torch.arange(B * 3, dtype=torch.float32).reshape(B, 3)
    got = map_sharded(lambda lo, hi: units[lo:hi] * 2 + 1, B)
    ok = ok and torch.equal(got, units * 2 + 1)
    lo, hi = shard_range(B, rank, world)
    q.put((rank, ok, lo, hi))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,B,repeat", [(2, 8, False), (2, 7, False), (3, 4, True), (2, 1, False)])
def test_sharded_equals_single(world, B, repeat):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, repeat, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    cover = sorted((lo, hi) for _, _, lo, hi in res)
    assert cover[0][0] == 0 and cover[-1][1] == B
    for (a, b), (c, d) in zip(cover, cover[1:]):
        assert b == c


def test_shard_range_partition():
    import mmx_b200  # noqa: F401
    from mmx_b200.distributed import shard_range
    for n in (0, 1, 7, 64, 2048):
        for w in (1, 2, 3, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
