"""CPU (gloo, world_size 2): the data-parallel gradient synchronisation used at N > 1 GPUs.

GradSync is backend-agnostic (SUM + scale), so the bucketing / ordering / averaging logic is exercised here with
gloo; on the GPU box the same code runs over NCCL (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, bucket_bytes, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from zeroshotvideoclassification_b200 import dist as zdist
    r, lr, w = zdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    sync = zdist.GradSync(bucket_bytes=bucket_bytes)
    g = torch.Generator().manual_seed(1234)            # same "true" gradients on both ranks ...
    shapes = {"layer4.w": (64, 32, 3), "layer4.gamma": (64,), "layer3.w": (16, 8, 3, 3), "stem.w": (5, 3, 7)}
    base = {k: torch.randn(*s, generator=g) for k, s in shapes.items()}
    local = {k: v * (rank + 1) for k, v in base.items()}          # ... scaled per rank: mean factor is 1.5
    # submitted in backward order, in two groups like two residual blocks
    sync.submit({k: local[k].clone() for k in ("layer4.w", "layer4.gamma")})
    sync.submit({k: local[k].clone() for k in ("layer3.w", "stem.w")})
    out = sync.finish()
    ok = set(out) == set(shapes)
    for k in shapes:
        ok &= bool(torch.allclose(out[k], base[k] * 1.5, rtol=1e-6, atol=1e-7)) and out[k].shape == base[k].shape
    # head-gradient helper and parameter broadcast
    lin = torch.nn.Linear(4, 3)
    with torch.no_grad():
        for p in lin.parameters():
            p.fill_(float(rank + 1))
    zdist.broadcast_module(lin, 0)
    ok &= all(bool((p == 1.0).all()) for p in lin.parameters())
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(rank))
    zdist.sync_head_grads(list(lin.parameters()))
    ok &= all(bool(torch.allclose(p.grad, torch.full_like(p, 0.5))) for p in lin.parameters())
    nbytes_dict = sync.bytes_reduced
    # flat arena: slices handed over block by block are averaged IN PLACE, merged until the bucket threshold is met
    arena = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    before = sync.bytes_reduced
    sync.submit_range(arena, 0, 300)
    sync.submit_range(arena, 300, 640)      # extends the pending slice
    sync.submit_range(arena, 700, 1000)     # a gap: the pending slice is flushed first
    sync.finish()
    expect = torch.arange(1000, dtype=torch.float32) * 1.5
    ok &= bool(torch.allclose(arena[:640], expect[:640])) and bool(torch.allclose(arena[700:], expect[700:]))
    ok &= bool(torch.equal(arena[640:700], torch.arange(640, 700, dtype=torch.float32) * (rank + 1)))   # untouched gap
    ok &= sync.bytes_reduced - before == 4 * (640 + 300)
    # parameters accumulated by autograd: reduced in place from a post-accumulate hook during backward
    lin2 = torch.nn.Linear(3, 2)
    with torch.no_grad():
        for p in lin2.parameters():
            p.fill_(1.0)
    sync.attach(list(lin2.parameters()))
    lin2(torch.full((1, 3), float(rank + 1))).sum().backward()
    sync.finish()
    ok &= bool(torch.allclose(lin2.weight.grad, torch.full((2, 3), 1.5))) and bool(torch.allclose(lin2.bias.grad, torch.ones(2)))
    sync.detach()
    q.put((rank, ok, nbytes_dict))
    dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [1, 1 << 20])
def test_gradsync_world2_gloo(bucket_bytes):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bucket_bytes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    nbytes = 4 * (64 * 32 * 3 + 64 + 16 * 8 * 3 * 3 + 5 * 3 * 7)
    assert all(b == nbytes for _, _, b in res)


def test_gradsync_single_process_is_a_noop():
    from zeroshotvideoclassification_b200 import dist as zdist
    sync = zdist.GradSync()
    sync.submit({"w": torch.ones(3)})
    assert sync.finish() == {} and sync.world == 1


def _acc_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from zeroshotvideoclassification_b200 import dist as zdist
    zdist.init_from_env("gloo")
    n = 10001
    g = torch.Generator().manual_seed(7)
    hit1 = torch.rand(n, generator=g) < 0.3          # the same synthetic per-row outcomes on every rank
    hit5 = hit1 | (torch.rand(n, generator=g) < 0.4)
    lo, hi = zdist.shard_rows(n, rank, world)
    counts = torch.tensor([int(hit1[lo:hi].sum()), int(hit5[lo:hi].sum()), hi - lo], dtype=torch.int64)
    top1, top5 = zdist.reduce_accuracy_counts(counts)
    ok = abs(top1 - 100.0 * float(hit1.float().mean())) < 1e-4 and abs(top5 - 100.0 * float(hit5.float().mean())) < 1e-4
    ok &= counts[2].item() == hi - lo            # the caller's tensor is not modified by the reduction
    q.put((rank, ok, (lo, hi)))
    dist.destroy_process_group()


def test_sharded_accuracy_world2_gloo():
    """Evaluation shards rows over ranks: the row ranges tile [0, N) and the all-reduced hit
    counts give the global top-1 / top-5 on every rank."""
    from zeroshotvideoclassification_b200 import dist as zdist
    for n, w in ((10, 3), (7, 8), (10000, 8), (0, 2)):
        rs = [zdist.shard_rows(n, r, w) for r in range(w)]
        assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
        assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_acc_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    with pytest.raises(ValueError):
        zdist.reduce_accuracy_counts(torch.zeros(3, dtype=torch.int64))
