"""CPU suite: the N>1 host logic (flat parameter/gradient buffers, one all-reduce per step, sharding) under
world_size=2 with the gloo backend."""
import os
import types

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

NS = types.SimpleNamespace


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import fpd_b200  # noqa: F401
    from fpd_b200 import parallel
    from fpd_b200.train_step import FlatParams
    from fpd_b200.lib.models import hourglass as H
    cfg = NS(MODEL=NS(EXTRA=NS(NUM_FEATURES=64, NUM_STACKS=1, NUM_BLOCKS=1), NUM_JOINTS=16))
    torch.manual_seed(100 + rank)            # ranks start different ...
    net = H.get_pose_net(cfg, True)
    parallel.broadcast_module(net, 0)        # ... and are made identical once
    flat = FlatParams(net)
    assert flat.is_intact()
    # parameters are views of the flat buffer: an update of the buffer is an update of the module
    before = net.conv1.weight.clone()
    flat.flat.add_(1.0)
    assert torch.allclose(net.conv1.weight, before + 1.0)
    flat.flat.sub_(1.0)
    # rank-dependent gradients -> one all-reduce -> mean
    for i, gv in enumerate(flat.grad_views):
        gv.fill_(float(rank + 1) * (i % 7 + 1))
    scale = parallel.allreduce_mean_(flat.grad)
    expect = sum(r + 1 for r in range(world)) / world
    ok = all(torch.allclose(gv * scale, torch.full_like(gv, expect * (i % 7 + 1))) for i, gv in enumerate(flat.grad_views))
    checksum = float(flat.flat.double().sum())
    sums = [None] * world
    dist.all_gather_object(sums, checksum)
    q.put((rank, ok, sums, parallel.shard_range(65, rank, world)))
    dist.destroy_process_group()


def test_flat_buffers_and_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, sums, shard in res:
        assert ok
        assert abs(sums[0] - sums[1]) < 1e-9          # identical replicas after the broadcast
    assert res[0][3] == (0, 33) and res[1][3] == (33, 65)  # shards tile the global batch


def test_shard_range_covers_batch():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import fpd_b200  # noqa: F401
    from fpd_b200.parallel import shard_range
    for gb in (1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


def _sync_math_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import fpd_b200  # noqa: F401
    from fpd_b200 import parallel
    g = torch.Generator().manual_seed(5)
    full = torch.randn(world * 6, 37, 8, generator=g) * 3 + 100.0          # [N, pixels, C], large common offset
    mine = full[rank * 6:(rank + 1) * 6].reshape(-1, 8)
    mean, var = parallel.merge_bn_stats(mine.mean(0), mine.var(0, unbiased=False))
    ref = full.reshape(-1, 8).double()
    ok = torch.allclose(mean.double(), ref.mean(0), rtol=1e-6) and torch.allclose(var.double(), ref.var(0, unbiased=False),
                                                                                  rtol=1e-5)
    s = torch.arange(16, dtype=torch.float32) * (rank + 1)
    avg = parallel.allreduce_avg(s)
    ok = ok and torch.allclose(avg, torch.arange(16, dtype=torch.float32) * (sum(range(1, world + 1)) / world))
    ok = ok and torch.equal(s, torch.arange(16, dtype=torch.float32) * (rank + 1))      # input left untouched
    # lib.core.function's DDP path: one flat all-reduce, mean over ranks
    from fpd_b200.lib.core.function import _allreduce_grads_
    grads = [torch.full((3, 2), float(rank + 1)), torch.full((5,), 10.0 * (rank + 1))]
    out = _allreduce_grads_(object(), grads)
    ok = ok and torch.allclose(out[0], torch.full((3, 2), 1.5)) and torch.allclose(out[1], torch.full((5,), 15.0))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_syncbn_merge_and_grad_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sync_math_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)
