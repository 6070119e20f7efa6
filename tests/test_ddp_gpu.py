"""Multi-GPU data-parallel path (needs >= 2 GPUs; skipped otherwise): two ranks, NCCL all-reduce of the flat gradient
buffer, replicas stay bit-identical, and the update equals a single-process emulation (per-shard forward/backward with
per-shard BatchNorm statistics = the reference's DataParallel semantics, gradients averaged, one Adam step)."""
import os
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
NS = types.SimpleNamespace


def _cfg(f, s):
    return NS(MODEL=NS(EXTRA=NS(NUM_FEATURES=f, NUM_STACKS=s, NUM_BLOCKS=1), NUM_JOINTS=16))


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import fpd_b200  # noqa: F401
    from fpd_b200 import ops, parallel
    from fpd_b200.lib.models import hourglass as H
    from fpd_b200.train_step import FPDTrainStep
    from bench import synthetic_batch
    torch.manual_seed(7 + rank)                      # ranks start different; broadcast makes them identical
    student = H.get_pose_net(_cfg(64, 2), True).cuda()
    teacher = H.get_pose_net(_cfg(64, 1), False).cuda()
    parallel.broadcast_module(student, 0)
    parallel.broadcast_module(teacher, 0)
    init = {k: v.clone() for k, v in student.state_dict().items()}
    step = FPDTrainStep(student, teacher, alpha=0.5, lr=1e-3, use_graph=False)
    shards = [tuple(t.cuda() for t in synthetic_batch(2, 100 + r, 128, 128)) for r in range(world)]
    step.step(*shards[rank])
    torch.cuda.synchronize()
    flat = step.flat.flat.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok_emul, err = True, 0.0
    if rank == 0:
        # single-process emulation on this GPU
        ref = H.get_pose_net(_cfg(64, 2), True).cuda()
        ref.load_state_dict(init)
        rstep = FPDTrainStep(ref, teacher, alpha=0.5, lr=1e-3, use_graph=False)
        acc = torch.zeros_like(rstep.flat.grad)
        for r in range(world):
            ref.load_state_dict({k: v for k, v in init.items() if "running" in k or "num_batches" in k}, strict=False)
            rstep._body(shards[r][0], shards[r][1], shards[r][2].reshape(2, -1))
            acc += rstep.flat.grad
        ops.adam_flat(rstep.flat.flat, acc, rstep.exp_avg, rstep.exp_avg_sq, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1,
                      grad_scale=1.0 / world)
        torch.cuda.synchronize()
        err = ((rstep.flat.flat - flat).abs().max() / flat.abs().max()).item()
        ok_emul = err < 1e-5
    q.put((rank, same, ok_emul, err))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_step_matches_single_process_emulation():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, same, ok_emul, err in res:
        assert same, "replicas diverged"
        assert ok_emul, "DDP update differs from the single-process emulation: %.3e" % err
