"""Data-parallel plumbing: one process per GPU, replicas kept identical by ONE all-reduce of the flat gradient
buffer per step (NCCL over NVLink/NVSwitch on the GPU box; gloo in the CPU tests).

Replaces the reference's single-process nn.DataParallel (tools/fpd_train.py:143,173), which every iteration
scatters the batch, re-broadcasts all parameters (13 MB student + 102 MB teacher), gathers all outputs to GPU 0 and
reduces gradients to GPU 0. Here weights stay resident on every rank; the teacher needs no
communication; BatchNorm statistics stay per-rank (= the reference's per-replica semantics)."""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world):
    """Contiguous per-image shard [lo, hi) of a global batch (per-image data parallelism, no data-path collective).
    NB: allreduce_mean_ averages the per-rank gradients with equal weights, which equals the global-batch mean only for
    EQUAL shards (the losses are per-rank batch means) -- keep global_batch a multiple of world for training."""
    per = global_batch // world
    rem = global_batch % world
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def broadcast_module(module, src=0, group=None):
    """Make every rank's parameters and buffers identical once, at start-up."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src, group=group)


def allreduce_mean_(flat_grad, group=None):
    """Sum the flat gradient buffer over ranks in place; returns the scale (1/world) the optimizer applies, so the
    division is fused into the Adam kernel instead of costing another pass over the buffer."""
    world = dist.get_world_size(group)
    if world > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world
