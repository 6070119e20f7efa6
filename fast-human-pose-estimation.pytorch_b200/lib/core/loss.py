"""Drop-in for the reference's lib/core/loss.py: `JointsMSELoss(use_target_weight)(output, target,
target_weight) -> 0-dim tensor` supporting .backward(), `+=`, .item() (as lib/core/function.py:127-152 uses
it). One fused CUDA pass (value + gradient) replaces the reference's per-joint Python loop of nn.MSELoss
calls (loss.py:24-39). JointsOHKMMSELoss is not provided: no entry point of the reference instantiates it."""
import torch
import torch.nn as nn


class _JointsMSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, target_weight):
        from fpd_b200 import ops
        out = output.contiguous().float()
        loss3, grad = ops.joints_mse(out, target.contiguous().float(),
                                     None if target_weight is None else target_weight.float(),
                                     want_grad=output.requires_grad)
        ctx.save_for_backward(grad)
        return loss3[0].clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (None if grad is None else grad * g), None, None


class JointsMSELoss(nn.Module):
    def __init__(self, use_target_weight):
        super().__init__()
        self.use_target_weight = use_target_weight

    def forward(self, output, target, target_weight):
        if not output.is_cuda:
            raise RuntimeError("fpd_b200 JointsMSELoss runs on CUDA tensors only (no CPU fallback)")
        tw = target_weight if self.use_target_weight else None
        return _JointsMSEFn.apply(output, target, tw)
