"""Bridge between torch.autograd and the engine's explicit backward tape, so the reference's loop
(`loss.backward(); optimizer.step()`, lib/core/function.py:145-147) works unchanged on the drop-in module."""
import torch

from . import ops


class _NetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, *params):
        eng = net.engine()
        ectx = eng.forward(x, net.training, record_tape=True)
        ctx.eng, ctx.ectx, ctx.params = eng, ectx, params
        outs = tuple(ops.nhwc_to_nchw(v.data) for v in ectx.outs)
        ctx.mark_non_differentiable()
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        g_nhwc = [None if g is None else ops.nchw_to_nhwc(g.contiguous()) for g in gouts]
        pg = ctx.eng.backward(ctx.ectx, g_nhwc)
        grads = []
        for p in ctx.params:
            g = pg.get(p)
            if g is None:
                g = torch.zeros_like(p)
            grads.append(g.reshape(p.shape))
        ctx.ectx = None
        return (None, None, *grads)


def run(net, x):
    params = [p for p in net.parameters()]
    needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    if needs_grad:
        outs = _NetFunction.apply(net, x, *params)
        return list(outs)
    ectx = net.engine().forward(x, net.training, record_tape=False)
    return [ops.nhwc_to_nchw(v.data) for v in ectx.outs]
