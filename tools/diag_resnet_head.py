"""Diagnostic: where does the pose_resnet backward first leave the oracle? Compares dL/d(activation) at the deconv-head
boundaries (engine Vars vs torch autograd through oracle/resnet_oracle.py) for a golden case. GPU only."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpd_b200  # noqa: E402,F401
from fpd_b200 import ops  # noqa: E402
import test_resnet_gpu as T  # noqa: E402
from oracle import hourglass_oracle as O  # noqa: E402
from oracle import resnet_oracle as RO  # noqa: E402


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def main(tag):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # unit: CUDA-core dgrad, 3x3, 17 output channels
    torch.manual_seed(0)
    w = torch.randn(17, 32, 3, 3, device="cuda")
    dy = torch.randn(2, 16, 16, 17, device="cuda")
    dx = ops.conv2d_simt_dgrad(dy, w, (16, 16), stride=1, pad=1)
    ref = F.conv_transpose2d(dy.permute(0, 3, 1, 2).double(), w.double(), padding=1).permute(0, 2, 3, 1)
    print("simt dgrad 3x3 Cout=17:", rel(dx, ref))

    g = T._gold(tag)
    net = T._net(T._cfg(*T.CASES[tag]))
    net.train()
    x = torch.from_numpy(g["x"]).cuda()
    target = torch.from_numpy(g["target"]).cuda()
    tw = torch.from_numpy(g["target_weight"]).cuda()
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    eng = net.engine()
    rec = []
    orig_deconv, orig_block = eng.deconv, eng.block_seq

    def deconv(ctx, xv, name, *a, **k):
        out = orig_deconv(ctx, xv, name, *a, **k)
        rec.append((name, xv, out))
        return out
    eng.deconv = deconv
    bnrec = {}
    orig_bnb = eng._bn_backward

    def bnb(c_, xv, bn_name, relu, da, aff):
        if bn_name.startswith("deconv_layers"):
            bnrec[bn_name] = (da.clone(), xv, aff, relu)
        return orig_bnb(c_, xv, bn_name, relu, da, aff)
    eng._bn_backward = bnb
    ctx = eng.forward(x, True, record_tape=True)
    _, grads = ops.fpd_loss([v.data for v in ctx.outs], target, None, tw, 0.0)
    eng.backward(ctx, grads)

    # oracle with retained intermediates (fp64)
    sd = {k: (v.double() if v.is_floating_point() else v) for k, v in sd0.items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    xx = x.double()
    h = F.relu(RO._bn(sd, "bn1", F.conv2d(xx, sd["conv1.weight"], stride=2, padding=3), True))
    h = F.max_pool2d(h, 3, 2, 1)
    for l in (1, 2, 3, 4):
        i = 0
        while RO._has(sd, "layer%d.%d.conv1" % (l, i)):
            h = RO._block(sd, "layer%d.%d" % (l, i), h, 2 if (i == 0 and l > 1) else 1, True)
            i += 1
    keep = []
    i = 0
    while RO._has(sd, "deconv_layers.%d" % (3 * i)):
        wd = sd["deconv_layers.%d.weight" % (3 * i)]
        pad, outpad = RO.DECONV_GEOMETRY[wd.shape[-1]]
        h.retain_grad()
        d = F.conv_transpose2d(h, wd, sd.get("deconv_layers.%d.bias" % (3 * i)), stride=2, padding=pad, output_padding=outpad)
        d.retain_grad()
        keep.append((h, d))
        h = F.relu(RO._bn(sd, "deconv_layers.%d" % (3 * i + 1), d, True))
        h.retain_grad()
        keep[-1] = keep[-1] + (h,)
        i += 1
    wf = sd["final_layer.weight"]
    out = F.conv2d(h, wf, sd["final_layer.bias"], padding=1 if wf.shape[-1] == 3 else 0)
    loss, _, _ = O.fpd_loss([out], target.double(), tw.double())
    loss.backward()
    print("forward out:", rel(ctx.outs[0].data.permute(0, 3, 1, 2), out.detach()))
    pg = ctx.pgrads
    named = dict(net.named_parameters())
    for j, (h_, d_, a_) in enumerate(keep):
        bn = "deconv_layers.%d" % (3 * j + 1)
        da, xv, aff, relu = bnrec[bn]
        ref_da = a_.grad.permute(0, 2, 3, 1)
        print(bn, "relu", relu, "da vs oracle:", rel(da, ref_da), " mean", rel(aff[2], d_.detach().mean(dim=(0, 2, 3))),
              " invstd", rel(aff[3], 1.0 / (d_.detach().var(dim=(0, 2, 3), unbiased=False) + 1e-5).sqrt()))
        mask = (a_.detach() > 0).permute(0, 2, 3, 1)
        mine = (da.double() * mask).sum(dim=(0, 1, 2))
        print("    dbeta: engine vs oracle", rel(pg[named[bn + ".bias"]], sd[bn + ".bias"].grad), " recomputed from engine da + oracle mask",
              rel(mine, sd[bn + ".bias"].grad), " mask mismatches",
              int(((xv.data.double() - aff[2].double()) * aff[0].double() + aff[1].double() > 0).ne(mask).sum()))
    for (name, xin, xout), (h_, d_, a_) in zip(rec, keep):
        print(name, "act in ", rel(xin.data.permute(0, 3, 1, 2), h_.detach()), " act out", rel(xout.data.permute(0, 3, 1, 2), d_.detach()))
        print(name, "grad out", rel(xout.grad.permute(0, 3, 1, 2), d_.grad), " grad in", rel(xin.grad.permute(0, 3, 1, 2), h_.grad))
        e = (xout.grad.permute(0, 3, 1, 2).double() - d_.grad).abs()
        print("   grad-out error by row   :", ["%.1e" % v for v in e.amax(dim=(0, 1, 3)).tolist()[:6]], "...",
              ["%.1e" % v for v in e.amax(dim=(0, 1, 3)).tolist()[-3:]])
        print("   grad-out error by column:", ["%.1e" % v for v in e.amax(dim=(0, 1, 2)).tolist()[:6]], "...",
              ["%.1e" % v for v in e.amax(dim=(0, 1, 2)).tolist()[-3:]])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r18")
