"""Shared helpers of the GPU parity tests: the oracle run in fp32 and fp64 on the same device, and the fp64-referenced
gradient criterion.

Why fp64-referenced: the train-mode hourglass / HRNet at default init amplify round-off ~1000x, so
the reference's OWN fp32 gradients sit 1e-3 .. 5e-2 away from the exact ones, tensor by tensor. A fixed tolerance is
therefore either meaningless (5e-2) or unattainable (1e-3). Instead every parameter gradient is compared with an fp64
evaluation of the reference algorithm and must be (almost) as close to it as the reference's fp32 arithmetic gets:

    err(ours, fp64)  <=  K * err(fp32 oracle, fp64) + floor          per tensor, max-abs error / max|g64|
    and the same for the whole-gradient relative L2 error.
"""
import json
import os

import torch


def cast_sd(sd, dtype, device=None):
    out = {}
    for k, v in sd.items():
        if torch.is_tensor(v):
            v = v.detach().clone()
            if v.is_floating_point():
                v = v.to(dtype)
            if device is not None:
                v = v.to(device)
        out[k] = v
    return out


def with_grad(sd):
    params = {k: v.requires_grad_(True) for k, v in sd.items()
              if torch.is_tensor(v) and v.is_floating_point() and "running" not in k and not k.startswith("__")}
    sd.update(params)
    return params


def rel_max(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def grad_report(ours, g32, g64, tag=None):
    """ours / g32 / g64: dict name -> gradient tensor. Returns a dict with per-tensor (e_ours, e_32) and the whole-gradient
    relative L2 errors; optionally appended to gpurun_out/parity_report.jsonl for the profiles/ summary."""
    rows = []
    n_o = n_3 = den = 0.0
    for k, ref in g64.items():
        ref = ref.double().cpu()
        o = ours[k].reshape(ref.shape).double().cpu()
        f = g32[k].reshape(ref.shape).double().cpu()
        m = ref.abs().max().clamp_min(1e-300).item()
        rows.append((k, (o - ref).abs().max().item() / m, (f - ref).abs().max().item() / m))
        n_o += float((o - ref).pow(2).sum())
        n_3 += float((f - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    rep = {"tag": tag, "tensors": len(rows), "l2_ours": (n_o / den) ** 0.5, "l2_fp32": (n_3 / den) ** 0.5,
           "worst_ours": max(r[1] for r in rows), "worst_fp32": max(r[2] for r in rows),
           "worst_ratio": max(r[1] / max(r[2], 1e-12) for r in rows),
           "median_ratio": sorted(r[1] / max(r[2], 1e-12) for r in rows)[len(rows) // 2], "rows": rows}
    if tag is not None:
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        try:
            os.makedirs(out, exist_ok=True)
            with open(os.path.join(out, "parity_report.jsonl"), "a") as fh:
                fh.write(json.dumps({k: v for k, v in rep.items() if k != "rows"}) + "\n")
        except OSError:
            pass
    return rep


# Per tensor our error may exceed the fp32 reference's by this factor (+ floor): the 3xFP16 / 3xTF32 operand pairs carry
# 22 significant bits against fp32's 24, so a little more than 2x the reference's own round-off is expected.
K_TENSOR, K_L2, FLOOR = 4.0, 3.0, 2e-6


def assert_grads_as_good_as_fp32(ours, g32, g64, tag):
    rep = grad_report(ours, g32, g64, tag)
    bad = [(k, eo, e3) for k, eo, e3 in rep["rows"] if eo > K_TENSOR * e3 + FLOOR]
    assert not bad, "%s: %d tensors further from fp64 than %.0fx the fp32 reference: %s" % (tag, len(bad), K_TENSOR, bad[:5])
    assert rep["l2_ours"] <= K_L2 * rep["l2_fp32"] + FLOOR, (tag, rep["l2_ours"], rep["l2_fp32"])
    return rep
