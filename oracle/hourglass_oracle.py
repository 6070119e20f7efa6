"""Functional fp32 restatement of the reference's stacked hourglass + FPD loss (test oracle, CPU).

Every function cites the reference lines it restates. The network is evaluated straight from a
`state_dict` carrying the reference's key names, with torch.nn.functional primitives (conv2d, batch_norm,
relu, max_pool2d, interpolate) -- the same ATen ops the reference's nn.Modules dispatch to -- so autograd
through these functions yields the reference's gradients as well.
"""
import torch
import torch.nn.functional as F

BN_MOMENTUM = 0.1  # reference lib/models/hourglass.py:10
BN_EPS = 1e-5      # nn.BatchNorm2d default


def _bn(sd, name, x, training):
    # nn.BatchNorm2d(ch, momentum=0.1): hourglass.py:18,21,25,117,162
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training, BN_MOMENTUM, BN_EPS)


def _conv(sd, name, x, stride=1, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def bottleneck(sd, p, x, training):
    """Bottleneck.forward, hourglass.py:32-52 (pre-activation; optional 1x1 `downsample` on the skip)."""
    out = _conv(sd, p + ".conv1", F.relu(_bn(sd, p + ".bn1", x, training)))
    out = _conv(sd, p + ".conv2", F.relu(_bn(sd, p + ".bn2", out, training)), padding=1)
    out = _conv(sd, p + ".conv3", F.relu(_bn(sd, p + ".bn3", out, training)))
    skip = _conv(sd, p + ".downsample.0", x) if (p + ".downsample.0.weight") in sd else x
    return out + skip


def _seq(sd, p, x, nblocks, training):
    for i in range(nblocks):
        x = bottleneck(sd, "%s.%d" % (p, i), x, training)
    return x


def hour_glass(sd, p, n, x, nblocks, training):
    """Hourglass._hour_glass_forward, hourglass.py:80-92."""
    up1 = _seq(sd, "%s.%d.0" % (p, n - 1), x, nblocks, training)
    low1 = F.max_pool2d(x, 2, stride=2)
    low1 = _seq(sd, "%s.%d.1" % (p, n - 1), low1, nblocks, training)
    if n > 1:
        low2 = hour_glass(sd, p, n - 1, low1, nblocks, training)
    else:
        low2 = _seq(sd, "%s.%d.3" % (p, n - 1), low1, nblocks, training)
    low3 = _seq(sd, "%s.%d.2" % (p, n - 1), low2, nblocks, training)
    return up1 + F.interpolate(low3, scale_factor=2, mode="nearest")


def hourglass_net(sd, x, num_stacks, num_blocks=1, training=True):
    """HourglassNet.forward, hourglass.py:170-192. `sd`: reference-keyed state_dict (tensors may require
    grad). In training mode the running statistics inside `sd` are updated in place like the reference."""
    x = F.relu(_bn(sd, "bn1", _conv(sd, "conv1", x, stride=2, padding=3), training))
    x = _seq(sd, "layer1", x, 1, training)
    x = F.max_pool2d(x, 2, stride=2)
    x = _seq(sd, "layer2", x, 1, training)
    x = _seq(sd, "layer3", x, 1, training)
    out = []
    for i in range(num_stacks):
        y = hour_glass(sd, "hg.%d.hg" % i, 4, x, num_blocks, training)
        y = _seq(sd, "res.%d" % i, y, num_blocks, training)
        y = F.relu(_bn(sd, "fc.%d.1" % i, _conv(sd, "fc.%d.0" % i, y), training))
        score = _conv(sd, "score.%d" % i, y)
        out.append(score)
        if i < num_stacks - 1:
            x = x + _conv(sd, "fc_.%d" % i, y) + _conv(sd, "score_.%d" % i, score)
    return out


def joints_mse(output, target, target_weight, use_target_weight=True):
    """JointsMSELoss.forward, lib/core/loss.py:21-39, in closed form:
    1/J * sum_j 0.5 * mean_{b,hw}((w_bj*pred - w_bj*gt)^2)."""
    B, J = output.shape[:2]
    d = (output - target).reshape(B, J, -1)
    if use_target_weight:
        d = d * target_weight.reshape(B, J, 1)
    return (0.5 * (d * d).mean(dim=(0, 2))).sum() / J


def fpd_loss(outputs, target, target_weight, teacher_out=None, alpha=0.5):
    """Loss combination of lib/core/function.py:127-134 (fpd_train) / :49-55 (train): the per-stack losses are
    SUMMED; the KD term compares every student stack with the teacher's LAST stack. Returns (loss, pose, kd)."""
    pose = sum(joints_mse(o, target, target_weight) for o in outputs)
    if teacher_out is None:
        return pose, pose, torch.zeros_like(pose)
    kd = sum(joints_mse(o, teacher_out, target_weight) for o in outputs)
    return (1 - alpha) * pose + alpha * kd, pose, kd
