"""Functional fp32 restatement of the reference's HRNet pose network (test oracle, CPU/any device).

Evaluated from a reference-keyed `state_dict` with torch.nn.functional primitives; the module structure (branches,
blocks, fuse rows) is recovered from the key names, so one function serves w32 / w48 / the small test config.
Reference: lib/models/pose_hrnet.py -- BasicBlock.forward :41-57, Bottleneck.forward :78-98,
HighResolutionModule.forward :247-265, PoseHighResolutionNet.forward :425-460.
"""
import torch.nn.functional as F

BN_MOMENTUM = 0.1


def _bn(sd, name, x, training, momentum=BN_MOMENTUM):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training, momentum, 1e-5)


def _conv(sd, name, x):
    w = sd[name + ".weight"]
    k = w.shape[-1]
    stride = sd.get("__stride__." + name, 1)
    return F.conv2d(x, w, sd.get(name + ".bias"), stride=stride, padding=k // 2)


def _has(sd, name):
    return (name + ".weight") in sd


def _block(sd, p, x, training):
    """BasicBlock (:41-57) or Bottleneck (:78-98), decided by the presence of conv3."""
    out = F.relu(_bn(sd, p + ".bn1", _conv(sd, p + ".conv1", x), training))
    if _has(sd, p + ".conv3"):
        out = F.relu(_bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out), training))
        out = _bn(sd, p + ".bn3", _conv(sd, p + ".conv3", out), training)
    else:
        out = _bn(sd, p + ".bn2", _conv(sd, p + ".conv2", out), training)
    skip = x
    if _has(sd, p + ".downsample.0"):
        skip = _bn(sd, p + ".downsample.1", _conv(sd, p + ".downsample.0", x), training)
    return F.relu(out + skip)


def _blocks(sd, p, x, training):
    i = 0
    while _has(sd, "%s.%d.conv1" % (p, i)):
        x = _block(sd, "%s.%d" % (p, i), x, training)
        i += 1
    return x


def _chain(sd, p, x, training, relu_last):
    """Sequential of Sequential(conv3x3 s2, BN[, ReLU]) (transition new branches :355-370, fuse down paths :213-239).
    NB: fuse-layer BatchNorms use the nn.BatchNorm2d default momentum (0.1) as well."""
    n = 0
    while _has(sd, "%s.%d.0" % (p, n)):
        n += 1
    for k in range(n):
        x = _bn(sd, "%s.%d.1" % (p, k), _conv(sd, "%s.%d.0" % (p, k), x), training)
        if relu_last or k < n - 1:
            x = F.relu(x)
    return x


def _module(sd, p, xs, training):
    """HighResolutionModule.forward (:247-265)."""
    nb = len(xs)
    xs = [_blocks(sd, "%s.branches.%d" % (p, b), xs[b], training) for b in range(nb)]
    if nb == 1:
        return xs
    rows = 0
    while any(k.startswith("%s.fuse_layers.%d." % (p, rows)) for k in sd):
        rows += 1
    outs = []
    for i in range(rows):
        y = None
        for j in range(nb):
            fp = "%s.fuse_layers.%d.%d" % (p, i, j)
            if j == i:
                t = xs[j]
            elif j > i:
                t = _bn(sd, fp + ".1", _conv(sd, fp + ".0", xs[j]), training)
                t = F.interpolate(t, scale_factor=2 ** (j - i), mode="nearest")
            else:
                t = _chain(sd, fp, xs[j], training, relu_last=False)
            y = t if y is None else y + t
        outs.append(F.relu(y))
    return outs


def annotate_strides(sd):
    """Strides are not part of a state_dict; they follow from the architecture (stem convs, transition new-branch
    chains and fuse down-chains are stride 2, everything else stride 1). Returns a copy of `sd` carrying them."""
    out = dict(sd)
    for k in sd:
        if not k.endswith(".weight") or sd[k].dim() != 4:
            continue
        name = k[:-7]
        parts = name.split(".")
        s2 = name in ("conv1", "conv2")
        if parts[0].startswith("transition") and len(parts) == 4:         # transitionN.i.j.0
            s2 = True
        if "fuse_layers" in parts and len(parts) - parts.index("fuse_layers") == 5:   # fuse_layers.i.j.k.0
            s2 = True
        out["__stride__." + name] = 2 if s2 else 1
    return out


def hrnet(sd, x, training=True):
    """PoseHighResolutionNet.forward (:425-460). `sd` must come from annotate_strides()."""
    x = F.relu(_bn(sd, "bn1", _conv(sd, "conv1", x), training))
    x = F.relu(_bn(sd, "bn2", _conv(sd, "conv2", x), training))
    x = _blocks(sd, "layer1", x, training)
    ys = [x]
    for s in (2, 3, 4):
        tname = "transition%d" % (s - 1)
        nb = 0
        while any(k.startswith("stage%d.0.branches.%d." % (s, nb)) for k in sd):
            nb += 1
        xs = []
        for i in range(nb):
            tp = "%s.%d" % (tname, i)
            if _has(sd, tp + ".0"):
                xs.append(F.relu(_bn(sd, tp + ".1", _conv(sd, tp + ".0", ys[-1]), training)))
            elif _has(sd, tp + ".0.0"):
                xs.append(_chain(sd, tp, ys[-1], training, relu_last=True))
            else:
                xs.append(ys[i])
        m = 0
        while any(k.startswith("stage%d.%d." % (s, m)) for k in sd):
            xs = _module(sd, "stage%d.%d" % (s, m), xs, training)
            m += 1
        ys = xs
    return _conv(sd, "final_layer", ys[0])
