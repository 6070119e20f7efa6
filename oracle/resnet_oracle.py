"""Functional fp32/fp64 restatement of the reference's ResNet + deconv-head pose network (test oracle, CPU/any device).

TEST INFRASTRUCTURE ONLY -- nothing under fast-human-pose-estimation.pytorch_b200/ imports this. Evaluated from a
reference-keyed `state_dict` with torch.nn.functional primitives; depth and block type are recovered from the key names
(conv3 present -> Bottleneck), the head geometry from the deconv kernel size, so one function serves ResNet-18 ... 152.
Reference: lib/models/pose_resnet.py -- BasicBlock.forward :43-59, Bottleneck.forward :80-101, _make_layer :140-157 (stride
on the FIRST block of layer2..4: conv1 of a BasicBlock, conv2 of a Bottleneck, and the 1x1 downsample), _get_deconv_cfg
:159-174, PoseResNet.forward :230-246.
Pinned against the reference module's outputs / gradients in tests/golden/resnet_*.npz (tests/test_oracle.py).
"""
import torch
import torch.nn.functional as F

BN_MOMENTUM = 0.1
DECONV_GEOMETRY = {4: (1, 0), 3: (1, 1), 2: (0, 0)}     # kernel -> (padding, output_padding)


def _bn(sd, name, x, training):
    return F.batch_norm(x, sd[name + ".running_mean"], sd[name + ".running_var"], sd[name + ".weight"],
                        sd[name + ".bias"], training, BN_MOMENTUM, 1e-5)


def _has(sd, name):
    return (name + ".weight") in sd


def _block(sd, p, x, stride, training):
    bottleneck = _has(sd, p + ".conv3")
    if bottleneck:
        out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"]), training))
        out = F.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), training))
        out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]), training)
    else:
        out = F.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"], stride=stride, padding=1), training))
        out = _bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], padding=1), training)
    skip = x
    if _has(sd, p + ".downsample.0"):
        skip = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), training)
    return F.relu(out + skip)


def resnet(sd, x, training=True):
    """PoseResNet.forward (:230-246) -> heat-maps [B, J, H/4, W/4]."""
    x = F.relu(_bn(sd, "bn1", F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), training))
    x = F.max_pool2d(x, 3, 2, 1)
    for l in (1, 2, 3, 4):
        i = 0
        while _has(sd, "layer%d.%d.conv1" % (l, i)):
            x = _block(sd, "layer%d.%d" % (l, i), x, 2 if (i == 0 and l > 1) else 1, training)
            i += 1
    i = 0
    while _has(sd, "deconv_layers.%d" % (3 * i)):
        w = sd["deconv_layers.%d.weight" % (3 * i)]
        pad, outpad = DECONV_GEOMETRY[w.shape[-1]]
        x = F.conv_transpose2d(x, w, sd.get("deconv_layers.%d.bias" % (3 * i)), stride=2, padding=pad,
                               output_padding=outpad)
        x = F.relu(_bn(sd, "deconv_layers.%d" % (3 * i + 1), x, training))
        i += 1
    w = sd["final_layer.weight"]
    return F.conv2d(x, w, sd["final_layer.bias"], padding=1 if w.shape[-1] == 3 else 0)


def synthetic_state(shapes, seed=0, dtype=torch.float32):
    """Deterministic, well-conditioned parameters for a state_dict layout {key: shape} (the ResNets are far too large to
    commit as golden weights): He-scaled convolutions so activations stay O(1) through 50+ layers, BatchNorm parameters and
    running statistics away from their trivial values. One generator per tensor, seeded by position."""
    out = {}
    for i, (k, shape) in enumerate(shapes.items()):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        shape = tuple(shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_mean"):
            out[k] = (0.1 * torch.randn(shape, generator=g)).to(dtype)
        elif k.endswith("running_var"):
            out[k] = (0.75 + 0.5 * torch.rand(shape, generator=g)).to(dtype)
        elif len(shape) == 4:
            if k.startswith("deconv_layers"):
                fan_in = shape[0] * (shape[2] * shape[3]) / 4.0     # a stride-2 transposed conv sees k*k/4 taps per output
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            out[k] = (torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5).to(dtype)
        elif k.endswith(".weight"):
            out[k] = (1.0 + 0.1 * torch.randn(shape, generator=g)).to(dtype)
        else:
            out[k] = (0.1 * torch.randn(shape, generator=g)).to(dtype)
    return out
