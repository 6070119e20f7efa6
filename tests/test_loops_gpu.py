"""The drop-in loops (`lib/core/function.py`: train / fpd_train / validate keep the reference signatures): the fused fast
path must give the same updates / predictions as the generic autograd route that mirrors the reference statement by
statement, on synthetic loaders."""
import logging
import os
import sys
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
NS = types.SimpleNamespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _cfg(f, s):
    return NS(MODEL=NS(EXTRA=NS(NUM_FEATURES=f, NUM_STACKS=s, NUM_BLOCKS=1), NUM_JOINTS=16, NAME="hourglass"),
              KD=NS(ALPHA=0.5), PRINT_FREQ=1,
              TEST=NS(FLIP_TEST=True, SHIFT_HEATMAP=True, POST_PROCESS=True), DEBUG=NS(DEBUG=False))


def _net(f, s, sd=None):
    import fpd_b200  # noqa: F401
    from fpd_b200.lib.models import hourglass as H
    net = H.get_pose_net(_cfg(f, s), True)
    if sd is not None:
        net.load_state_dict(sd)
    return net.cuda()


class _Hide(torch.nn.Module):
    """Wrapper that hides the engine attributes -> function.py takes the generic (reference-like) route."""

    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x):
        return self.net(x.cuda())


def _loader(n, B=2, hw=128):
    from bench import synthetic_batch
    out = []
    for i in range(n):
        x, t, w = synthetic_batch(B, 70 + i, hw, hw)
        meta = {"center": torch.tensor([[64.0, 64.0]] * B), "scale": torch.tensor([[0.64, 0.64]] * B),
                "score": torch.ones(B), "image": ["img%d_%d" % (i, b) for b in range(B)]}
        out.append((x, t, w, meta))
    return out


class _Writer:
    def __init__(self):
        self.scalars = []

    def add_scalar(self, tag, v, step):
        self.scalars.append((tag, float(v), step))

    def add_scalars(self, tag, d, step):
        pass


def test_fpd_train_fast_path_equals_generic_route():
    from fpd_b200.lib.core.function import fpd_train, train
    from fpd_b200.lib.core.loss import JointsMSELoss
    logging.basicConfig(level=logging.INFO)
    torch.manual_seed(5)
    init_s = {k: v.clone() for k, v in _net(64, 2).state_dict().items()}
    init_t = {k: v.clone() for k, v in _net(64, 1).state_dict().items()}
    cfg = _cfg(64, 2)
    loader = _loader(2)
    results = []
    for generic in (False, True):
        s, t = _net(64, 2, init_s), _net(64, 1, init_t)
        model = _Hide(s) if generic else s
        opt = torch.optim.Adam(s.parameters(), lr=1e-3)
        wd = {"writer": _Writer(), "train_global_steps": 0}
        fpd_train(cfg, loader, model, t, JointsMSELoss(True), JointsMSELoss(True), opt, 0, "/tmp", "/tmp", wd)
        torch.cuda.synchronize()
        assert wd["train_global_steps"] == 2
        tags = {tag for tag, _, _ in wd["writer"].scalars}
        assert {"train_loss", "train_pose_loss", "train_kd_pose_loss", "train_acc"} <= tags
        results.append(({k: v.clone() for k, v in s.state_dict().items()}, wd["writer"].scalars))
    (sd_f, sc_f), (sd_g, sc_g) = results
    for (tf, vf, _), (tg, vg, _) in zip(sc_f, sc_g):
        assert tf == tg and abs(vf - vg) <= 1e-4 * max(1.0, abs(vg)), (tf, vf, vg)
    # the two routes run the same kernels in the same order -> weights agree to round-off
    worst = max(((sd_f[k].double() - sd_g[k].double()).abs().max() / sd_g[k].double().abs().max().clamp_min(1e-12)).item()
                for k in sd_f if sd_f[k].is_floating_point())
    assert worst < 1e-4, worst
    # plain train() (no teacher) runs through the same machinery
    s = _net(64, 2, init_s)
    opt = torch.optim.Adam(s.parameters(), lr=1e-3)
    wd = {"writer": _Writer(), "train_global_steps": 0}
    train(cfg, loader[:1], s, JointsMSELoss(True), opt, 0, "/tmp", "/tmp", wd)
    assert wd["train_global_steps"] == 1 and not torch.equal(s.state_dict()["conv1.weight"], init_s["conv1.weight"].cuda())


class _ValSet:
    flip_pairs = [[0, 5], [1, 4], [2, 3], [10, 15], [11, 14], [12, 13]]

    def __init__(self, n):
        self.n = n
        self.seen = None

    def __len__(self):
        return self.n

    def evaluate(self, cfg, preds, output_dir, all_boxes, image_path, filenames, imgnums):
        self.seen = (preds.copy(), all_boxes.copy(), list(image_path))
        return {"Mean": 12.5}, 12.5


def test_validate_fast_path_equals_generic_route():
    from fpd_b200.lib.core.function import validate
    from fpd_b200.lib.core.loss import JointsMSELoss
    torch.manual_seed(6)
    net = _net(64, 2)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    cfg = _cfg(64, 2)
    loader = _loader(2)
    seen = []
    sys.path.insert(0, os.path.join(ROOT, "fast-human-pose-estimation.pytorch_b200", "lib"))  # `utils.transforms` (generic route)
    for generic in (False, True):
        ds = _ValSet(4)
        perf = validate(cfg, loader, ds, _Hide(net) if generic else net, JointsMSELoss(True), "/tmp", "/tmp",
                        {"writer": _Writer(), "valid_global_steps": 0})
        assert perf == 12.5
        seen.append(ds.seen)
    (pf, bf, nf), (pg, bg, ng) = seen
    assert nf == ng and np.allclose(bf, bg)
    assert pf.shape == (4, 16, 3)
    # key points: identical arg-max cells; max values to round-off
    assert np.abs(pf[:, :, :2] - pg[:, :, :2]).max() < 1e-3
    assert np.abs(pf[:, :, 2] - pg[:, :, 2]).max() < 1e-5 * np.abs(pg[:, :, 2]).max() + 1e-7
