"""GPU parity of the drop-in HRNet (engine_hrnet + libfpd_b200) against the reference golden vectors (small config
covering every structural case) and the oracle at the real w32 / w48 widths (BASELINE configs[3] shapes)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-3


def _load(name):
    z = np.load(os.path.join(GOLD, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _sd(gold, prefix="sd/"):
    return {k[len(prefix):]: torch.from_numpy(v.copy()) for k, v in gold.items() if k.startswith(prefix)}


def _rel(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


class Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def wrap(d):
    return Cfg({k: wrap(v) for k, v in d.items()}) if isinstance(d, dict) else d


def stage(nmod, chans, blocks):
    return dict(NUM_MODULES=nmod, NUM_BRANCHES=len(chans), BLOCK='BASIC', NUM_BLOCKS=[blocks] * len(chans),
                NUM_CHANNELS=chans, FUSE_METHOD='SUM')


def hr_cfg(kind):
    if kind == "small":
        st = (stage(1, [8, 16], 1), stage(2, [8, 16, 32], 1), stage(1, [8, 16, 32, 64], 1))
    else:  # experiments/fpd_coco/hrnet/w32|w48_256x192_adam_lr1e-3.yaml
        w = 32 if kind == "w32" else 48
        st = (stage(1, [w, 2 * w], 4), stage(4, [w, 2 * w, 4 * w], 4), stage(3, [w, 2 * w, 4 * w, 8 * w], 4))
    return wrap(dict(MODEL=dict(NUM_JOINTS=17, INIT_WEIGHTS=False, PRETRAINED='', EXTRA=dict(
        PRETRAINED_LAYERS=['*'], FINAL_CONV_KERNEL=1, STAGE2=st[0], STAGE3=st[1], STAGE4=st[2]))))


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference_mode():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _net(kind, sd=None):
    import fpd_b200  # noqa: F401
    from fpd_b200.lib.models import pose_hrnet as H
    net = H.get_pose_net(hr_cfg(kind), is_train=False)
    if sd is not None:
        net.load_state_dict(sd, strict=True)
    return net.cuda()


def test_hrnet_small_matches_reference_golden():
    from fpd_b200.lib.core.loss import JointsMSELoss
    g = _load("hrnet_small.npz")
    net = _net("small", _sd(g))
    x = torch.from_numpy(g["x"]).cuda()
    net.eval()
    with torch.no_grad():
        out_eval = net(x)
    assert torch.is_tensor(out_eval) and tuple(out_eval.shape) == g["out_eval"].shape
    assert _rel(out_eval, g["out_eval"]) < TOL
    net.train()
    out = net(x)
    assert _rel(out.detach(), g["out_train"]) < TOL
    loss = JointsMSELoss(True)(out, torch.from_numpy(g["target"]).cuda(), torch.from_numpy(g["target_weight"]).cuda())
    assert abs(loss.item() - float(g["loss"])) < TOL * abs(float(g["loss"]))
    loss.backward()
    named = dict(net.named_parameters())
    # whole-gradient checks (train-mode BN gradients are ill-conditioned per tensor; see test_hourglass_gpu.py)
    norms = dict(zip(g["grad_names"].tolist(), g["grad_norms"].tolist()))
    tot_ref = sum(v * v for v in norms.values()) ** 0.5
    tot = sum(p.grad.double().norm().item() ** 2 for p in named.values()) ** 0.5
    assert abs(tot - tot_ref) < 5e-2 * tot_ref, (tot, tot_ref)
    num = den = 0.0
    for k in [k[5:] for k in g if k.startswith("grad/")]:
        ref = torch.from_numpy(g["grad/" + k]).double()
        num += float((named[k].grad.double().cpu() - ref).pow(2).sum())
        den += float(ref.pow(2).sum())
    assert (num / den) ** 0.5 < 5e-2, (num / den) ** 0.5


@pytest.mark.parametrize("kind,B,training", [("w32", 2, True), ("w32", 2, False), ("w48", 2, False)])
def test_hrnet_forward_matches_oracle_at_baseline_widths(kind, B, training):
    from oracle import hrnet_oracle as HO
    torch.manual_seed(0)
    net = _net(kind)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    net.train(training)
    x = torch.randn(B, 3, 256, 192, device="cuda")
    sd = HO.annotate_strides({k: v.clone() for k, v in net.state_dict().items()})
    with torch.no_grad():
        ref = HO.hrnet(sd, x, training=training)
        got = net(x)
    assert tuple(got.shape) == (B, 17, 64, 48)
    assert _rel(got, ref) < TOL, "rel %.3e" % _rel(got, ref)


def test_hrnet_fpd_train_step_matches_oracle_losses():
    """BASELINE configs[3] path at test size: HRNet student (train) + frozen HRNet teacher (eval), FPD loss, backward,
    flat Adam through FPDTrainStep; loss terms against the oracle (hrnet_oracle + fpd_loss)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from fpd_b200.train_step import FPDTrainStep
    from oracle import hourglass_oracle as O
    from oracle import hrnet_oracle as HO
    torch.manual_seed(21)
    student = _net("small")
    teacher = _net("small")
    for m in teacher.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    B = 2
    x = torch.randn(B, 3, 128, 96, device="cuda")
    target = torch.rand(B, 17, 32, 24, device="cuda")
    tw = (torch.rand(B, 17, 1, device="cuda") > 0.2).float()
    s_sd = HO.annotate_strides({k: v.clone() for k, v in student.state_dict().items()})
    t_sd = HO.annotate_strides({k: v.clone() for k, v in teacher.state_dict().items()})
    with torch.no_grad():
        out = HO.hrnet(s_sd, x, training=True)
        tout = HO.hrnet(t_sd, x, training=False)
        ref_total, ref_pose, ref_kd = O.fpd_loss([out], target, tw, tout, 0.5)
    step = FPDTrainStep(student, teacher, alpha=0.5, lr=1e-3, use_graph=False)
    w0 = step.flat.flat.clone()
    losses = step.step(x, target, tw)
    torch.cuda.synchronize()
    pose, kd, total = [float(v) for v in losses.cpu()]
    assert abs(pose - ref_pose.item()) < TOL * abs(ref_pose.item())
    assert abs(kd - ref_kd.item()) < TOL * abs(ref_kd.item())
    assert abs(total - ref_total.item()) < TOL * abs(ref_total.item())
    assert not torch.equal(w0, step.flat.flat)
    assert torch.isfinite(step.flat.flat).all()


def test_hrnet_w32_w48_fpd_train_step_at_baseline_widths():
    """BASELINE configs[3] at the real widths (w32 student, frozen w48 teacher, 256x192) and a small batch: the whole
    train step runs (incl. the wide stride-2 convolutions' weight gradients, Cout 256 / 384) and its loss terms match
    the oracle."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from fpd_b200.train_step import FPDTrainStep
    from oracle import hourglass_oracle as O
    from oracle import hrnet_oracle as HO
    torch.manual_seed(23)
    student, teacher = _net("w32"), _net("w48")
    for m in teacher.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    B = 2
    x = torch.randn(B, 3, 256, 192, device="cuda")
    target = torch.rand(B, 17, 64, 48, device="cuda")
    tw = (torch.rand(B, 17, 1, device="cuda") > 0.2).float()
    s_sd = HO.annotate_strides({k: v.clone() for k, v in student.state_dict().items()})
    t_sd = HO.annotate_strides({k: v.clone() for k, v in teacher.state_dict().items()})
    with torch.no_grad():
        out = HO.hrnet(s_sd, x, training=True)
        tout = HO.hrnet(t_sd, x, training=False)
        ref_total, ref_pose, ref_kd = O.fpd_loss([out], target, tw, tout, 0.5)
    step = FPDTrainStep(student, teacher, alpha=0.5, lr=1e-3, use_graph=True)
    w0 = step.flat.flat.clone()
    losses = step.step(x, target, tw)
    torch.cuda.synchronize()
    pose, kd, total = [float(v) for v in losses.cpu()]
    assert abs(pose - ref_pose.item()) < TOL * abs(ref_pose.item())
    assert abs(kd - ref_kd.item()) < TOL * abs(ref_kd.item())
    assert abs(total - ref_total.item()) < TOL * abs(ref_total.item())
    assert torch.isfinite(step.flat.flat).all() and not torch.equal(w0, step.flat.flat)
    assert torch.isfinite(step.flat.grad).all() and float(step.flat.grad.abs().max()) > 0
