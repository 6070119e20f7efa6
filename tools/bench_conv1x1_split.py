"""Which warp split suits which 1x1 convolution? Times conv2d_tc_h (3xFP16) on the 1x1 shapes of the hourglass student /
teacher, with and without a residual, under the split chosen by FPD_CONV_EPI8_MODE (read once per process: run this script
once per mode). Rotating operand sets larger than L2, CUDA events, warm."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fpd_b200  # noqa: E402,F401
from fpd_b200 import ops  # noqa: E402

SHAPES = [  # (H, Cin, Cout): teacher conv1 / conv3, student conv1 / conv3 and their data gradients, fc / score convs
    (64, 256, 128), (64, 128, 256), (32, 256, 128), (32, 128, 256), (64, 128, 64), (64, 64, 128), (32, 128, 64),
    (32, 64, 128), (64, 256, 256), (64, 128, 128), (64, 256, 16), (64, 16, 256)]


def main():
    B = 32
    mode = os.environ.get("FPD_CONV_EPI8_MODE", os.environ.get("FPD_CONV_EPI8", "default"))
    for (H, Cin, Cout) in SHAPES:
        if not ops.conv2d_tc_h_supported(Cin, Cout, 1, H, H, True):
            continue
        sets = []
        nset = max(2, int(300e6 // (B * H * H * (Cin + 2 * Cout) * 4)) + 1)
        for i in range(nset):
            x = torch.randn(B, H, H, Cin, device="cuda")
            res = torch.randn(B, H, H, Cout, device="cuda")
            y = torch.empty(B, H, H, Cout, device="cuda")
            sets.append((x, res, y))
        w = torch.randn(Cout, Cin, 1, 1, device="cuda") * 0.05
        w_hi, w_lo = ops.weight_prep_f16(w)
        mean = torch.zeros(Cin, device="cuda"); scale = torch.ones(Cin, device="cuda"); shift = torch.zeros(Cin, device="cuda")
        bias = torch.randn(Cout, device="cuda")
        for with_res in (False, True):
            def run(i):
                x, res, y = sets[i % nset]
                ops.conv2d_tc_h(x, w_hi, w_lo, 1, mean=mean, scale=scale, shift=shift, relu=True, bias=bias,
                                residual=res if with_res else None, out=y)
            for i in range(nset):
                run(i)
            torch.cuda.synchronize()
            n = 4 * nset
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1000.0
            gb = B * H * H * (Cin + Cout * (2 if with_res else 1)) * 4 / 1e9
            print("mode=%s  %3dx%-3d %3d->%-3d res=%d  %7.1f us  %5.2f TB/s" % (mode, H, H, Cin, Cout, int(with_res), us,
                                                                              gb / us * 1e3 / 1e3), flush=True)
        del sets


if __name__ == "__main__":
    main()
