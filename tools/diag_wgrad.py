"""Timing ablation of the fused weight-gradient kernel (csrc/wgrad_tc2.cu): FPD_WGRAD_DBG bits 1 no MMA, 2 no TMA,
8 no transform. Several launches per timed region so the host launch cost does not dominate."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fpd_b200  # noqa: F401
    from fpd_b200 import ops
    shapes = [(32, 64, 64, 64, 64, 3), (32, 64, 64, 128, 64, 1), (32, 64, 64, 64, 128, 1), (32, 32, 32, 64, 64, 3),
              (32, 16, 16, 64, 64, 3), (32, 128, 128, 32, 32, 3)]
    masks = [0, 1, 8, 2, 9, 10, 11]
    print("%-28s " % "shape" + " ".join("%7d" % m for m in masks) + "   (us per launch, 5 launches per timed region)")
    for (B, H, W, Cin, Cout, k) in shapes:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
        dy = torch.randn(B, H, W, Cout, device="cuda", generator=g)
        mean = torch.zeros(Cin, device="cuda"); scale = torch.ones(Cin, device="cuda"); shift = torch.zeros(Cin, device="cuda")
        row = []
        for m in masks:
            os.environ["FPD_WGRAD_DBG"] = str(m)
            for _ in range(2):
                ops.conv2d_wgrad_tc_fused(x, dy, k, mean=mean, scale=scale, shift=shift, relu=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.conv2d_wgrad_tc_fused(x, dy, k, mean=mean, scale=scale, shift=shift, relu=True)
            e1.record()
            torch.cuda.synchronize()
            row.append(e0.elapsed_time(e1) / 5 * 1000.0)
        os.environ["FPD_WGRAD_DBG"] = "0"
        print("%-28s " % str((B, H, W, Cin, Cout, k)) + " ".join("%7.1f" % t for t in row), flush=True)


if __name__ == "__main__":
    main()
