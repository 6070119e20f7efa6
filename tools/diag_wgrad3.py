"""Correctness (against torch) and timing of the halo-tile weight-gradient kernel (csrc/wgrad_tc3.cu). Probes the
TMEM row mapping of M = 64 accumulators (FPD_WGRAD3_LANEMAP 0/1/2) and reports each shape; never stops at a failure."""
import os
import sys
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fpd_b200  # noqa: F401
    from fpd_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    lib = ops.N.lib()
    shapes = [(2, 32, 32, 64, 64), (2, 64, 64, 64, 64), (3, 16, 16, 64, 64), (4, 8, 8, 64, 64), (2, 64, 48, 64, 32),
              (2, 32, 32, 128, 128), (2, 16, 16, 128, 64), (2, 32, 24, 64, 96)]
    for lm in (1, 0, 2):
        os.environ["FPD_WGRAD3_LANEMAP"] = str(lm)
        for (B, H, W, Cin, Cout) in shapes:
            try:
                g = torch.Generator(device="cuda").manual_seed(3)
                x = torch.randn(B, Cin, H, W, device="cuda", generator=g) * 2 + 0.3
                dy = torch.randn(B, Cout, H, W, device="cuda", generator=g)
                mean = torch.randn(Cin, device="cuda", generator=g)
                scale = torch.rand(Cin, device="cuda", generator=g) + 0.5
                shift = torch.randn(Cin, device="cuda", generator=g) * 0.5
                a = F.relu((x - mean.view(1, -1, 1, 1)) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
                ref = torch.nn.grad.conv2d_weight(a, (Cout, Cin, 3, 3), dy, padding=1)
                dw = ops.conv2d_wgrad_tc_fused(x.permute(0, 2, 3, 1).contiguous(), dy.permute(0, 2, 3, 1).contiguous(), 3,
                                               mean=mean, scale=scale, shift=shift, relu=True)
                torch.cuda.synchronize()
                err = ((dw - ref).abs().max() / ref.abs().max()).item()
                # per-tap error to localise layout problems
                pt = [((dw[:, :, t // 3, t % 3] - ref[:, :, t // 3, t % 3]).abs().max() / ref.abs().max()).item() for t in range(9)]
                print("lanemap=%d %s rel err %.3e %s taps %s" % (lm, (B, H, W, Cin, Cout), err, "OK" if err < 5e-5 else "MISMATCH",
                                                               " ".join("%.0e" % e for e in pt)), flush=True)
            except Exception:
                print("lanemap=%d %s EXC\n%s" % (lm, (B, H, W, Cin, Cout), traceback.format_exc()), flush=True)
                try:
                    torch.cuda.synchronize()
                except Exception:
                    print("device unusable; stopping")
                    return 1
    os.environ["FPD_WGRAD3_LANEMAP"] = "1"
    print("timing (us per launch incl. reduce, 5 launches per region): new vs FPD_WGRAD3=0 needs a separate process")
    for (B, H, W, Cin, Cout) in [(32, 64, 64, 64, 64), (32, 32, 32, 64, 64), (32, 16, 16, 64, 64), (32, 8, 8, 64, 64)]:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
        dy = torch.randn(B, H, W, Cout, device="cuda", generator=g)
        mean = torch.zeros(Cin, device="cuda"); scale = torch.ones(Cin, device="cuda"); shift = torch.zeros(Cin, device="cuda")
        for _ in range(2):
            ops.conv2d_wgrad_tc_fused(x, dy, 3, mean=mean, scale=scale, shift=shift, relu=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.conv2d_wgrad_tc_fused(x, dy, 3, mean=mean, scale=scale, shift=shift, relu=True)
        e1.record()
        torch.cuda.synchronize()
        print("%s %.1f us" % ((B, H, W, Cin, Cout), e0.elapsed_time(e1) / 5 * 1000.0), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
