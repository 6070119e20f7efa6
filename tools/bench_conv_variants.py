"""Per-shape timing of the three forward-conv formulations on the hot shapes (B=32, 64x64 unless noted):
  presplit = affine_act_split kernel + conv_tc (SS, hi/lo operands from HBM)
  ss_fused = conv_tc2 (raw x, transform in smem, SS MMA)
  ts_fused = conv_tc3 (raw x, transform -> TMEM, TS MMA)
CUDA events, L2 flushed between launches; prints microseconds."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, flush, iters=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1000.0


def main():
    import fpd_b200  # noqa: F401
    from fpd_b200 import ops
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = [(32, 64, 64, 128, 128, 3), (32, 64, 64, 64, 64, 3), (32, 64, 64, 128, 64, 1), (32, 64, 64, 64, 128, 1),
              (32, 64, 64, 256, 128, 1), (32, 64, 64, 128, 256, 1), (32, 64, 64, 128, 128, 1), (32, 32, 32, 128, 128, 3),
              (32, 32, 32, 64, 64, 3), (32, 16, 16, 64, 64, 3), (32, 128, 128, 32, 32, 3)]
    print("%-28s %10s %10s %10s %10s %10s   GF" % ("B,H,W,Cin,Cout,k", "split", "conv_tc", "ss_fused", "ts_fused", "g_fused"))
    for (B, H, W, Cin, Cout, k) in shapes:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
        w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * 0.03
        mean = torch.zeros(Cin, device="cuda"); scale = torch.ones(Cin, device="cuda"); shift = torch.zeros(Cin, device="cuda")
        w_hi, w_lo = ops.weight_prep(w)
        a_hi = torch.empty_like(x); a_lo = torch.empty_like(x)
        y = torch.empty(B, H, W, Cout, device="cuda")
        t_split = timed(lambda: ops.affine_act_split(x, scale, shift, True, out_hi=a_hi, out_lo=a_lo, mean=mean), flush)
        t_conv = timed(lambda: ops.conv2d_tc(a_hi, a_lo, w_hi, w_lo, k, out=y), flush)
        t_ss = timed(lambda: ops.conv2d_tc_fused(x, w_hi, w_lo, k, mean=mean, scale=scale, shift=shift, relu=True, out=y,
                                                 impl="ss"), flush)
        t_ts = timed(lambda: ops.conv2d_tc_fused(x, w_hi, w_lo, k, mean=mean, scale=scale, shift=shift, relu=True, out=y,
                                                 impl="ts"), flush)
        t_g = timed(lambda: ops.conv2d_tc_fused(x, w_hi, w_lo, k, mean=mean, scale=scale, shift=shift, relu=True, out=y,
                                                impl="g"), flush)
        res = torch.randn(B, H, W, Cout, device="cuda", generator=g)
        bias = torch.randn(Cout, device="cuda", generator=g)
        t_res = timed(lambda: ops.conv2d_tc_fused(x, w_hi, w_lo, k, mean=mean, scale=scale, shift=shift, relu=True, out=y,
                                                  bias=bias, residual=res, impl="ts"), flush)
        gf = 2.0 * B * H * W * Cin * Cout * k * k / 1e9
        print("%-28s %10.1f %10.1f %10.1f %10.1f %10.1f   %.1f   ts+bias+residual %.1f" % (
            str((B, H, W, Cin, Cout, k)), t_split, t_conv, t_ss, t_ts, t_g, gf, t_res), flush=True)


if __name__ == "__main__":
    main()
