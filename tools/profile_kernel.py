"""Run one libfpd_b200 kernel shape a few times (for `ncu --set full` captures of the hot kernels).

    ncu --set full --clock-control none --import-source on -k regex:conv_tc_h_kernel -s 2 -c 1 \
        -o gpurun_out/prof_conv_h python tools/profile_kernel.py conv_h_f16 32 64 64 128 128 3
    ncu --set full ... -k regex:wgrad_tc3_kernel ... python tools/profile_kernel.py wgrad_fused 32 64 64 64 64 3
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fpd_b200  # noqa: F401
    from fpd_b200 import ops
    kind = sys.argv[1]
    B, H, W, Cin, Cout, k = map(int, sys.argv[2:8])
    iters = int(sys.argv[8]) if len(sys.argv) > 8 else 4
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * 0.03
    if kind == "conv_ts":
        w_hi, w_lo = ops.weight_prep(w)
        y = torch.empty(B, H, W, Cout, device="cuda")
        mean = torch.zeros(Cin, device="cuda"); scale = torch.ones(Cin, device="cuda"); shift = torch.zeros(Cin, device="cuda")
        for _ in range(iters):
            ops.conv2d_tc_fused(x, w_hi, w_lo, k, mean=mean, scale=scale, shift=shift, relu=True, out=y)
    elif kind in ("conv_h_f16", "conv_h_tf32"):
        prep = ops.weight_prep_f16 if kind == "conv_h_f16" else ops.weight_prep
        w_hi, w_lo = prep(w)
        y = torch.empty(B, H, W, Cout, device="cuda")
        res = torch.randn(B, H, W, Cout, device="cuda", generator=g)
        bias = torch.randn(Cout, device="cuda", generator=g)
        mean = torch.zeros(Cin, device="cuda"); scale = torch.ones(Cin, device="cuda"); shift = torch.zeros(Cin, device="cuda")
        for _ in range(iters):
            ops.conv2d_tc_h(x, w_hi, w_lo, k, mean=mean, scale=scale, shift=shift, relu=True, out=y, bias=bias,
                            residual=res if k == 1 else None)
    elif kind == "wgrad_fused":
        dy = torch.randn(B, H, W, Cout, device="cuda", generator=g)
        for _ in range(iters):
            ops.conv2d_wgrad_tc_fused(x, dy, k, relu=True)
    else:
        raise SystemExit("unknown kind " + kind)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
