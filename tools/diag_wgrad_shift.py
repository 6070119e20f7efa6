"""Experiment: are the TMA SWIZZLE_128B_ATOM_32B write pattern and the UMMA SWIZZLE_128B_BASE32B read pattern both keyed
on absolute shared-memory address bits? Runs the fused weight-gradient kernel with every operand tile shifted 0..3 rows
off the swizzle-atom boundary (FPD_WGRAD_DBG 16/32/48) and reports the error against torch (identity pre-op, so the
transform is position-independent)."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fpd_b200  # noqa: F401
    from fpd_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    for (B, H, W, Cin, Cout, k) in [(2, 32, 32, 64, 64, 1), (2, 32, 32, 64, 64, 3), (2, 64, 64, 128, 64, 1)]:
        g = torch.Generator(device="cuda").manual_seed(3)
        x = torch.randn(B, Cin, H, W, device="cuda", generator=g)
        dy = torch.randn(B, Cout, H, W, device="cuda", generator=g)
        ref = torch.nn.grad.conv2d_weight(x, (Cout, Cin, k, k), dy, padding=k // 2)
        xh = x.permute(0, 2, 3, 1).contiguous()
        dyh = dy.permute(0, 2, 3, 1).contiguous()
        for shift in (0, 1, 2, 3):
            os.environ["FPD_WGRAD_DBG"] = str(shift << 4)
            dw = ops.conv2d_wgrad_tc_fused(xh, dyh, k)
            torch.cuda.synchronize()
            err = ((dw - ref).abs().max() / ref.abs().max()).item()
            print("%s shift=%d rows: rel err %.3e %s" % ((B, H, W, Cin, Cout, k), shift, err, "OK" if err < 1e-4 else "MISMATCH"),
                  flush=True)
        os.environ["FPD_WGRAD_DBG"] = "0"


if __name__ == "__main__":
    main()
