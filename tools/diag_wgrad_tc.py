"""Per-shape subprocess diagnosis of the tcgen05 weight-gradient kernel (see diag_conv_tc.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHAPES = [(1, 8, 8, 128, 32, 1), (1, 8, 8, 128, 64, 1), (2, 64, 64, 128, 64, 1), (1, 8, 8, 64, 128, 1),
          (1, 8, 8, 64, 64, 3), (2, 64, 64, 64, 64, 3), (2, 128, 128, 32, 32, 3), (1, 32, 32, 128, 128, 3)]


def run_one(B, H, W, Cin, Cout, k, passes):
    import torch
    import torch.nn.functional as F
    from fpd_b200 import ops as o
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(B, Cin, H, W, device="cuda", generator=g)
    dy = torch.randn(B, Cout, H, W, device="cuda", generator=g)
    w = torch.zeros(Cout, Cin, k, k, device="cuda", requires_grad=True)
    (ref,) = torch.autograd.grad(F.conv2d(a, w, None, padding=k // 2), w, dy)
    nh = lambda t: t.permute(0, 2, 3, 1).contiguous()
    a_hi, a_lo = o.affine_act_split(nh(a), split=(passes == 3))
    g_hi, g_lo = o.affine_act_split(nh(dy), split=(passes == 3))
    dw = o.conv2d_wgrad_tc(a_hi, a_lo, g_hi, g_lo, k)
    torch.cuda.synchronize()
    err = ((dw - ref).abs().max() / ref.abs().max()).item()
    print("wgrad shape", (B, H, W, Cin, Cout, k), "passes", passes, "rel_err %.3e" % err, flush=True)
    if not err < (1e-4 if passes == 3 else 5e-3):
        d = (dw - ref).abs()
        print("  per-tap max err", d.amax(dim=(0, 1)).flatten().tolist())
        print("  bad co count", int((d.amax(dim=(1, 2, 3)) > 1e-2 * ref.abs().max()).sum()), "bad ci count",
              int((d.amax(dim=(0, 2, 3)) > 1e-2 * ref.abs().max()).sum()))
        print("  dw[0,0:4,0,0]", dw[0, :4, 0, 0].tolist(), "ref", ref[0, :4, 0, 0].tolist())
        print("  dw[0:4,0,0,0]", dw[:4, 0, 0, 0].tolist(), "ref", ref[:4, 0, 0, 0].tolist())
        print("  corr", torch.corrcoef(torch.stack([dw.flatten(), ref.flatten()]))[0, 1].item())
        # is dw a transposed / permuted version?
        if k == 1 and Cin == Cout:
            print("  err vs transposed", ((dw.transpose(0, 1) - ref).abs().max() / ref.abs().max()).item())


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_one(*map(int, sys.argv[1:]))
    else:
        for s in SHAPES:
            for passes in (1, 3):
                r = subprocess.run([sys.executable, __file__, *map(str, s), str(passes)], capture_output=True,
                                   text=True, timeout=300)
                sys.stdout.write(r.stdout)
                if r.returncode != 0:
                    print("wgrad shape", s, "passes", passes, "FAILED rc", r.returncode, r.stderr[-600:])
