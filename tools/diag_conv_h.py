"""Correctness diagnostics + timing of the generation-5 convolution (csrc/conv_tc5.cu) against torch fp32 and the TS
kernel (csrc/conv_tc3.cu). Never stops at the first failure: one gpurun call reports every shape / mode.

    python tools/diag_conv_h.py            # errors (small batches) then timings (B=32, L2 flushed, CUDA events)
"""
import os
import sys
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def relerr(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def timed(fn, flush, iters=6):
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


def localize(y, ref, shape):
    """Where is the error? per-channel-block / per-pixel-position summaries of |y - ref| (NCHW tensors)."""
    d = (y - ref).abs()
    B, C, H, W = d.shape
    per_c = d.amax(dim=(0, 2, 3))
    per_h = d.amax(dim=(0, 1, 3))
    per_w = d.amax(dim=(0, 1, 2))
    per_b = d.amax(dim=(1, 2, 3))
    top = lambda t: [(int(i), float(t[i])) for i in torch.argsort(t, descending=True)[:4]]
    return "    worst channels %s | rows %s | cols %s | images %s | zero-out frac %.3f nan %d" % (
        top(per_c), top(per_h), top(per_w), top(per_b), float((y == 0).float().mean()), int(torch.isnan(y).sum()))


def ablate():
    """Which stage of the pipeline bounds conv_tc_h? Times the kernel with parts of the work switched off
    (FPD_CONV_DBG bit mask: 1 no MMA, 2 no weight TMA, 4 no x TMA, 8 no transform/copy, 16 no epilogue traffic,
    32 no halo split). Outputs are garbage under a non-zero mask; only the times mean something."""
    import fpd_b200  # noqa: F401
    from fpd_b200 import ops
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    shapes = [(32, 64, 64, 128, 128, 3), (32, 64, 64, 64, 64, 3), (32, 64, 64, 256, 128, 1), (32, 64, 64, 128, 256, 1),
              (32, 64, 64, 64, 128, 1), (32, 32, 32, 128, 128, 3)]
    masks = [0, 1, 2, 4, 8, 16, 32, 1 | 2, 1 | 8, 2 | 8, 1 | 2 | 8, 1 | 2 | 4 | 8, 1 | 2 | 4 | 8 | 16, 2 | 4 | 8 | 16 | 32,
             1 | 16, 2 | 4]
    print("ablation (us): mask bits 1=noMMA 2=noW 4=noX 8=noXform 16=noEpi 32=noSplit", flush=True)
    print("%-30s %-5s " % ("shape", "mode") + " ".join("%6d" % m for m in masks), flush=True)
    for (B, H, W, Cin, Cout, k) in shapes:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
        w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * 0.03
        mean = torch.zeros(Cin, device="cuda"); scale = torch.ones(Cin, device="cuda"); shift = torch.zeros(Cin, device="cuda")
        res = torch.randn(B, H, W, Cout, device="cuda", generator=g)
        bias = torch.randn(Cout, device="cuda", generator=g)
        y = torch.empty(B, H, W, Cout, device="cuda")
        kw = dict(mean=mean, scale=scale, shift=shift, relu=True, out=y, bias=bias, residual=res)
        for mode, prep in (("tf32", ops.weight_prep), ("f16", ops.weight_prep_f16)):
            w_hi, w_lo = prep(w)
            row = []
            for m in masks:
                os.environ["FPD_CONV_DBG"] = str(m)
                row.append(timed(lambda: ops.conv2d_tc_h(x, w_hi, w_lo, k, **kw), flush, iters=4))
            os.environ["FPD_CONV_DBG"] = "0"
            print("%-30s %-5s " % (str((B, H, W, Cin, Cout, k)), mode) + " ".join("%6.1f" % t for t in row), flush=True)
    return 0


def stalls():
    """Per-role stall cycles of conv_tc_h (fpd_conv2d_tc_h_set_profile_buffer): where does each warp role wait?"""
    import ctypes
    import fpd_b200  # noqa: F401
    from fpd_b200 import ops
    lib = ops.N.lib()
    shapes = [(32, 64, 64, 128, 128, 3), (32, 64, 64, 64, 64, 3), (32, 64, 64, 256, 128, 1), (32, 64, 64, 128, 256, 1),
              (32, 64, 64, 64, 128, 1), (32, 16, 16, 128, 128, 3), (32, 4, 4, 64, 64, 3)]
    names = ["P:w_empty", "P:raw_empty", "P:total", "M:w_full", "M:a_ready", "M:tmem_empty", "M:total", "E:tmem_full",
             "E:total", "X:raw_full", "X:a_empty", "X:barrier", "X:total"]
    print("per-role stall cycles, mean over CTAs, in us at 1.965 GHz (P producer, M MMA issuer, E epilogue, X transform)")
    for (B, H, W, Cin, Cout, k) in shapes:
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
        w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * 0.03
        mean = torch.zeros(Cin, device="cuda"); scale = torch.ones(Cin, device="cuda"); shift = torch.zeros(Cin, device="cuda")
        res = torch.randn(B, H, W, Cout, device="cuda", generator=g)
        bias = torch.randn(Cout, device="cuda", generator=g)
        y = torch.empty(B, H, W, Cout, device="cuda")
        kw = dict(mean=mean, scale=scale, shift=shift, relu=True, out=y, bias=bias, residual=res)
        for mode, prep in (("tf32", ops.weight_prep), ("f16", ops.weight_prep_f16)):
            if not lib.fpd_conv2d_tc_h_supported(Cin, Cout, k, H, W, int(mode == "f16")):
                continue
            w_hi, w_lo = prep(w)
            for dbg in (0, 1, 16):
                os.environ["FPD_CONV_DBG"] = str(dbg)
                for _ in range(2):
                    ops.conv2d_tc_h(x, w_hi, w_lo, k, **kw)
                buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
                lib.fpd_conv2d_tc_h_set_profile_buffer(ctypes.c_void_p(buf.data_ptr()))
                ops.conv2d_tc_h(x, w_hi, w_lo, k, **kw)
                torch.cuda.synchronize()
                lib.fpd_conv2d_tc_h_set_profile_buffer(None)
                os.environ["FPD_CONV_DBG"] = "0"
                m = buf.view(148, 16).double()
                used = m[:, 6] > 0
                avg = m[used].mean(dim=0) / 1965.0
                print("%-28s %-4s dbg=%-2d " % (str((B, H, W, Cin, Cout, k)), mode, dbg) +
                      " ".join("%s=%.1f" % (n, avg[i]) for i, n in enumerate(names)), flush=True)
    return 0


def main():
    if "--ablate" in sys.argv:
        return ablate()
    if "--stalls" in sys.argv:
        return stalls()
    import fpd_b200  # noqa: F401
    from fpd_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    lib = ops.N.lib()
    shapes = [(2, 64, 64, 64, 64, 1), (2, 64, 64, 64, 64, 3), (2, 64, 64, 128, 128, 3), (2, 64, 64, 128, 64, 1),
              (2, 64, 64, 64, 128, 1), (1, 64, 64, 256, 256, 1), (4, 32, 32, 64, 64, 3), (4, 16, 16, 64, 64, 3),
              (8, 8, 8, 64, 64, 3), (8, 4, 4, 64, 64, 3), (2, 4, 4, 64, 64, 3), (3, 8, 8, 128, 64, 1),
              (2, 128, 128, 32, 32, 3), (2, 64, 64, 128, 16, 1), (2, 64, 64, 16, 128, 1), (2, 64, 48, 32, 32, 3),
              (2, 32, 24, 48, 96, 3), (2, 16, 12, 128, 128, 3), (2, 64, 64, 160, 32, 1), (5, 16, 16, 256, 128, 3)]
    nfail = 0
    for f16 in (False, True):
        for passes in (3, 1):
            for (B, H, W, Cin, Cout, k) in shapes:
                tag = "%s passes=%d %s" % ("f16 " if f16 else "tf32", passes, (B, H, W, Cin, Cout, k))
                if not lib.fpd_conv2d_tc_h_supported(Cin, Cout, k, H, W, int(f16)):
                    print("SKIP  %s (unsupported)" % tag, flush=True)
                    continue
                try:
                    g = torch.Generator(device="cuda").manual_seed(7)
                    x = torch.randn(B, Cin, H, W, device="cuda", generator=g) * 2 + 0.5
                    mean = torch.randn(Cin, device="cuda", generator=g)
                    scale = torch.rand(Cin, device="cuda", generator=g) + 0.5
                    shift = torch.randn(Cin, device="cuda", generator=g) * 0.5
                    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * (1.0 / (Cin * k * k) ** 0.5)
                    bias = torch.randn(Cout, device="cuda", generator=g)
                    res = torch.randn(B, Cout, H, W, device="cuda", generator=g)
                    a = F.relu((x - mean.view(1, -1, 1, 1)) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
                    ref = F.conv2d(a, w, bias, padding=k // 2) + res
                    prep = ops.weight_prep_f16 if f16 else ops.weight_prep
                    w_hi, w_lo = prep(w, split=(passes == 3))
                    xh = x.permute(0, 2, 3, 1).contiguous()
                    y = ops.conv2d_tc_h(xh, w_hi, w_lo, k, mean=mean, scale=scale, shift=shift, relu=True, bias=bias,
                                        residual=res.permute(0, 2, 3, 1).contiguous())
                    torch.cuda.synchronize()
                    yn = y.permute(0, 3, 1, 2)
                    e = relerr(yn, ref)
                    tol = 2e-5 if passes == 3 else 3e-3
                    ok = e < tol
                    print("%s %s rel err %.3e" % ("ok   " if ok else "FAIL ", tag, e), flush=True)
                    if not ok:
                        nfail += 1
                        print(localize(yn - res - bias.view(1, -1, 1, 1), ref - res - bias.view(1, -1, 1, 1), None), flush=True)
                except Exception:
                    nfail += 1
                    print("EXC   %s\n%s" % (tag, traceback.format_exc()), flush=True)
                    try:
                        torch.cuda.synchronize()
                    except Exception:
                        print("device unusable after the exception; stopping", flush=True)
                        return 1
    print("failures: %d" % nfail, flush=True)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tshapes = [(32, 64, 64, 128, 128, 3), (32, 64, 64, 64, 64, 3), (32, 64, 64, 128, 64, 1), (32, 64, 64, 64, 128, 1),
               (32, 64, 64, 256, 128, 1), (32, 64, 64, 128, 256, 1), (32, 64, 64, 256, 256, 1), (32, 64, 64, 128, 128, 1),
               (32, 32, 32, 128, 128, 3), (32, 32, 32, 64, 64, 3), (32, 32, 32, 256, 128, 1), (32, 16, 16, 128, 128, 3),
               (32, 16, 16, 64, 64, 3), (32, 8, 8, 128, 128, 3), (32, 4, 4, 128, 128, 3), (32, 128, 128, 32, 32, 3),
               (32, 128, 128, 64, 128, 1)]
    print("%-30s %9s %9s %9s   %s" % ("B,H,W,Cin,Cout,k", "ts(us)", "h_tf32", "h_f16", "GF   f16: TFLOP/s(alg)"), flush=True)
    for (B, H, W, Cin, Cout, k) in tshapes:
        try:
            g = torch.Generator(device="cuda").manual_seed(0)
            x = torch.randn(B, H, W, Cin, device="cuda", generator=g)
            w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) * 0.03
            mean = torch.zeros(Cin, device="cuda"); scale = torch.ones(Cin, device="cuda"); shift = torch.zeros(Cin, device="cuda")
            res = torch.randn(B, H, W, Cout, device="cuda", generator=g)
            bias = torch.randn(Cout, device="cuda", generator=g)
            w_hi, w_lo = ops.weight_prep(w)
            h_hi, h_lo = ops.weight_prep_f16(w)
            y = torch.empty(B, H, W, Cout, device="cuda")
            kw = dict(mean=mean, scale=scale, shift=shift, relu=True, out=y, bias=bias, residual=res)
            t_ts = timed(lambda: ops.conv2d_tc_fused(x, w_hi, w_lo, k, **kw), flush)
            t_h = t_f = float("nan")
            if lib.fpd_conv2d_tc_h_supported(Cin, Cout, k, H, W, 0):
                t_h = timed(lambda: ops.conv2d_tc_h(x, w_hi, w_lo, k, **kw), flush)
            if lib.fpd_conv2d_tc_h_supported(Cin, Cout, k, H, W, 1):
                t_f = timed(lambda: ops.conv2d_tc_h(x, h_hi, h_lo, k, **kw), flush)
            gf = 2.0 * B * H * W * Cin * Cout * k * k / 1e9
            print("%-30s %9.1f %9.1f %9.1f   %.1f  %.0f" % (str((B, H, W, Cin, Cout, k)), t_ts, t_h, t_f, gf,
                                                         gf / t_f * 1e-3 if t_f == t_f else 0), flush=True)
        except Exception:
            print("EXC timing %s\n%s" % ((B, H, W, Cin, Cout, k), traceback.format_exc()), flush=True)
            return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
