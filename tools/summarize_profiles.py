"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py r1      # reads gpurun_out/launches_r1.csv, gpurun_out/prof_*.ncu-rep
"""
import collections
import csv
import gzip
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        # tcgen05 (UTC*MMA) work shows up in the tensor-pipe cycle counters, not in the legacy HMMA sub-pipe one
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed"]


def launches(tag):
    path = os.path.join(SRC, "launches_%s.csv" % tag)
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in data:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(CUtensorMap.*|\(float.*|\(fpd::.*|\(double.*|\(int.*", "", r[ix["Kernel Name"]]).replace("fpd::<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += float(r[ix["Metric Value"]].replace(",", "")) / 1000.0
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, "%s_launches_summary.md" % tag), "w") as f:
        f.write("# %s: ncu launch list of ONE eager FPD training step (B=32, student s4 f128 + teacher s8 f256)\n\n" % tag)
        f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv "
                "--log-file gpurun_out/launches_%s.csv python tools/profile_step.py --ncu` (cold-cache, serialised: "
                "compare SHARES, not absolutes). Raw list: `%s_launches.csv.gz`.\n\n" % (tag, tag))
        f.write("%d kernel launches, %.1f ms summed device time.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n"
                % (sum(v[0] for v in agg.values()), tot / 1000.0))
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k[-90:], v[0], v[1], 100 * v[1] / tot))
    with open(path, "rb") as fi, gzip.open(os.path.join(OUT, "%s_launches.csv.gz" % tag), "wb") as fo:
        shutil.copyfileobj(fi, fo)


def report(tag, name, title):
    rep = os.path.join(SRC, name + ".ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(os.path.join(OUT, "%s_%s.md" % (tag, name)), "w") as f:
        f.write("# %s: `ncu --set full --clock-control none --import-source on` -- %s\n\n" % (tag, title))
        f.write("| metric | value | unit |\n|---|---:|---|\n")
        for i, h in enumerate(hdr):
            base = h.split(".TriageCompute.")[-1]
            if base in KEYS or h in KEYS or h.startswith("sm__pipe_tensor") or "pipe_tensor_op" in h:
                f.write("| `%s` | %s | %s |\n" % (h, vals[i], units[i]))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    report(tag, "prof_conv_tc_ss", "conv_tc_kernel (SS, pre-split operands), 3x3 128->128 @64x64, B=32, 3xTF32")
    report(tag, "prof_conv_tc_ts", "conv_tc_ts_kernel (fused transform, A in TMEM), 3x3 128->128 @64x64, B=32, 3xTF32")
    report(tag, "prof_wgrad_tc_fused", "wgrad_tc_fused_kernel, 3x3 64->64 @64x64, B=32, 3xTF32")
    report(tag, "prof_conv_h_3x3", "conv_tc_h_kernel<f16> (halo reuse, 3xFP16, A in TMEM), 3x3 128->128 @64x64, B=32")
    report(tag, "prof_conv_h_1x1", "conv_tc_h_kernel<f16> (3xFP16, A in TMEM), 1x1 128->256 + bias + residual @64x64, B=32")
    report(tag, "prof_wgrad_tc3", "wgrad_tc3_kernel (halo tile, taps as shifted start rows), 3x3 64->64 @64x64, B=32, 3xTF32")
    report(tag, "prof_wgrad_tc3_pair", "wgrad_tc3_kernel, pair mode (Cin = 32: two taps per M = 64 MMA), 3x3 32->32 @64x48, B=24, 3xTF32")
    # launch-weighted roofline of the dominant kernel over one whole step: conv FLOPs it executes / its summed duration
    path = os.path.join(SRC, "launches_%s.csv" % tag)
    if os.path.exists(path):
        rows = list(csv.reader(open(path)))
        hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
        ix = {h: i for i, h in enumerate(rows[hi])}
        us = sum(float(r[ix["Metric Value"]].replace(",", "")) / 1000.0 for r in rows[hi + 1:]
                 if len(r) > ix["Metric Value"] and "conv_tc_h_kernel" in r[ix["Kernel Name"]])
        flop = 32 * (7.8145e9 + 7.8145e9 + 56.189e9)      # student fwd + student dgrad + teacher fwd
        peak = 1640.6e12
        with open(os.path.join(OUT, "%s_launches_summary.md" % tag), "a") as f:
            f.write("\nLaunch-weighted roofline of `conv_tc_h_kernel`: %.3f TFLOP of forward + data-gradient convolutions in "
                    "%.2f ms summed = %.1f TFLOP/s = **%.3f** of the measured f16 peak (1640.6 TFLOP/s; cap 0.333 for three "
                    "passes).\n" % (flop / 1e12, us / 1000.0, flop / (us * 1e-6) / 1e12, flop / (us * 1e-6) / peak))
    cup = os.path.join(SRC, "step_cupti_%s.txt" % tag)
    if os.path.exists(cup):
        shutil.copy(cup, os.path.join(OUT, "%s_step_cupti.txt" % tag))
    print(sorted(os.listdir(OUT)))
