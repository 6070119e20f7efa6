"""Per-kernel SASS evidence that the hot kernels are Blackwell-native: counts of the tcgen05 / TMEM / TMA mnemonics in every
kernel of libfpd_b200.so (cuobjdump -sass; runs without a GPU).

    python tools/sass_summary.py > profiles/sass_summary.txt

UTCHMMA = tcgen05.mma kind::f16/tf32 (UTCQMMA/UTCOMMA: fp8/fp4 kinds, unused here), LDTM / STTM = tcgen05.ld / .st (TMEM),
UTMALDG / UTMASTG = TMA bulk-tensor load / store, UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "fast-human-pose-estimation.pytorch_b200", "libfpd_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCATOMSWS", "SYNCS",
             "HMMA", "FFMA", "DFMA", "DADD", "SHFL", "LDG", "STG", "LDS", "STS", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    counts[cur][mn] += 1
            counts[cur]["_total"] += 1
    names = list(counts)
    try:
        dm = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, dm))
    except Exception:
        pass
    cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "FFMA", "DFMA", "LDG", "STG", "LDS", "STS"]
    print("# SASS mnemonic counts per kernel of libfpd_b200.so (sm_100a); tool: tools/sass_summary.py")
    print("%-86s %7s " % ("kernel", "instrs") + " ".join("%7s" % c for c in cols))
    tot = collections.Counter()
    for fn in names:
        c = counts[fn]
        nm = demangle.get(fn, fn)
        nm = re.sub(r"fpd::\(anonymous namespace\)::", "", nm)
        nm = re.sub(r"\(.*", "", nm)[:86]
        print("%-86s %7d " % (nm, c["_total"]) + " ".join("%7d" % c[k] for k in cols))
        tot.update(c)
    print("%-86s %7d " % ("TOTAL", tot["_total"]) + " ".join("%7d" % tot[k] for k in cols))
    tc = [demangle.get(f, f) for f in names if counts[f]["UTCHMMA"]]
    print("# kernels issuing tcgen05.mma (UTCHMMA): %d; with TMEM loads/stores: %d; with TMA loads: %d; with TMA stores: %d" % (
        len(tc), sum(1 for f in names if counts[f]["LDTM"] or counts[f]["STTM"]), sum(1 for f in names if counts[f]["UTMALDG"]),
        sum(1 for f in names if counts[f]["UTMASTG"])))


if __name__ == "__main__":
    main()
