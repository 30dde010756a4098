"""Opcode histogram per kernel of the shipped libunetb200.so (cuobjdump -sass), the SASS evidence that
the tcgen05 / TMEM / TMA path is what was built (B200_PROFILING.md "What proves a Blackwell-native
kernel"): UTCHMMA(.2CTA) = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG(.IM2COL) = TMA loads,
UTMASTG = TMA stores, UTCCP = tcgen05.cp, UTCBAR = tcgen05.commit, SYNCS = mbarrier.
    python scripts_dev/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unet_segmentation_b200", "lib", "libunetb200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMALDG.IM2COL", "UTMASTG", "UTCCP", "UTCBAR",
        "SYNCS", "HMMA", "STG.E.128", "LDG.E.128", "FFMA", "FFMA2", "DFMA", "ATOM", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            c = funcs[cur]
            c["_total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + "."):
                    c[k] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                c["UTCHMMA.2CTA"] += 1
            if op.startswith("UTMALDG") and "IM2COL" in op:
                c["UTMALDG.IM2COL"] += 1
    names = list(funcs)
    dem = subprocess.run(["cu++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {len(names)} kernels; opcode counts per kernel")
    print("# (columns with all zeros omitted per row; kernels without any listed opcode are summarised at the end)")
    tot = collections.Counter()
    plain = 0
    for n, d in zip(names, dem):
        c = funcs[n]
        for k in KEYS:
            tot[k] += c[k]
        shown = {k: c[k] for k in KEYS if c[k] and k not in ("FFMA", "FFMA2", "LDG.E.128", "STG.E.128", "ATOM", "RED", "DFMA")}
        short = (d.split(">(")[0] + ">") if ">(" in d else re.sub(r"\(.*", "", d)
        short = short.replace("ub::", "").replace("void ", "")
        if not shown:
            plain += 1
            continue
        extra = {k: c[k] for k in ("STG.E.128", "LDG.E.128") if c[k]}
        print(f"{short:58s} instr {c['_total']:6d}  " + "  ".join(f"{k} {v}" for k, v in {**shown, **extra}.items()))
    print(f"# {plain} further kernels are CUDA-core / elementwise only (no tensor, TMA or mbarrier opcode)")
    print("# library totals: " + "  ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))
    if tot["HMMA"]:
        print("# WARNING: legacy HMMA (mma.sync) present")


if __name__ == "__main__":
    sys.exit(main())
