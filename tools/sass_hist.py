#!/usr/bin/env python
"""SASS opcode histogram per kernel of the built library (runs without a GPU):

    python tools/sass_hist.py > profiles/r02_sass_opcodes.txt

Counts the Blackwell-specific mnemonics per kernel: UTCHMMA(.2CTA) = tcgen05.mma (CTA pairs), LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit, FFMA2 / FMUL2 / FADD2 = packed fp32,
LDG.E.128 = 128-bit global loads, plus the size of each kernel."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "news_recommendation_project_v2_b200", "libnrb200.so")
WATCH = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "FFMA2",
         "FMUL2", "FADD2", "MUFU", "LDG.E.128", "STG.E.128", "LDS.128", "STS.128", "SHFL", "REDG", "FFMA", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=900).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    kernels[cur][w] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass libnrb200.so (sm_100a): opcode counts per kernel (static instruction counts)")
    print("# UTCHMMA = tcgen05.mma, .2CTA = cta_group::2; LDTM/STTM = tcgen05.ld/st; UTMALDG/UTMASTG = TMA load/store;")
    print("# FFMA2/FMUL2/FADD2 = packed fp32 pairs\n")
    for (name, cnt), nice in zip(kernels.items(), demangle):
        nice = re.sub(r"\(.*", "", nice)
        if not nice.startswith(("void nrb::", "nrb::")):
            continue
        hot = " ".join(f"{w}={cnt[w]}" for w in WATCH if cnt[w])
        print(f"{nice[:90]:90s} instrs={cnt['_total']:5d}  {hot}")


if __name__ == "__main__":
    sys.exit(main())
