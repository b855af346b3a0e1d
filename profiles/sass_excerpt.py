"""SASS evidence for the tcgen05 / TMEM / bulk-TMA kernels: per kernel of liblcn_b200.so the count of every Blackwell
tensor-memory / tensor-core / bulk-copy mnemonic and the first occurrences in context.
    python profiles/sass_excerpt.py > profiles/r2/sass_tcgen05.txt        (cuobjdump, no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lcn_pose_b200", "liblcn_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UBLKCP", "UBLKRED", "UBLKPF", "SYNCS", "UTMALDG", "UTMASTG",
             "ACQBULK", "UCGABAR", "ELECT", "HMMA", "FFMA", "R2UR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = re.split(r"\n\s*Function : ", sass)[1:]
    print("cuobjdump -sass lcn_pose_b200/liblcn_b200.so  (sm_100a); mnemonic counts per kernel, then excerpts\n")
    rows = []
    for k in kernels:
        name = k.split("\n", 1)[0].strip()
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
        cnt = collections.Counter()
        for line in k.splitlines():
            m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1).split(".")[0]
                if op in MNEMONICS:
                    cnt[op] += 1
        if any(cnt[x] for x in ("UTCHMMA", "LDTM", "UBLKCP", "UBLKRED", "UTCBAR")):
            rows.append((demangled, cnt, k))
    print("%-46s" % "kernel" + "".join("%9s" % m for m in MNEMONICS))
    for name, cnt, _ in rows:
        print("%-46s" % name[:46] + "".join("%9d" % cnt[m] for m in MNEMONICS))
    for name, cnt, k in rows:
        print("\n==== %s ====" % name)
        shown = collections.Counter()
        for line in k.splitlines():
            for m in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKRED"):
                if re.search(r"\b" + m + r"\b", line) and shown[m] < 3:
                    shown[m] += 1
                    print("   " + re.sub(r"\s+", " ", line.strip())[:150])


if __name__ == "__main__":
    main()
