#!/usr/bin/env python
"""Per-kernel histogram of the SASS opcodes that show what hardware path a kernel uses (tcgen05 = UTC*MMA, TMEM loads =
LDTM, TMA = UTMALDG / UTMASTG, legacy warp MMA = HMMA, packed fp32 = FFMA2 / FADD2 / FMUL2, warp reductions = REDUX).
The .so is git-ignored, so this listing is the committed evidence:  python scripts/sass_histogram.py > profiles/rNN_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ldt_b200", "csrc", "libldt_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "FFMA2",
         "FADD2", "FMUL2", "FFMA", "MUFU", "REDUX", "LDGSTS", "ATOMS", "FMNMX3", "FMNMX"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    out = subprocess.run(["c++filt"], input=out, capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (.*)$", line)
        if m:
            cur = re.sub(r"\(.*$", "", m.group(1).strip())
            kernels.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur:
            op, mods = m.group(1), m.group(2)
            kernels[cur]["__total__"] += 1
            for w in WATCH:
                if op == w:
                    key = w
                    if w in ("UTCHMMA", "UTCQMMA", "UTCOMMA") and ".2CTA" in mods:
                        key += ".2CTA"
                    if w == "UTCBAR" and ".MULTICAST" in mods:
                        key += ".MULTICAST"
                    if w == "HMMA":
                        key += mods.split(".F32")[0] if ".F32" in mods else mods
                    kernels[cur][key] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: watched SASS opcodes per kernel (cuobjdump -sass, CUDA {os.environ.get('CUDA_VERSION', '12.9')})")
    tot = collections.Counter()
    for k, c in kernels.items():
        watched = {a: b for a, b in c.items() if a != "__total__"}
        tot.update(watched)
        if not watched:
            continue
        print(f"{k}\n    instructions {c['__total__']}: " + "  ".join(f"{a} {b}" for a, b in sorted(watched.items())))
    print("# library total: " + "  ".join(f"{a} {b}" for a, b in sorted(tot.items())))


if __name__ == "__main__":
    sys.exit(main())
