#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libmrssm_b200.so: counts of the mnemonics that prove the Blackwell paths
(UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = bulk copy,
UTMAPF / UBLKPF = TMA / bulk prefetch, SYNCS = mbarrier ops) per kernel.  No GPU needed (cuobjdump reads the cubin).

    python profiles/sass_histogram.py > profiles/rNN_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "multimodal-rssm_b200", "lib", "libmrssm_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UBLKPF", "SYNCS", "REDUX", "ATOMG", "RED", "STL", "LDL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["_total"] += 1
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    cur[o] += 1
    total = collections.Counter()
    print("%-64s %7s " % ("kernel", "instr") + " ".join("%7s" % o for o in OPS))
    for name, c in kernels.items():
        if not any(c[o] for o in OPS[:9]):
            continue
        print("%-64s %7d " % (name[:64], c["_total"]) + " ".join("%7d" % c[o] for o in OPS))
        total.update(c)
    print("%-64s %7d " % ("TOTAL (kernels listed)", total["_total"]) + " ".join("%7d" % total[o] for o in OPS))


if __name__ == "__main__":
    sys.exit(main())
