"""Per-kernel counts of the SASS opcodes that prove a Blackwell-native kernel (B200_PROFILING.md: tcgen05.mma ->
UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG), read from the built shared library with cuobjdump.

    python tools/sass_summary.py > profiles/r02_sass_opcodes.txt
"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "embeddingnet_b200", "libembeddingnet_b200.so")
WANT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "HMMA", "SYNCS", "RED", "ATOM"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WANT:
                if op.startswith(w):
                    kernels[cur][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# SASS opcode counts per kernel, %s (cuobjdump -sass; sm_100a)" % os.path.relpath(LIB, ROOT))
    print("# tcgen05.mma -> UTC*MMA | tcgen05.ld/st -> LDTM/STTM | TMA (cp.async.bulk.tensor) -> UTMALDG | mbarrier -> SYNCS")
    print("%-9s %s" % ("instrs", "  ".join("%7s" % w for w in WANT)) + "  kernel")
    tot = Counter()
    for (name, c), dn in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("en::(anonymous namespace)::", "").replace("en::", "")
        print("%-9d %s  %s" % (c["_total"], "  ".join("%7d" % c[w] for w in WANT), short[:110]))
        tot.update(c)
    print("%-9d %s  TOTAL (%d kernels)" % (tot["_total"], "  ".join("%7d" % tot[w] for w in WANT), len(kernels)))


if __name__ == "__main__":
    sys.exit(main())
