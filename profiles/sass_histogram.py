#!/usr/bin/env python
"""SASS opcode histogram of the shipped library, per kernel: the Blackwell-native evidence the profiling guide asks for
(tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, bulk copies -> UBLKCP, commit -> UTCBAR, mbarrier -> SYNCS).

    python profiles/sass_histogram.py [path/to/libcql_b200.so] > profiles/r02_sass_histogram.txt
"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

LIB = Path(sys.argv[1]) if len(sys.argv) > 1 else Path(__file__).resolve().parent.parent / "replay_cql_b200" / "libcql_b200.so"
OPS = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR.2CTA.MULTICAST", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTCATOMSWS", "SYNCS", "UCGABAR",
       "F2FP", "FFMA2", "HMMA", "CCTL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
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
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            kernels[cur]["_total"] += 1
            op = m.group(1)
            for want in OPS:
                if op == want or op.startswith(want + "."):
                    kernels[cur][want] += 1
                    break
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {LIB.name}: opcode counts per kernel (static instruction counts; UTCHMMA.2CTA = tcgen05.mma.cta_group::2)")
    print(f"# {'kernel':88s} {'instrs':>7s} " + " ".join(f"{o:>9s}" for o in OPS))
    tot = Counter()
    for (name, c), dm in zip(kernels.items(), demangle):
        if not any(c[o] for o in OPS[:8]):
            continue
        short = re.sub(r"\(.*", "", dm).replace("void ", "").replace("cql::", "")
        print(f"  {short[:88]:88s} {c['_total']:7d} " + " ".join(f"{c[o]:9d}" for o in OPS))
        tot.update(c)
    print(f"  {'TOTAL (kernels with tensor-core / TMEM / bulk-copy instructions)':88s} {tot['_total']:7d} " + " ".join(f"{tot[o]:9d}" for o in OPS))


if __name__ == "__main__":
    main()
