#!/usr/bin/env python
"""SASS mnemonic counts of the shipped library (whole library and per kernel) -> profiles/r<N>_sass_counts.txt.

usage: python tools/sass_counts.py [round]      (runs cuobjdump -sass here; no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kaldi-lstm_b200", "_lib", "liblstmp_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT",
         "ST.E.128", "LD.E.128", "STS.128", "LDS.128", "STG.E.128", "LDG.E.128", "FFMA2", "FFMA", "HMMA", "MUFU.EX2", "MUFU.TANH",
         "FENCE.VIEW.ASYNC", "MEMBAR", "RED.E", "ATOM"]


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    ins = re.compile(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = ins.match(line)
        if m and cur is not None:
            op = m.group(1)
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    # FFMA must not swallow FFMA2, ST.E.128 is a prefix match on the full mnemonic
                    if w == "FFMA" and op.startswith("FFMA2"):
                        continue
                    cur[w] += 1
                    break
            else:
                # generic 128-bit accesses are spelled ST.E.128 / LD.E.128 with optional qualifiers in between
                if re.match(r"^ST\.E(\.[A-Z0-9]+)*\.128", op):
                    cur["ST.E.128"] += 1
                elif re.match(r"^LD\.E(\.[A-Z0-9]+)*\.128", op):
                    cur["LD.E.128"] += 1
    total = collections.Counter()
    for c in per.values():
        total.update(c)
    out = os.path.join(ROOT, "profiles", "r%s_sass_counts.txt" % rnd)
    with open(out, "w") as f:
        f.write("SASS mnemonic counts of kaldi-lstm_b200/_lib/liblstmp_b200.so (cuobjdump -sass, sm_100a), round %s; made by "
                "tools/sass_counts.py\n" % rnd)
        f.write("UTCHMMA = tcgen05.mma kind::f16/tf32, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTMALDG = cp.async.bulk.tensor,\n"
                "LDGSTS = cp.async, SYNCS = mbarrier ops, UCGABAR = barrier.cluster, ST.E.128 / LD.E.128 = GENERIC 128-bit accesses "
                "(round 1: 1208 ST.E.128)\n\n")
        f.write("TOTAL %s\n\n" % dict(total))
        for name, c in per.items():
            f.write("%s\n    %s\n" % (name, dict(c)))
    print(out, dict(total))


if __name__ == "__main__":
    main()
