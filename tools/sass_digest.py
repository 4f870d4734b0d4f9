#!/usr/bin/env python
"""Instruction histogram per kernel of the built library (cuobjdump -sass): which kernels carry tcgen05 (UTCIMMA / UTCHMMA),
TMEM loads (LDTM), TMA (UTMALDG / UTMASTG), mbarrier (SYNCS), FP64 tensor (DMMA), FP64 ALU and POPC instructions.
usage: python tools/sass_digest.py > profiles/r02_sass_digest.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "everest_b200", "lib", "libeverest_b200.so")
KEYS = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "DMNMX",
        "DSETP", "MUFU", "POPC", "IMAD", "LDG", "STG", "LDS", "STS", "LDGSTS", "BAR", "SHFL", "ATOM", "RED"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    funcs = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and cur:
            funcs[cur][m.group(1).split(".")[0]] += 1
    demangle = subprocess.run(["c++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS digest of everest_b200/lib/libeverest_b200.so (sm_100a), HEAD {head}; columns: total instructions, then the counts of")
    print("# " + " ".join(KEYS))
    tot = collections.Counter()
    for (name, cnt), dn in zip(funcs.items(), demangle):
        total = sum(cnt.values())
        for k in KEYS:
            tot[k] += cnt.get(k, 0)
        short = re.sub(r"\(.*", "", dn)[:70]
        cols = " ".join(f"{k}={cnt[k]}" for k in KEYS if cnt.get(k))
        print(f"{short:70s} {total:7d}  {cols}")
    print("# library totals: " + " ".join(f"{k}={tot[k]}" for k in KEYS if tot[k]))


if __name__ == "__main__":
    main()
