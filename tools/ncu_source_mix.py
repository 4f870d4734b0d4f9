"""Instruction mix of one kernel from an ncu report's source page: warp / thread instruction totals, average active lanes,
top opcodes, and the heaviest SASS regions by warp instructions and by stall samples.
usage: python tools/ncu_source_mix.py <report.ncu-rep> [top_n]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
name = rows[0][1]
h = rows[1]
ci = {k: h.index(k) for k in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
warp = thr = smp = 0
ops = collections.defaultdict(lambda: [0, 0, 0])
lines = []
for r in rows[2:]:
    if len(r) <= ci["Thread Instructions Executed"]:
        continue
    w, t, s = int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0), int(r[ci["# Samples"]] or 0)
    src = r[ci["Source"]].strip()
    op = src.split()[1] if src.startswith("@") else src.split()[0]
    op = op.split(".")[0]
    warp += w; thr += t; smp += s
    o = ops[op]; o[0] += w; o[1] += t; o[2] += s
    lines.append((w, t, s, src))
print(name)
print(f"warp instructions {warp:.4g}  thread instructions {thr:.4g}  avg active lanes {thr / max(warp, 1):.1f}  stall samples {smp}")
for op, (w, t, s) in sorted(ops.items(), key=lambda x: -x[1][0])[:top]:
    print(f"  {op:10s} {100.0 * w / warp:5.1f} % of warp instr   lanes {t / max(w, 1):5.1f}   {100.0 * s / max(smp, 1):5.1f} % of samples")
