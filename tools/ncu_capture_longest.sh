#!/bin/bash
# usage: tools/ncu_capture_longest.sh <workload> <kernel regex> <out name>
# Two passes over tools/probe_forward.py: list the launches of the kernel with their durations, then capture (ncu --set full)
# the first launch that is at least half as long as the longest one -- set-up launches of the same kernel come first and are
# tiny next to the screen's.
W=$1; K=$2; OUT=$3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$K --csv --log-file gpurun_out/${OUT}_list.csv python tools/probe_forward.py $W > /dev/null 2>&1
SKIP=$(python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${OUT}_list.csv")))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID"); hdr=rows[h]; ix={k:j for j,k in enumerate(hdr)}
d=[]
for r in rows[h+1:]:
    if len(r)<len(hdr) or r[ix["Metric Name"]]!="gpu__time_duration.sum": continue
    v=float(r[ix["Metric Value"]].replace(",","")); v*={"ns":1e-3,"us":1.0,"ms":1e3,"s":1e6}.get(r[ix["Metric Unit"]],1.0)
    d.append(v)
m=max(d); print(next(i for i,v in enumerate(d) if v>=0.5*m))
PY
)
echo "$OUT: capturing launch #$SKIP of $K"
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/$OUT python tools/probe_forward.py $W > gpurun_out/${OUT}_ncu.log 2>&1
echo "$OUT rc=$?"
