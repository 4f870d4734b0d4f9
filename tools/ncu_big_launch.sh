#!/bin/bash
# usage: ncu_big_launch.sh <kernel regex> <out name> -- finds the launch with the largest grid among the first 60 launches of
# the kernel inside tools/probe_ozaki.py full and captures that one with --set full
K=$1; OUT=$2
ncu --metrics launch__grid_size,gpu__time_duration.sum --clock-control none -k regex:$K -c 60 --csv --log-file /tmp/list_$OUT.csv \
    ${PROBE:-python tools/probe_ozaki.py full} > /dev/null 2>&1
IDX=$(python - <<PY
import csv
rows=list(csv.reader(open("/tmp/list_$OUT.csv")))
h=next(i for i,r in enumerate(rows) if r and r[0]=="ID")
ix={k:j for j,k in enumerate(rows[h])}
best=(-1,-1); ids=[]
for r in rows[h+1:]:
    if len(r)<len(rows[h]) or r[ix["Metric Name"]]!="launch__grid_size": continue
    ids.append(float(r[ix["Metric Value"]].replace(",","")))
m=max(ids); i=[k for k,v in enumerate(ids) if v==m]
print(i[-1])
PY
)
echo "largest-grid launch index of $K: $IDX"
ncu --set full --clock-control none --import-source on -k regex:$K -s $IDX -c 1 -o gpurun_out/$OUT -f ${PROBE:-python tools/probe_ozaki.py full} > /dev/null 2>&1
ls -la gpurun_out/$OUT.ncu-rep
