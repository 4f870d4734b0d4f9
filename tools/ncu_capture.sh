#!/bin/bash
# usage: tools/ncu_capture.sh <workload> <kernel regex> <out name> [skip] [count]
# plain run first (must exit 0), then ONE launch of the kernel under `ncu --set full` (B200_PROFILING.md recipe);
# `skip` = matching launches to pass over first (set-up launches of the same kernel, warm-up screens)
W=$1; K=$2; OUT=$3; SKIP=${4:-2}; CNT=${5:-1}
python tools/probe_forward.py $W > gpurun_out/${OUT}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c $CNT -f -o gpurun_out/$OUT python tools/probe_forward.py $W > gpurun_out/${OUT}_ncu.log 2>&1
echo "$OUT rc=$?"; tail -2 gpurun_out/${OUT}_ncu.log
