#!/bin/bash
# Session-4 experiment: cell pre-filter + single-point fast path of the MC/HVI tiled kernel (EVEREST_MC_FAST), then
# parity tests, plain bench and a full-set ncu capture of the kernel.
EVEREST_MC_FAST=0 python tools/probe_mc.py base
EVEREST_MC_FAST=1 python tools/probe_mc.py fast
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/s4c_bench_n1.json 2> gpurun_out/s4c_bench_n1.err || exit 1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:mc_hvi_tiled_kernel -s 4 -c 1 -o gpurun_out/s4c_mc_hvi_tiled_kernel -f \
      python tools/probe_ozaki.py full > gpurun_out/s4c_ncu_mc.log 2>&1
ls -la gpurun_out/s4c_mc_hvi_tiled_kernel.ncu-rep
