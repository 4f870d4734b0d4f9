#!/bin/bash
# Session-4 experiment: cell pre-filter + single-point fast path in the many-cell (chunked) MC/HVI kernel on config 4
# (DTLZ2, 4 objectives, q = 8), parity tests, config-4 and config-3 benches.
EVEREST_MC_FAST=0 python tools/probe_mc.py base dtlz2
EVEREST_MC_FAST=1 python tools/probe_mc.py fast dtlz2
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload dtlz2 > gpurun_out/s4_bench_dtlz2.json 2> gpurun_out/s4_bench_dtlz2.err
python bench.py > gpurun_out/s4e_bench_n1.json 2> gpurun_out/s4e_bench_n1.err
