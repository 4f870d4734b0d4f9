#!/bin/bash
# Session-4 experiment: up-front Fp loads of the MC/HVI tiled kernel (EVEREST_MC_FPBATCH), parity tests, plain bench.
EVEREST_MC_FPBATCH=0 python tools/probe_mc.py base
EVEREST_MC_FPBATCH=1 python tools/probe_mc.py fpbatch
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/s4d_bench_n1.json 2> gpurun_out/s4d_bench_n1.err || exit 1
