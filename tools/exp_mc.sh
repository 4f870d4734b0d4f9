#!/bin/bash
# Session-4 experiment + evidence: MC/HVI staging switches, parity tests, plain bench, ncu launch list of the bench command,
# full-set captures of the two kernels changed this session.
EVEREST_MC_STAGE=0 EVEREST_MC_PREFETCH=0 python tools/probe_mc.py base
EVEREST_MC_STAGE=1 EVEREST_MC_PREFETCH=0 python tools/probe_mc.py stage
EVEREST_MC_STAGE=0 EVEREST_MC_PREFETCH=1 python tools/probe_mc.py prefetch
EVEREST_MC_STAGE=1 EVEREST_MC_PREFETCH=1 python tools/probe_mc.py both
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/s4b_bench_n1.json 2> gpurun_out/s4b_bench_n1.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s4_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s4_ncu_bench.log 2>&1
for k in mc_hvi_tiled_kernel cond_root_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/s4_$k -f \
      python tools/probe_ozaki.py full > gpurun_out/s4_ncu_$k.log 2>&1
  ls -la gpurun_out/s4_$k.ncu-rep
done
