#!/bin/bash
# Session-3 ncu evidence (run on the GPU box after the plain bench has exited 0): launch list of the bench command and one
# full-set capture each of the dominant kernels of the device-resident screen (probe_ozaki.py full = config 3).
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench_plain.json 2> gpurun_out/s3_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/s3_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/s3_ncu_bench.log 2>&1
for k in ozaki_gemm2p_kernel crosscov_fast_kernel mc_hvi_tiled_kernel cond_root_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/s3_$k -f \
      python tools/probe_ozaki.py full > gpurun_out/s3_ncu_$k.log 2>&1
  ls -la gpurun_out/s3_$k.ncu-rep
done
