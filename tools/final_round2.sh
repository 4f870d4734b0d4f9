#!/bin/bash
# Final measurement batch of round 2 (one gpurun call): GPU test suite, the four BASELINE workloads through bench.py, the launch
# list of the headline bench command and `ncu --set full` captures of the kernels that changed in the last session.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; tail -1 gpurun_out/r02_final_pytest_gpu.log
for w in zdt1 himmelblau dtlz2 mixed; do
  timeout 600 python bench.py --workload $w > gpurun_out/r02_final_bench_$w.json 2> gpurun_out/r02_final_bench_$w.err
  echo "$w rc=$?"; tail -c 300 gpurun_out/r02_final_bench_$w.json | head -c 10 > /dev/null
done
timeout 300 python bench.py --no-cpu-baseline --no-ask --steps 2 --warmup 3 > gpurun_out/r02_final_plain.json 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_final_launches.csv \
  python bench.py --no-cpu-baseline --no-ask --steps 2 --warmup 3 > gpurun_out/r02_final_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 300 bash tools/ncu_capture.sh mixed crosscov_kernel2 r02_ncu_crosscov2_mixed 2 1
timeout 300 bash tools/ncu_capture.sh dtlz2 cond_root_kernel r02_ncu_condroot_dtlz2 5 1
