# ncu evidence for the INT8 digit-plane GEMM and the fused cross-covariance + slicing kernel (round 1, session 2)
set -x
python bench.py > gpurun_out/bench_r1_s2_final_n1.json 2> gpurun_out/bench_r1_s2_final_n1.err
ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm_kernel -s 4 -c 1 -o gpurun_out/s2_ozaki_final -f python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:crosscov_fast_kernel -s 8 -c 1 -o gpurun_out/s2_crosscov_final -f python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r01_s2_final_launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out/*final*
