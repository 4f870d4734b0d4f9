# ncu --set full captures of the kernels added / changed in the second session of round 1 (one launch each)
set -x
ncu --set full --clock-control none --import-source on -k regex:posterior_gemm_tma -s 4 -c 1 -o gpurun_out/s2_gemm -f python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
for k in skinny_gemm_kernel kernel_grad_kernel mc_hvi_grad_kernel cond_root_bwd_kernel grad_reduce_kernel small_gram_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/s2_$k -f python tools/probe_grad.py > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:mc_loghvi_kernel -s 1 -c 1 -o gpurun_out/s2_mc_loghvi -f python tools/probe_family.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:mll_grad_kernel -s 2 -c 1 -o gpurun_out/s2_mll_grad -f python tools/probe_fit.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_s2_bench_launches.csv python bench.py --no-cpu-baseline --steps 2 --warmup 3 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
