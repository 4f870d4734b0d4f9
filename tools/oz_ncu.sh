ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm_kernel -s 3 -c 1 -o gpurun_out/s2_ozaki -f python tools/probe_ozaki.py full > /dev/null 2>&1
ls -la gpurun_out/s2_ozaki.ncu-rep
for g in 4 10 16; do EVEREST_OZAKI_GROUPS=$g python tools/probe_ozaki.py full 2>&1 | grep "ozaki=1"; done
