for g in 1 4 8 16; do
  EVEREST_GEMM_GROUPS=$g python bench.py --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('groups', $g, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'gemm_ms', round(d['roofline']['launch_ms'],3), 'frac', round(d['roofline']['frac'],4), d['roofline']['step_time_share'])
"
done
for g in 1 8; do
EVEREST_GEMM_GROUPS=$g ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:posterior_gemm_tma -c 3 --csv --log-file gpurun_out/ncu_gemm_groups_$g.csv python bench.py --no-cpu-baseline --steps 1 --warmup 3 > /dev/null 2>&1
tail -12 gpurun_out/ncu_gemm_groups_$g.csv
done
