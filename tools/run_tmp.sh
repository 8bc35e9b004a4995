python -m pytest tests -m gpu -q 2>&1 | tail -4
python bench.py > gpurun_out/r01_final2_c2.json 2> gpurun_out/r01_final2_c2.err; tail -c 300 gpurun_out/r01_final2_c2.err
B="python bench.py --steps 5 --warmup 3 --no-e2e"
run() { tag=$1; shift; $B "$@" 2> gpurun_out/r01_final2_$tag.err > gpurun_out/r01_final2_$tag.json; python -c "
import json,sys; d=json.loads(open('gpurun_out/r01_final2_$tag.json').read().strip().splitlines()[-1]); u=d.get('uniform_grid') or {}; c=d.get('cpu_baseline') or {}
print('$tag', round(d['value']/1e6,2), 'M el/s', round(d['ms_per_step'],3), 'ms; kernel', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'hbm', round(d['roofline']['hbm_GBps_algorithmic']), '| uniform', round(u.get('value',0)/1e6,1), u.get('ms_per_step'), '| cpu', c.get('value'), c.get('cores'))"; }
run c3 --phys elasticity --topo tet --grid 113
run c5share --phys elasticity --grid 80
run c1 --p 1 --grid 32
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r01_final2_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
grep -c assemble gpurun_out/r01_final2_launches_c2.csv
bash tools/ncu_capture.sh r01_final2_team_hexp2elast assemble_gram_team --phys elasticity --grid 48 2>&1 | tail -2
