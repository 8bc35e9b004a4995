python -m pytest tests -m gpu -q 2>&1 | tail -6
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { tag=$1; shift; $B "$@" 2> gpurun_out/s3_$tag.err > gpurun_out/s3_$tag.json; python -c "
import json,sys; d=json.loads(open('gpurun_out/s3_$tag.json').read().strip().splitlines()[-1]); u=d.get('uniform_grid') or {}
print('$tag', round(d['value']/1e6,1), 'M el/s', round(d['ms_per_step'],3), 'ms; kernel', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'hbm', round(d['roofline']['hbm_GBps_algorithmic']), '| uniform', round(u.get('value',0)/1e6,1), u.get('ms_per_step'))"; }
run c3_g64 --phys elasticity --topo tet --grid 64
run c2_g128 --grid 128
run c5_g64 --phys elasticity --grid 64
run c1_g128 --p 1 --grid 128
run c1e_g96 --p 1 --phys elasticity --grid 96
