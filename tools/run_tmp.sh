python -m pytest tests -m gpu -q -x 2>&1 | tail -4
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { tag=$1; shift; $B "$@" 2> gpurun_out/s3_$tag.err > gpurun_out/s3_$tag.json; python -c "
import json,sys; d=json.loads(open('gpurun_out/s3_$tag.json').read().strip().splitlines()[-1]); u=d.get('uniform_grid') or {}
print('$tag', round(d['value']/1e6,1), 'M el/s', round(d['ms_per_step'],3), 'ms; kernel', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'hbm', round(d['roofline']['hbm_GBps_algorithmic']), '| uniform', round(u.get('value',0)/1e6,1), u.get('ms_per_step'))"; }
run c5_g64_rt --phys elasticity --grid 64
run c1e_g96_rt --p 1 --phys elasticity --grid 96
