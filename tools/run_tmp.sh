python -m pytest tests/test_gpu_parity.py -q -k "locality or overlapped or tetra or parallelepiped or fixtures" 2>&1 | tail -3
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { tag=$1; shift; $B "$@" 2> gpurun_out/s3_$tag.err > gpurun_out/s3_$tag.json; python -c "
import json,sys; d=json.loads(open('gpurun_out/s3_$tag.json').read().strip().splitlines()[-1]); u=d.get('uniform_grid') or {}
print('$tag', round(d['value']/1e6,2), 'M el/s', round(d['ms_per_step'],3), 'ms; kernel', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'traffic', d['roofline'].get('traffic'), '| uniform', round(u.get('value',0)/1e6,1), u.get('ms_per_step'))"; }
run c3_loc1 --phys elasticity --topo tet --grid 113
run c3_loc0 --phys elasticity --topo tet --grid 113 --locality 0
run c5_loc1 --phys elasticity --grid 80
run c2_loc1 --grid 128
