B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
run() { tag=$1; shift; $B "$@" 2> gpurun_out/s3_$tag.err > gpurun_out/s3_$tag.json; python -c "
import json,sys; d=json.loads(open('gpurun_out/s3_$tag.json').read().strip().splitlines()[-1]); u=d.get('uniform_grid') or {}
print('$tag', round(d['value']/1e6,2), 'M el/s', round(d['ms_per_step'],3), 'ms; kernel', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'hbm', round(d['roofline']['hbm_GBps_algorithmic']), '| uniform', round(u.get('value',0)/1e6,1), u.get('ms_per_step'))"; }
run c5_g64_v3 --phys elasticity --grid 64 --variant 3
run c5_g64_v0 --phys elasticity --grid 64
run c4_g32_v1 --p 4 --grid 32 --variant 1
run c4_g32_v0 --p 4 --grid 32
for pin in 0 1; do tests/_bin/dropin_test 48 2 0 0 1 0 16 0 0 $pin | tail -1 | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('dropin 48^3 p2 poisson pin', r['pin_host'], 'nnz', r['nnz'], 'gpu first/second assemble s', r['gpu_first_assemble_s'], r['gpu_second_assemble_s'], 'cpu threaded', r['cpu_threaded_assemble_s'], 'ok', r['ok'])"; done
python -m pytest tests/test_gpu_dropin.py -q 2>&1 | tail -3
