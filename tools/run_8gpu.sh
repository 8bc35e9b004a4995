run() { tag=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 "$@" > gpurun_out/r01_final3_n8_$tag.json 2> gpurun_out/r01_final3_n8_$tag.err; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/r01_final3_n8_$tag.json').read().strip().splitlines()[-1]); print('$tag', d['config']['workload'][:70], '| el/s %.4g dof/s %.4g ms %.3f frac %.3f'%(d['value'],d['dof_per_s'],d['ms_per_step'],d['roofline']['frac']), 'dof/gpu', d['config']['dof_per_gpu'])
except Exception as e: print('$tag FAILED', e); print(open('gpurun_out/r01_final3_n8_$tag.err').read()[-800:])
"; }
run c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --phys elasticity --grid 81
run c3 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --phys elasticity --topo tet --grid 80
run c5uniform --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --phys elasticity --grid 81 --perturb 0
