#!/bin/bash
# one GPU: the gather kernels (their tests and the drop-in case that gathers, full failure text), then timing on uniform C2 / C5 and C3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zzzzz_round2.py -q -p no:cacheprovider --tb=short -k "gather or baseline_sizes" > gpurun_out/r02_gather_tests.log 2>&1
tail -5 gpurun_out/r02_gather_tests.log
rm -f gpurun_out/r02_gather_bench.jsonl gpurun_out/r02_gather_bench.err
for args in "--topo hex --p 2 --phys poisson --grid 128 --perturb 0" "--topo hex --p 2 --phys elasticity --grid 81 --perturb 0" "--topo hex --p 1 --phys elasticity --grid 128 --perturb 0"; do
  for g in 1; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-e2e --gather $g $args 2>>gpurun_out/r02_gather_bench.err | tee -a gpurun_out/r02_gather_bench.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('gather=$g', d['config']['workload'][:50], '%.1f M el/s' % (d['value'] / 1e6), '%.3f ms' % d['ms_per_step'], 'kernel %.3f ms' % d['roofline']['kernel_ms'])
"
  done
done
