#!/bin/bash
# quick one-GPU pass: the tests around the kernels touched last + the headline configuration (device-resident timing only)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zzz_sumfact.py tests/test_gpu_zzzzz_round2.py tests/test_gpu_parity.py -q -p no:cacheprovider --tb=short -x -k "${QUICK_K:-gather or sumfact or hex}" > gpurun_out/r02_quick_tests.log 2>&1
tail -4 gpurun_out/r02_quick_tests.log
rm -f gpurun_out/r02_quick_bench.jsonl
for args in "${@:-"--topo hex --p 2 --phys poisson --grid 128"}"; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-extra --no-cpu-baseline --no-e2e $args 2>>gpurun_out/r02_quick_bench.err | tee -a gpurun_out/r02_quick_bench.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config']['workload'][:50], '%.1f M el/s' % (d['value'] / 1e6), '%.3f ms' % d['ms_per_step'], 'kernel %.3f ms' % d['roofline']['kernel_ms'], d['roofline']['kernel'][:40])
"
done
